"""Reduce `ncu --page raw --csv` output to the columns the roofline discussion needs.

  python scripts/ncu_summarize.py raw.csv [--require REGEX ...] [--meta out.json key=value ...] > summary.csv

--require: fail (exit 3) unless at least one kernel name matches each REGEX - the round-1 capture described a kernel
that was no longer the shipped one; `--require 'conv_tc_persist_kernel<\\d+, 2>'` makes that impossible to repeat.
--meta: write a sidecar JSON (kernel mix of the capture, csrc hash / commit of the build it was taken from).
"""
import collections
import csv
import json
import re
import sys

KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "lts__t_sector_hit_rate.pct"]


def main():
    args = sys.argv[1:]
    path = args.pop(0)
    require, meta_path, meta = [], None, {}
    while args:
        a = args.pop(0)
        if a == "--require":
            require.append(args.pop(0))
        elif a == "--meta":
            meta_path = args.pop(0)
        elif "=" in a:
            k, v = a.split("=", 1)
            meta[k] = v
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    cols = [i for i, h in enumerate(hdr) if h in KEEP or "tensor" in h.lower()]
    body = [r for r in rows[hdr_i + 2:] if len(r) >= len(hdr)]
    kn = hdr.index("Kernel Name")
    names = [r[kn].replace("(int)", "") for r in body]   # ncu prints template arguments as "(int)128" in some pages
    for rx in require:
        if not any(re.search(rx, n) for n in names):
            sys.stderr.write("ncu_summarize: no kernel matches required pattern %r (capture is not of the shipped kernel mix)\n" % rx)
            sys.exit(3)
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] for i in cols])
    w.writerow([units[i] for i in cols])
    for r in body:
        w.writerow([r[i] for i in cols])
    if meta_path:
        mix = collections.Counter(re.sub(r"^.*?(\w+<[^>]*>|\w+)\(.*$", r"\1", n) if "(" in n else n for n in names)
        meta["kernel_mix"] = dict(mix)
        meta["launches"] = len(body)
        json.dump(meta, open(meta_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
