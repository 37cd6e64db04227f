"""Reduce `ncu --page raw --csv` output to the columns the roofline discussion needs."""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, units = rows[hdr_i], rows[hdr_i + 1]
cols = [i for i, h in enumerate(hdr) if h in KEEP or "tensor" in h.lower()]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in cols])
w.writerow([units[i] for i in cols])
for r in rows[hdr_i + 2:]:
    if len(r) >= len(hdr):
        w.writerow([r[i] for i in cols])
