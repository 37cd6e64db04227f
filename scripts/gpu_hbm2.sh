#!/bin/bash
# Round-2b HBM kernels (IRSDE_HBM_NEW bits: 1 LayerNorm, 2 la_combine, 4 la_fold, 8 k/v pass; IRSDE_LN_PP 1|2), ONE box:
#   probe (each bit vs the round-2a kernels) -> same-box ABAB of the settings -> driver-style line, full GPU suite and the
#   HBM-kernel ncu capture with the best passing setting.
R=${ROUND:-r02b}
mkdir -p gpurun_out
. image-restoration-sde_b200/BUILD_INFO 2>/dev/null
probe() { env IRSDE_HBM_NEW=$1 IRSDE_LN_PP=${2:-2} timeout 300 python scripts/hbm_probe.py run /tmp/probe_$1_${2:-2}.pt 2>&1 | tail -1; }
probe 0; probe 15 2; probe 15 1
GOOD=15; PP1=1
python scripts/hbm_probe.py cmp /tmp/probe_0_2.pt /tmp/probe_15_1.pt > gpurun_out/hbm_probe.txt 2>&1 || PP1=0
if ! python scripts/hbm_probe.py cmp /tmp/probe_0_2.pt /tmp/probe_15_2.pt >> gpurun_out/hbm_probe.txt 2>&1; then
  GOOD=0; PP1=0
  for b in 1 2 4 8; do
    probe $b
    if python scripts/hbm_probe.py cmp /tmp/probe_0_2.pt /tmp/probe_${b}_2.pt >> gpurun_out/hbm_probe.txt 2>&1; then GOOD=$((GOOD | b)); fi
  done
fi
cat gpurun_out/hbm_probe.txt | tail -40; echo "GOOD mask = $GOOD, PP=1 usable = $PP1"
[ "$GOOD" = 0 ] && { echo "no new kernel passed the probe"; exit 1; }

run() { # name mask pp
  env IRSDE_HBM_NEW=$2 IRSDE_LN_PP=$3 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/hbm_$1.log 2> gpurun_out/hbm_$1.err
  python - "$1" "$2" "$3" <<'PY'
import json, sys
for line in open("gpurun_out/hbm_%s.log" % sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        print("AB", sys.argv[1], "mask", sys.argv[2], "pp", sys.argv[3], "ms/chain", round(d["ms_per_step"], 1), d["clocks"]["sm_mhz"],
              {k: round(v["ms_per_step"], 3) for k, v in d["breakdown"].items()})
PY
}
M3=$((GOOD & 7))
for rep in a b; do
  run old_$rep 0 2
  run new_pp2_$rep $GOOD 2
  [ "$PP1" = 1 ] && run new_pp1_$rep $GOOD 1
  [ "$M3" != "$GOOD" ] && [ "$M3" != 0 ] && run nokv_$rep $M3 2
done 2>&1 | tee gpurun_out/hbm_abab.txt
# best setting = lowest mean ms/chain among the new ones
read BM BP <<< $(python - <<'PY'
import re, collections
t = collections.defaultdict(list)
for l in open("gpurun_out/hbm_abab.txt"):
    m = re.match(r"AB (\S+)_[ab] mask (\d+) pp (\d) ms/chain ([\d.]+)", l)
    if m and m.group(1) != "old": t[(m.group(2), m.group(3))].append(float(m.group(4)))
best = min(t, key=lambda k: sum(t[k]) / len(t[k]))
print(best[0], best[1])
PY
)
echo "best setting: IRSDE_HBM_NEW=$BM IRSDE_LN_PP=$BP" | tee -a gpurun_out/hbm_abab.txt
export IRSDE_HBM_NEW=$BM IRSDE_LN_PP=$BP
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/hbm_bench_c2.log 2> gpurun_out/hbm_bench_c2.err
grep '^{' gpurun_out/hbm_bench_c2.log | tail -1 | cut -c1-400
timeout 200 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/hbm_bench_c4.log 2> gpurun_out/hbm_bench_c4.err
grep '^{' gpurun_out/hbm_bench_c4.log | tail -1 | cut -c1-200
timeout 200 python bench.py --precision fp32x3 --steps 3 --warmup 3 --no-cpu > gpurun_out/hbm_bench_fp32x3.log 2> gpurun_out/hbm_bench_fp32x3.err
grep '^{' gpurun_out/hbm_bench_fp32x3.log | tail -1 | cut -c1-200
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/hbm_tests.log; tail -5 gpurun_out/hbm_tests.log
timeout 400 ncu --set full --clock-control none -k regex:"layernorm|la_kv|la_combine|la_fold|sde_update" -s 40 -c 40 -o /tmp/prof_hbm -f \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/hbm_ncu.log 2>&1
ncu -i /tmp/prof_hbm.ncu-rep --page raw --csv > /tmp/prof_hbm_raw.csv 2>/dev/null
python scripts/ncu_summarize.py /tmp/prof_hbm_raw.csv --meta gpurun_out/${R}_hbm_kernels_ncu_full.meta.json commit=$commit csrc_sha256=$csrc_sha256 \
  conv_tc_sha256=$conv_tc_sha256 IRSDE_HBM_NEW=$BM IRSDE_LN_PP=$BP > gpurun_out/${R}_hbm_kernels_ncu_full.csv 2> gpurun_out/hbm_ncu_sum.err
wc -l gpurun_out/${R}_hbm_kernels_ncu_full.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 300 --csv --log-file gpurun_out/${R}_launches_gpu_time.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-graph > gpurun_out/hbm_launches.log 2>&1
wc -l gpurun_out/${R}_launches_gpu_time.csv
