#!/usr/bin/env bash
# gpu_final_single.sh — the short form of measure_round.sh for the last GPU minutes of a round (1 GPU, ~3 min):
#   1. pytest -m gpu                                      -> gpurun_out/<tag>_tests.log
#   2. python bench.py --steps 20 --warmup 5 (the driver's command, CPU baseline + parity included)
#                                                         -> gpurun_out/<tag>_bench_k20.json
#   3. ncu launch list of a short bench run               -> gpurun_out/<tag>_launches.csv + <tag>_launch_summary.txt
#   4. ncu --set full of the two histogram kernels        -> gpurun_out/<tag>_full.ncu-rep + <tag>_ncu_full.txt
# Every step has its own timeout.   gpurun --timeout 260 -- 'bash scripts/gpu_final_single.sh r2f'
set -u
TAG=${1:-rXf}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
T0=$(date +%s)
el() { echo "   [$(( $(date +%s) - T0 )) s]"; }
echo "== 1 tests";  timeout 90 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_tests.log" 2>&1; echo "rc=$?"; tail -2 "$OUT/${TAG}_tests.log"; el
echo "== 2 bench";  timeout 80 python bench.py --steps 20 --warmup 5 > "$OUT/${TAG}_bench_k20.json" 2> "$OUT/${TAG}_bench_k20.err"; echo "rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_k20.json"; el
echo "== 3 launch list"
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_ncu_l.log" 2>&1
python scripts/launch_breakdown.py "$OUT/${TAG}_launches.csv" 4 > "$OUT/${TAG}_launch_summary.txt" 2>&1; head -14 "$OUT/${TAG}_launch_summary.txt"; el
echo "== 4 full capture (histogram kernels)"
RLB_NO_GRAPH=1 timeout 45 ncu --set full --clock-control none --import-source on -k regex:"k_hist_root|k_hist_child" -s 0 -c 5 \
    -f -o "$OUT/${TAG}_full" python scripts/prof_iter.py 1.0 2 > "$OUT/${TAG}_ncu_f.log" 2>&1
python scripts/summarise_ncu.py "$OUT/${TAG}_full.ncu-rep" "$OUT/${TAG}_ncu_full.txt" "$OUT/${TAG}_hist_root_traffic.json" > /dev/null 2>&1; ls -la "$OUT/${TAG}_full.ncu-rep" 2>/dev/null; el
echo "== done"
