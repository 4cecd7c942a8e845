#!/usr/bin/env bash
# measure_round.sh — everything profiles/ needs from ONE gpurun call (1 GPU):
#
#   gpurun --timeout 2700 -- 'bash scripts/measure_round.sh r2'
#
#   1. pytest -m gpu                                   -> gpurun_out/<tag>_tests.log
#   2. python bench.py (default K/W)                   -> gpurun_out/<tag>_bench_default.json   (the judged line; NOT under a profiler)
#   3. python bench.py --impl reference                -> gpurun_out/<tag>_bench_reference.json
#   4. ncu launch list of a short bench run            -> gpurun_out/<tag>_launches.csv + <tag>_launch_summary.txt
#   5. ncu --set full of the dominant kernels          -> gpurun_out/<tag>_full.ncu-rep + <tag>_ncu_full.txt + hist_root_traffic.json
#   6. event timeline without a profiler               -> gpurun_out/<tag>_event_timeline.txt
#   7. compute-sanitizer memcheck + racecheck (small)  -> gpurun_out/<tag>_memcheck.log, <tag>_racecheck.log
#   8. GPU vs oracle in lockstep on the FULL workload  -> gpurun_out/<tag>_full_size_parity.txt
# Afterwards, here:  cp gpurun_out/<tag>_{bench_default.json,bench_reference.json,launches.csv,launch_summary.txt,ncu_full.txt,event_timeline.txt} profiles/
#                    cp gpurun_out/hist_root_traffic.json profiles/
# Every step is bounded by its own timeout so that one hanging step cannot take the box (and a strike) with it.
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
step() { echo "== $1" | tee -a "$OUT/${TAG}_steps.log"; }

step "1 tests";      timeout 300 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_tests.log" 2>&1; tail -2 "$OUT/${TAG}_tests.log"
step "2 bench";      timeout 300 python bench.py > "$OUT/${TAG}_bench_default.json" 2> "$OUT/${TAG}_bench_default.err"; cut -c1-400 "$OUT/${TAG}_bench_default.json"
step "3 reference";  timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > "$OUT/${TAG}_bench_reference.json" 2> "$OUT/${TAG}_bench_reference.err"
step "4 launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_ncu_l.log" 2>&1
python scripts/launch_breakdown.py "$OUT/${TAG}_launches.csv" 4 > "$OUT/${TAG}_launch_summary.txt" 2>&1; head -12 "$OUT/${TAG}_launch_summary.txt"
step "5 full capture"
RLB_NO_GRAPH=1 timeout 420 ncu --set full --clock-control none --import-source on \
    -k regex:"k_hist_root|k_hist_child|k_query_warp|k_query_block|k_chain_sim|k_leaf_chain|k_part_fused|k_finish" -s 20 -c 36 \
    -f -o "$OUT/${TAG}_full" python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_ncu_f.log" 2>&1
python scripts/summarise_ncu.py "$OUT/${TAG}_full.ncu-rep" "$OUT/${TAG}_ncu_full.txt" "$OUT/hist_root_traffic.json" > /dev/null 2>&1
step "6 timeline";   timeout 200 python scripts/trace_iter.py > "$OUT/${TAG}_event_timeline.txt" 2>&1; head -12 "$OUT/${TAG}_event_timeline.txt"
step "7 compute-sanitizer (memcheck, then racecheck; small run, graph off)"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > "$OUT/${TAG}_memcheck.log" 2>&1; echo "memcheck rc=$?"; tail -3 "$OUT/${TAG}_memcheck.log"
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > "$OUT/${TAG}_racecheck.log" 2>&1; echo "racecheck rc=$?"; tail -3 "$OUT/${TAG}_racecheck.log"
step "8 full-size lockstep with the oracle (north_star acceptance line)"
timeout 900 python scripts/full_size_parity.py --trees 10 > "$OUT/${TAG}_full_size_parity.txt" 2>&1; tail -4 "$OUT/${TAG}_full_size_parity.txt"
step done
