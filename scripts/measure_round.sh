#!/usr/bin/env bash
# measure_round.sh — everything profiles/ needs from ONE gpurun call (1 GPU):
#
#   gpurun --timeout 2700 -- 'bash scripts/measure_round.sh r2'
#
#   1. pytest -m gpu                                   -> gpurun_out/<tag>_tests.log
#   2. python bench.py --steps 20 --warmup 5 (what the driver runs) and the 100-step default
#                                                      -> gpurun_out/<tag>_bench_k20.json, <tag>_bench_default.json   (NOT under a profiler)
#   3. python bench.py --impl reference --steps 20 --warmup 5  -> gpurun_out/<tag>_bench_reference.json
#   4. side legs: --op eval, --workload c4, --workload c5      -> gpurun_out/<tag>_bench_{eval,c4,c5}.json
#   5. ncu launch list of a short bench run            -> gpurun_out/<tag>_launches.csv + <tag>_launch_summary.txt
#   6. ncu --set full of the dominant kernels          -> gpurun_out/<tag>_full.ncu-rep + <tag>_ncu_full.txt + hist_root_traffic.json
#   7. event timeline without a profiler               -> gpurun_out/<tag>_event_timeline.txt
#   8. compute-sanitizer memcheck + racecheck (small)  -> gpurun_out/<tag>_memcheck.log, <tag>_racecheck.log
# Afterwards, here: copy the files to be judged into profiles/.
# Every step is bounded by its own timeout so that one hanging step cannot take the box (and a strike) with it.
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
step() { echo "== $1" | tee -a "$OUT/${TAG}_steps.log"; }

step "1 tests";      timeout 600 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_tests.log" 2>&1; tail -2 "$OUT/${TAG}_tests.log"
step "2 bench";      timeout 300 python bench.py --steps 20 --warmup 5 > "$OUT/${TAG}_bench_k20.json" 2> "$OUT/${TAG}_bench_k20.err"; python scripts/brief.py "$OUT/${TAG}_bench_k20.json"
                     timeout 300 python bench.py > "$OUT/${TAG}_bench_default.json" 2> "$OUT/${TAG}_bench_default.err"; python scripts/brief.py "$OUT/${TAG}_bench_default.json"
step "3 reference";  timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/${TAG}_bench_reference.json" 2> "$OUT/${TAG}_bench_reference.err"; cut -c1-300 "$OUT/${TAG}_bench_reference.json"
step "4 side legs"
timeout 300 python bench.py --op eval > "$OUT/${TAG}_bench_eval.json" 2> "$OUT/${TAG}_bench_eval.err"; python scripts/brief.py "$OUT/${TAG}_bench_eval.json"
timeout 300 python bench.py --workload c4 --steps 50 --no-cpu-baseline > "$OUT/${TAG}_bench_c4.json" 2> "$OUT/${TAG}_bench_c4.err"; python scripts/brief.py "$OUT/${TAG}_bench_c4.json"
timeout 300 python bench.py --workload c5 --steps 10 --warmup 2 > "$OUT/${TAG}_bench_c5.json" 2> "$OUT/${TAG}_bench_c5.err"; python scripts/brief.py "$OUT/${TAG}_bench_c5.json"
step "5 launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file "$OUT/${TAG}_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_ncu_l.log" 2>&1
python scripts/launch_breakdown.py "$OUT/${TAG}_launches.csv" 4 > "$OUT/${TAG}_launch_summary.txt" 2>&1; head -12 "$OUT/${TAG}_launch_summary.txt"
step "6 full capture"
RLB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"k_hist_root|k_hist_child|k_query_warp|k_query_block|k_chain_sim|k_leaf_chain|k_part_fused|k_finish|k_ensemble_eval" -s 8 -c 40 \
    -f -o "$OUT/${TAG}_full" python scripts/prof_iter.py 1.0 3 > "$OUT/${TAG}_ncu_f.log" 2>&1
python scripts/summarise_ncu.py "$OUT/${TAG}_full.ncu-rep" "$OUT/${TAG}_ncu_full.txt" "$OUT/hist_root_traffic.json" > /dev/null 2>&1
step "7 timeline";   timeout 200 python scripts/trace_iter.py > "$OUT/${TAG}_event_timeline.txt" 2>&1; head -12 "$OUT/${TAG}_event_timeline.txt"
step "8 compute-sanitizer (memcheck, then racecheck; small run, graph off)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > "$OUT/${TAG}_memcheck.log" 2>&1; echo "memcheck rc=$?"; tail -3 "$OUT/${TAG}_memcheck.log"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_small.py > "$OUT/${TAG}_racecheck.log" 2>&1; echo "racecheck rc=$?"; tail -3 "$OUT/${TAG}_racecheck.log"
step done
