#!/usr/bin/env bash
# gpu_check2.sh <tag> — tests, bench, event timeline, init breakdown, warm-cache ncu launch list (1 GPU)
set -u
TAG=${1:-chk2}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
timeout 900 python -m pytest tests -m gpu -x -q -s > "$OUT/${TAG}_tests.log" 2>&1; echo "tests rc=$?"; tail -6 "$OUT/${TAG}_tests.log"; grep -h "VALIDATION\|full-size" "$OUT/${TAG}_tests.log" | tail
timeout 400 python bench.py --no-cpu-baseline > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"; echo "bench rc=$?"; cut -c1-2200 "$OUT/${TAG}_bench.json"; tail -3 "$OUT/${TAG}_bench.err"
timeout 200 python scripts/trace_iter.py > "$OUT/${TAG}_event_timeline.txt" 2>&1; head -34 "$OUT/${TAG}_event_timeline.txt"
timeout 200 python scripts/init_profile.py > "$OUT/${TAG}_init_profile.txt" 2>&1; tail -4 "$OUT/${TAG}_init_profile.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file "$OUT/${TAG}_launches_warm.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/${TAG}_ncu_l.log" 2>&1
python scripts/launch_breakdown.py "$OUT/${TAG}_launches_warm.csv" 4 > "$OUT/${TAG}_launch_summary_warm.txt" 2>&1; head -34 "$OUT/${TAG}_launch_summary_warm.txt"
echo done
