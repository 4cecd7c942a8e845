#!/usr/bin/env bash
# gpu_check4.sh <tag> — GPU tests, the default bench line (with CPU baseline + parity), the 20-step line, the event timeline (1 GPU)
set -u
TAG=${1:-chk4}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_tests.log" 2>&1; echo "tests rc=$?"; tail -4 "$OUT/${TAG}_tests.log"
timeout 400 python bench.py > "$OUT/${TAG}_bench_default.json" 2> "$OUT/${TAG}_bench_default.err"; echo "bench rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_default.json"; tail -3 "$OUT/${TAG}_bench_default.err"
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > "$OUT/${TAG}_bench_k20.json" 2> "$OUT/${TAG}_bench_k20.err"; echo "bench k20 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_k20.json"
timeout 200 python scripts/trace_iter.py > "$OUT/${TAG}_event_timeline.txt" 2>&1; head -30 "$OUT/${TAG}_event_timeline.txt"
echo done
