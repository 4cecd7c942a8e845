#!/usr/bin/env bash
# gpu_check.sh <tag> — the GPU tests, then the default bench line and the side legs (1 GPU); every step bounded by its own timeout.
set -u
TAG=${1:-chk}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
timeout 900 python -m pytest tests -m gpu -x -q -s > "$OUT/${TAG}_tests.log" 2>&1; echo "tests rc=$?"; tail -15 "$OUT/${TAG}_tests.log"; grep -h "PARITY\|full-size" "$OUT/${TAG}_tests.log" | tail -20
timeout 400 python bench.py > "$OUT/${TAG}_bench_default.json" 2> "$OUT/${TAG}_bench_default.err"; echo "bench rc=$?"; cut -c1-3000 "$OUT/${TAG}_bench_default.json"; tail -3 "$OUT/${TAG}_bench_default.err"
if [ "${2:-}" != "short" ]; then
timeout 300 python bench.py --op eval > "$OUT/${TAG}_bench_eval.json" 2> "$OUT/${TAG}_bench_eval.err"; echo "eval rc=$?"; cut -c1-1500 "$OUT/${TAG}_bench_eval.json"; tail -3 "$OUT/${TAG}_bench_eval.err"
timeout 300 python bench.py --workload c4 --steps 50 --no-cpu-baseline > "$OUT/${TAG}_bench_c4.json" 2> "$OUT/${TAG}_bench_c4.err"; echo "c4 rc=$?"; cut -c1-1500 "$OUT/${TAG}_bench_c4.json"; tail -3 "$OUT/${TAG}_bench_c4.err"
timeout 300 python bench.py --workload c5 --steps 10 --warmup 2 > "$OUT/${TAG}_bench_c5.json" 2> "$OUT/${TAG}_bench_c5.err"; echo "c5 rc=$?"; cut -c1-1500 "$OUT/${TAG}_bench_c5.json"; tail -3 "$OUT/${TAG}_bench_c5.err"
fi
echo done
