#!/usr/bin/env bash
# gpu_variants.sh — one short GPU-box call for the histogram-kernel variants (RLB_HIST_VARIANT):
#   1. scripts/variant_bench.py: the variants side by side on the full C2 workload (time per iteration, root / child
#      histogram time, tree CRC against variant 0)          -> gpurun_out/<tag>_variants.jsonl
#   2. pytest -m gpu with the variant under test switched on (parity against the oracle, full-size lockstep included)
#                                                            -> gpurun_out/<tag>_tests_v<variant>.log
# Usage: gpurun --timeout 400 -- 'bash scripts/gpu_variants.sh r2v 1'
set -u
TAG=${1:-rXv}
V=${2:-1}
LIST=${3:-0,1}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
echo "== variants"
timeout 240 python scripts/variant_bench.py "$LIST" 3 20 > "$OUT/${TAG}_variants.jsonl" 2> "$OUT/${TAG}_variants.err"
echo "rc=$?"; cat "$OUT/${TAG}_variants.jsonl"; tail -3 "$OUT/${TAG}_variants.err"
if [ "$V" != "notest" ]; then
echo "== tests with RLB_HIST_VARIANT=$V"
RLB_HIST_VARIANT=$V timeout 240 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_tests_v${V}.log" 2>&1
echo "rc=$?"; tail -3 "$OUT/${TAG}_tests_v${V}.log"
fi
echo "== done"
