#!/usr/bin/env bash
# gpu_check3.sh <tag> — tests, bench with PDL on / off, event timeline, init phases (1 GPU)
set -u
TAG=${1:-chk3}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_tests.log" 2>&1; echo "tests rc=$?"; tail -6 "$OUT/${TAG}_tests.log"
timeout 400 python bench.py --no-cpu-baseline > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"; echo "bench rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench.json"; tail -3 "$OUT/${TAG}_bench.err"
RLB_PDL=0 timeout 400 python bench.py --no-cpu-baseline > "$OUT/${TAG}_bench_nopdl.json" 2> "$OUT/${TAG}_bench_nopdl.err"; echo "bench nopdl rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_nopdl.json"
timeout 400 python bench.py --no-cpu-baseline --steps 20 --warmup 3 > "$OUT/${TAG}_bench_k20.json" 2> "$OUT/${TAG}_bench_k20.err"; echo "bench k20 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_k20.json"
timeout 200 python scripts/trace_iter.py > "$OUT/${TAG}_event_timeline.txt" 2>&1; head -34 "$OUT/${TAG}_event_timeline.txt"
RLB_INIT_PROFILE=1 timeout 200 python scripts/init_profile.py > "$OUT/${TAG}_init_profile.txt" 2>&1; tail -24 "$OUT/${TAG}_init_profile.txt"
echo done
