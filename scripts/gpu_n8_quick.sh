#!/usr/bin/env bash
# gpu_n8_quick.sh <tag> <N> — C2 bench line at N GPUs (100 and 20 steps) + per-launch timeline on rank 0
set -u
TAG=${1:-n8q}
N=${2:-8}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29613 bench.py --gpus $N --steps 100 --warmup 5 > "$OUT/${TAG}_bench_c2_n$N.json" 2> "$OUT/${TAG}_bench_c2_n$N.err"; echo "bench rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n$N.json"; python -c "import json,sys; d=json.loads(open('$OUT/${TAG}_bench_c2_n$N.json').read().strip().splitlines()[-1]); print(d.get('comm'))"
timeout 400 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/${TAG}_bench_c2_n${N}_k20.json" 2> "$OUT/${TAG}_bench_c2_n${N}_k20.err"; echo "bench20 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n${N}_k20.json"
timeout 300 $TR --master-port 29700 scripts/trace_iter_multi.py > "$OUT/${TAG}_trace_n$N.txt" 2>&1; grep -v "^W\|^\*\*\*\|OMP" "$OUT/${TAG}_trace_n$N.txt" | tail -34 | head -26
echo done
