#!/usr/bin/env bash
# gpu_check_n8.sh <tag> <N> — the N-GPU lines of every BASELINE config: C3 (C2 sharded), C4, C5 (bag-parallel), + bit identity
set -u
TAG=${1:-n8}
N=${2:-8}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 scripts/mgpu_check.py 0.05 4 > "$OUT/${TAG}_mgpu_p2p.log" 2>&1; echo "mgpu p2p rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" "$OUT/${TAG}_mgpu_p2p.log" | tail -7
timeout 400 $TR --master-port 29613 bench.py --gpus $N --steps 100 --warmup 5 > "$OUT/${TAG}_bench_c2_n$N.json" 2> "$OUT/${TAG}_bench_c2_n$N.err"; echo "bench rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n$N.json"; grep -v "^W\|^\*\*\*\|OMP_NUM" "$OUT/${TAG}_bench_c2_n$N.err" | tail -3
timeout 400 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/${TAG}_bench_c2_n${N}_k20.json" 2> "$OUT/${TAG}_bench_c2_n${N}_k20.err"; echo "bench20 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n${N}_k20.json"
timeout 400 $TR --master-port 29615 bench.py --gpus $N --workload c4 --steps 50 --warmup 5 > "$OUT/${TAG}_bench_c4_n$N.json" 2> "$OUT/${TAG}_bench_c4_n$N.err"; echo "c4 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c4_n$N.json"
timeout 400 $TR --master-port 29616 bench.py --gpus $N --workload c5 --steps 8 --warmup 2 > "$OUT/${TAG}_bench_c5_n$N.json" 2> "$OUT/${TAG}_bench_c5_n$N.err"; echo "c5 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c5_n$N.json"; grep -v "^W\|^\*\*\*\|OMP_NUM" "$OUT/${TAG}_bench_c5_n$N.err" | tail -3
echo done
