#!/usr/bin/env bash
# gpu_check_multi.sh <tag> <N> — N-GPU bit-identity check (exchange window and NCCL fallback) and the N-GPU bench line.
set -u
TAG=${1:-mchk}
N=${2:-2}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
RLB_P2P_VERBOSE=1 timeout 300 $TR --master-port 29611 scripts/mgpu_check.py 0.05 6 > "$OUT/${TAG}_mgpu_p2p.log" 2>&1; echo "mgpu p2p rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" "$OUT/${TAG}_mgpu_p2p.log" | tail -12
RLB_P2P=0 timeout 300 $TR --master-port 29612 scripts/mgpu_check.py 0.05 4 > "$OUT/${TAG}_mgpu_nccl.log" 2>&1; echo "mgpu nccl rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" "$OUT/${TAG}_mgpu_nccl.log" | tail -7
timeout 400 $TR --master-port 29613 bench.py --gpus $N --steps 100 --warmup 5 > "$OUT/${TAG}_bench_n$N.json" 2> "$OUT/${TAG}_bench_n$N.err"; echo "bench rc=$?"; cut -c1-2500 "$OUT/${TAG}_bench_n$N.json"; grep -v "^W\|^\*\*\*\|OMP_NUM" "$OUT/${TAG}_bench_n$N.err" | tail -5
timeout 400 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 3 > "$OUT/${TAG}_bench_n${N}_k20.json" 2> "$OUT/${TAG}_bench_n${N}_k20.err"; echo "bench20 rc=$?"; cut -c1-1200 "$OUT/${TAG}_bench_n${N}_k20.json"
if [ "${3:-}" = "tests" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_tests.log" 2>&1; echo "tests rc=$?"; tail -8 "$OUT/${TAG}_tests.log"
fi
echo done
