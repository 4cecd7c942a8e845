#!/usr/bin/env bash
# gpu_final_multi.sh <tag> <N> [full] — the N-GPU lines of the round, one box session:
#   C2 at the driver's flags (--steps 20 --warmup 5), the same with RLB_XW_SEQ_LOADS=1 (A/B of the batched NVLink loads);
#   with "full": bit identity (mgpu_check), C4 and C5 (bag-parallel) as well; with "tests": the >= 2-GPU pytest cases.
set -u
TAG=${1:-multi}
N=${2:-8}
MODE=${3:-}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
comm() { python -c "import json,sys; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('  comm', d.get('comm'))"; }
timeout 300 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 5 > "$OUT/${TAG}_bench_c2_n${N}_k20.json" 2> "$OUT/${TAG}_bench_c2_n${N}_k20.err"; echo "k20 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n${N}_k20.json"; comm "$OUT/${TAG}_bench_c2_n${N}_k20.json"
RLB_XW_SEQ_LOADS=1 timeout 300 $TR --master-port 29615 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > "$OUT/${TAG}_bench_c2_n${N}_k20_seqloads.json" 2> "$OUT/${TAG}_bench_c2_n${N}_k20_seqloads.err"; echo "k20 seq-loads rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n${N}_k20_seqloads.json"; comm "$OUT/${TAG}_bench_c2_n${N}_k20_seqloads.json"
if [ "$MODE" = "full" ]; then
  timeout 300 $TR --master-port 29616 bench.py --gpus $N --steps 100 --warmup 5 > "$OUT/${TAG}_bench_c2_n${N}.json" 2> "$OUT/${TAG}_bench_c2_n${N}.err"; echo "k100 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n${N}.json"
  timeout 300 $TR --master-port 29617 bench.py --gpus $N --workload c4 --steps 50 --warmup 5 > "$OUT/${TAG}_bench_c4_n$N.json" 2> "$OUT/${TAG}_bench_c4_n$N.err"; echo "c4 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c4_n$N.json"
  timeout 300 $TR --master-port 29618 bench.py --gpus $N --workload c5 --steps 8 --warmup 2 > "$OUT/${TAG}_bench_c5_n$N.json" 2> "$OUT/${TAG}_bench_c5_n$N.err"; echo "c5 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c5_n$N.json"
  timeout 200 $TR --master-port 29611 scripts/mgpu_check.py 0.05 4 > "$OUT/${TAG}_mgpu_n$N.log" 2>&1; echo "mgpu rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" "$OUT/${TAG}_mgpu_n$N.log" | tail -5
fi
if [ "$MODE" = "tests" ]; then
  timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > "$OUT/${TAG}_tests_multi.log" 2>&1; tail -2 "$OUT/${TAG}_tests_multi.log"
  for M in 2; do
    TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1"
    timeout 300 $TR2 --master-port 29624 bench.py --gpus $M --steps 20 --warmup 5 --no-cpu-baseline > "$OUT/${TAG}_bench_c2_n${M}_k20.json" 2> "$OUT/${TAG}_bench_c2_n${M}_k20.err"; echo "n=$M k20 rc=$?"; python scripts/brief.py "$OUT/${TAG}_bench_c2_n${M}_k20.json"; comm "$OUT/${TAG}_bench_c2_n${M}_k20.json"
  done
fi
echo done
