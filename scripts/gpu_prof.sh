#!/usr/bin/env bash
# gpu_prof.sh <tag> — ncu --set full captures of the lambda kernels and the split-step kernels at full size (1 GPU)
set -u
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p "$OUT"
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
RLB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_query" -s 5 -c 5 -f -o "$OUT/${TAG}_lambda" \
    python scripts/prof_iter.py 1.0 3 > "$OUT/${TAG}_ncu_lambda.log" 2>&1; echo "lambda rc=$?"; tail -3 "$OUT/${TAG}_ncu_lambda.log"
ncu -i "$OUT/${TAG}_lambda.ncu-rep" --page raw --csv > "$OUT/${TAG}_lambda_raw.csv" 2>/dev/null
ncu -i "$OUT/${TAG}_lambda.ncu-rep" --page source --csv > "$OUT/${TAG}_lambda_source.csv" 2>/dev/null
RLB_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hist_child|k_part_fused|k_finish|k_leaf_chain|k_chain_sim" -s 27 -c 12 -f -o "$OUT/${TAG}_split" \
    python scripts/prof_iter.py 1.0 3 > "$OUT/${TAG}_ncu_split.log" 2>&1; echo "split rc=$?"; tail -3 "$OUT/${TAG}_ncu_split.log"
ncu -i "$OUT/${TAG}_split.ncu-rep" --page raw --csv > "$OUT/${TAG}_split_raw.csv" 2>/dev/null
ls -la "$OUT" | grep "$TAG"
echo done
