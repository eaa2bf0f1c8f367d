#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
P=$PWD/swin-transformer-object-detection_b200
for rep in 1 2; do
  echo "== before"; SWIN_B200_LIB=$P/libswin_b200_base.so python tools/ln_bench.py ln_bwd 7 2>&1 | grep -E "emit|gather"
  echo "== scale prefetch"; python tools/ln_bench.py ln_bwd 7 2>&1 | grep -E "emit|gather"
done | tee gpurun_out/ln_bwd_ab.txt
echo "== w12 before"; SWIN_B200_LIB=$P/libswin_b200_base.so python tools/attn_w12_bench.py 2>&1 | tail -8
echo "== w12 after"; python tools/attn_w12_bench.py 2>&1 | tail -8
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
