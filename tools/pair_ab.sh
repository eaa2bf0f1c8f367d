#!/bin/bash
# A/B of the CTA-pair GEMM: correctness with the pair kernels forced, then the micro-benchmark with them off / on.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SWIN_GEMM_PAIR=2 timeout 300 python tools/pair_check.py > gpurun_out/pair_check.log 2>&1
echo "pair_check exit=$?"; tail -n 45 gpurun_out/pair_check.log
if [ "$1" != "check" ]; then
  SWIN_GEMM_PAIR=0 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_pair0.log 2>&1; echo "bench0 exit=$?"
  SWIN_GEMM_PAIR=1 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_pair1.log 2>&1; echo "bench1 exit=$?"
  paste -d'\n' gpurun_out/gemm_pair0.log gpurun_out/gemm_pair1.log
fi
