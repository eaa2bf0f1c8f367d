#!/bin/bash
# A/B of the CTA-pair GEMM: correctness (1-CTA kernels, then pair kernels forced), optional role-wait profile, then the
# micro-benchmark with pair mode off / on.   usage: pair_ab.sh [check] [prof] [bench]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for a in "$@"; do
  case $a in
    check)
      SWIN_GEMM_PAIR=0 timeout 300 python tools/pair_check.py > gpurun_out/pair_check0.log 2>&1; echo "check(1-CTA) exit=$? $(tail -n 1 gpurun_out/pair_check0.log)"
      SWIN_GEMM_PAIR=2 timeout 300 python tools/pair_check.py > gpurun_out/pair_check2.log 2>&1; echo "check(pair)  exit=$? $(tail -n 1 gpurun_out/pair_check2.log)"
      grep -h FAIL gpurun_out/pair_check0.log gpurun_out/pair_check2.log | head -20 ;;
    prof)
      export SWIN_B200_LIB=$PWD/swin-transformer-object-detection_b200/libswin_b200_prof.so
      SWIN_GEMM_PAIR=0 timeout 120 python tools/gemm_prof.py > gpurun_out/prof0.log 2>&1
      SWIN_GEMM_PAIR=2 timeout 120 python tools/gemm_prof.py > gpurun_out/prof2.log 2>&1
      unset SWIN_B200_LIB
      cat gpurun_out/prof0.log gpurun_out/prof2.log ;;
    bench)
      SWIN_GEMM_PAIR=0 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_pair0.log 2>&1; echo "bench0 exit=$?"
      SWIN_GEMM_PAIR=2 timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_pair2.log 2>&1; echo "bench2 exit=$?"
      paste -d'\n' gpurun_out/gemm_pair0.log gpurun_out/gemm_pair2.log ;;
  esac
done
