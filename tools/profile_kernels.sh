#!/bin/bash
# ncu --set full captures of ONE launch each of the hot kernels, taken from the micro-benchmarks (same kernels, same
# shapes as the benchmark step, but a few hundred MB of device memory instead of 16 GB, so ncu's save/restore replay
# stays fast).  A capture of the whole step took 19 GPU-minutes and its 100 MB of reports could not be copied back.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${1:-r01}
N="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, skip, count, command...
  name=$1; k=$2; s=$3; c=$4; shift 4
  timeout 300 $N -k regex:$k -s $s -c $c -o gpurun_out/${R}_$name "$@" > gpurun_out/${R}_$name.log 2>&1
  echo "== $name exit=$?"
  # text summaries are what gets committed under profiles/; reports above 6 MB (whole-file source import) are dropped so
  # that gpurun_out/ stays under the 64 MiB copy-back limit
  rep=gpurun_out/${R}_$name.ncu-rep
  if [ -f $rep ]; then
    python tools/ncu_extract.py $rep > gpurun_out/${R}_$name.summary.txt 2>&1
    ncu -i $rep --page details 2>/dev/null | grep -vE "^ *-+$" > gpurun_out/${R}_$name.details.txt
    ncu -i $rep --page source --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_hot.py /tmp/src.csv 30 > gpurun_out/${R}_$name.hot.txt 2>&1
    if [ $(stat -c %s $rep) -gt 6000000 ]; then rm -f $rep; fi
  fi
}
cap gemm_qkv_s0      gemm_tc 1 1 python tools/gemm_bench.py qkv_s0 2
cap gemm_fc1_gelu_s2 gemm_tc 1 1 python tools/gemm_bench.py fc1_gelu_s2 2
cap gemm_dW_fc1_s2   gemm_tc 1 1 python tools/gemm_bench.py dW_fc1_s2 2
cap gemm_square_8k   gemm_tc 1 1 python tools/gemm_bench.py square_8k 2
cap attn_s0          attn_tc 2 2 python tools/attn_bench.py 2 1
cap ln_fwd_gather_s0 ln_fwd_kernel 2 1 python tools/ln_bench.py ln_fwd_gather_s0_shift3 2
cap ln_bwd_emit_s0   ln_bwd 1 1 python tools/ln_bench.py ln_bwd_plain_emit_s0 2
cap ln_bwd_gather_s0 ln_bwd 1 1 python tools/ln_bench.py ln_bwd_gather_s0_shift3 2
cap gemm_dgelu_s0    gemm_tc 1 1 python tools/gemm_bench.py dgelu_s0 2
cap window_gather_s0 row_map_copy 1 1 python tools/ln_bench.py window_gather_s0 2
cap ln_fwd_merge_s0  ln_fwd_kernel 4 1 python tools/ln_bench.py ln_fwd_merge_s0 2
ls -la gpurun_out/${R}_*; du -sh gpurun_out
