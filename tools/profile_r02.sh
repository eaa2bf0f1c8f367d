#!/bin/bash
# Round-2 evidence in one GPU call: launch list of one benchmark step (ncu, durations + DRAM bytes), ncu --set full captures
# of the kernels this round changed (fused QKV + attention at the stage-0/1/2 widths, attention forward / backward, the GELU
# and qkv GEMMs), summaries under gpurun_out/<R>_*.  usage: profile_r02.sh [tag]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${1:-r02}
bash tools/profile_step.sh $R
N="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, skip, count, command...
  name=$1; k=$2; s=$3; c=$4; shift 4
  timeout 300 $N -k regex:$k -s $s -c $c -o gpurun_out/${R}_$name "$@" > gpurun_out/${R}_$name.log 2>&1
  echo "== $name exit=$?"
  rep=gpurun_out/${R}_$name.ncu-rep
  if [ -f $rep ]; then
    python tools/ncu_extract.py $rep > gpurun_out/${R}_$name.summary.txt 2>&1
    ncu -i $rep --page details 2>/dev/null | grep -vE "^ *-+$" > gpurun_out/${R}_$name.details.txt
    ncu -i $rep --page source --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_hot.py /tmp/src.csv 30 > gpurun_out/${R}_$name.hot.txt 2>&1
    rm -f $rep
  fi
}
# attn_qkv_bench.py 1 N: per case one "fused + qkv" launch, then one "fused" launch of attn_qkv_fwd_kernel
cap attn_qkv_s0      attn_qkv_fwd 1 1 python tools/attn_qkv_bench.py 1 1
cap attn_qkv_s1      attn_qkv_fwd 5 1 python tools/attn_qkv_bench.py 1 3
cap attn_qkv_s2      attn_qkv_fwd 7 1 python tools/attn_qkv_bench.py 1 4
cap attn_s0          attn_tc 2 2 python tools/attn_bench.py 2 1
cap attn_s2          attn_tc 14 2 python tools/attn_bench.py 2 4
cap gemm_fc1_gelu_s2 gemm_tc 1 1 python tools/gemm_bench.py fc1_gelu_s2 2
cap gemm_fc1_gelu_s0 gemm_tc 1 1 python tools/gemm_bench.py fc1_gelu_s0 2
cap gemm_qkv_s2      gemm_tc 1 1 python tools/gemm_bench.py qkv_s2 2
ls gpurun_out/${R}_*summary.txt
