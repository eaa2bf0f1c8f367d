#!/bin/bash
# One GPU call: all kernel/model parity groups, the benchmark of record with its per-kernel table, the step launch list and
# the ncu captures of the GEMM / attention kernels.  usage: round_validate.sh [tag]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
R=${1:-r01}
bash tools/gpu_check.sh index ln gemm32 gemm16 tiles attn32 attn16 attnfull optim model
python bench.py --breakdown gpurun_out/${R}_step_breakdown.txt > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench exit=$?"
cat gpurun_out/${R}_bench.json
bash tools/profile_step.sh $R
N="ncu --set full --clock-control none --import-source on -f"
cap() {
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
cap gemm_qkv_s2      gemm_tc 1 1 python tools/gemm_bench.py qkv_s2 2
cap gemm_fc1_gelu_s2 gemm_tc 1 1 python tools/gemm_bench.py fc1_gelu_s2 2
cap gemm_dW_fc1_s2   gemm_tc 1 1 python tools/gemm_bench.py dW_fc1_s2 2
cap gemm_square_8k   gemm_tc 1 1 python tools/gemm_bench.py square_8k 2
cap gemm_qkv_s0      gemm_tc 1 1 python tools/gemm_bench.py qkv_s0 2
cap attn_s0          attn_tc 2 2 python tools/attn_bench.py 2 1
