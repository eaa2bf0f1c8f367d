#!/bin/bash
# Round-end evidence: launch list of one benchmark step + ncu --set full captures of the stage-0 forward
# and backward kernels (the HBM-heavy end) and of the stage-2 GEMMs (the tensor-heavy end).  Outputs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${1:-r01}
python bench.py --ncu-step --warmup 3 > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${R}_launches_step.csv python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_a.log 2>&1
# forward, stage 0 + first block of stage 1 (first 45 kernels of the step)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tc|attn_tc|ln_fwd|ln_bwd|ln_nchw|scale_cast|colsum|patch_unfold|rel_bias" -c 45 \
    -o gpurun_out/${R}_fwd_stage0 python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_b.log 2>&1
# backward, stage 0 (last kernels of the step)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tc|attn_tc|ln_fwd|ln_bwd|ln_nchw|scale_cast|colsum|patch_unfold|rel_bias" --launch-skip 225 -c 60 \
    -o gpurun_out/${R}_bwd_stage0 python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_c.log 2>&1
# stage-2 block (tensor-bound GEMMs): forward kernels 60..75
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tc|attn_tc|ln_fwd|ln_bwd|ln_nchw|scale_cast|colsum|patch_unfold|rel_bias" --launch-skip 60 -c 16 \
    -o gpurun_out/${R}_fwd_stage2 python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/${R}_*
