#!/bin/bash
# Launch list of ONE benchmark step (the same command bench.py times): per-launch duration + DRAM bytes, three metrics
# only so the kernel-replay passes stay cheap.  Outputs -> gpurun_out/<R>_step_launches.{csv,txt} + gemm_traffic.json.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${1:-r01}
python bench.py --ncu-step --warmup 3 > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/${R}_step_launches.csv python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_step.log 2>&1
echo "ncu exit=$?"
python tools/summarize_launches.py gpurun_out/${R}_step_launches.csv --json gpurun_out/${R}_gemm_traffic.json > gpurun_out/${R}_step_launches.txt
head -12 gpurun_out/${R}_step_launches.txt
