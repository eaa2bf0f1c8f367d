#!/bin/bash
# Runs the GPU kernel tests group by group in separate processes (a trapped kernel poisons only its own
# process) with hard timeouts; logs go to gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, timeout, pytest args...
  name=$1; to=$2; shift 2
  timeout $to python -m pytest "$@" -q -rf --tb=short --timeout 120 -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "== $name exit=$? $(tail -n 1 gpurun_out/$name.log)"
}
for g in "$@"; do
  case $g in
    index) run index 300 tests/test_gpu_kernels.py -m gpu -k "index or roundtrip or rel_bias" ;;
    ln) run ln 300 tests/test_gpu_kernels.py -m gpu -k "layernorm or scale_cast or ln_nchw or patch_unfold" ;;
    gemm32) run gemm32 300 tests/test_gpu_kernels.py -m gpu -k "gemm and f32 and not bf16_matches and not tile_modes" ;;
    gemm16) run gemm16 300 tests/test_gpu_kernels.py -m gpu -k "gemm and bf16 and not tile_modes" ;;
    tiles) run tiles 300 tests/test_gpu_kernels.py -m gpu -k "tile_modes" ;;
    attn32) run attn32 300 tests/test_gpu_kernels.py -m gpu -k "window_attention_core and f32" ;;
    attn16) run attn16 300 tests/test_gpu_kernels.py -m gpu -k "window_attention_core and bf16" ;;
    attnfull) run attnfull 300 tests/test_gpu_kernels.py -m gpu -k "full_size_tcgen05" ;;
    optim) run optim 300 tests/test_optim.py -m gpu ;;
    model) run model 600 tests/test_gpu_model.py -m gpu ;;
    *) echo "unknown group $g" ;;
  esac
done
