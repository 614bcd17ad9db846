#!/bin/bash
# ncu --set full (with source) of single kernels at Llama-2-7B size -> gpurun_out/k_*.ncu-rep
mkdir -p gpurun_out
run() { # name regex script-arg skip
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c 1 -f \
    -o "gpurun_out/k_$1" python tools/one_kernel.py "$3" 3 > "gpurun_out/ncu_$1.log" 2>&1
  tail -1 "gpurun_out/ncu_$1.log"
}
run attn_fwd attn_fwd_kernel attn_fwd 2
run attn_dkdv attn_bwd_dkdv attn_bwd 2
run attn_dq attn_bwd_dq attn_bwd 2
run gemm_wgrad gemm_bf16 gemm_wgrad 2
run gemm_wgrad_acc gemm_bf16 gemm_wgrad_acc 2
run gemm_wgrad_bf16 gemm_bf16 gemm_wgrad_bf16 2
run gemm_fwd gemm_bf16 gemm_fwd 2
ls -la gpurun_out/*.ncu-rep
