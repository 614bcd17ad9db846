#!/bin/bash
# ncu --set full of the dominant GEMM launches at the micro-batch-2 shape (M = 8192 tokens)
set -u; mkdir -p gpurun_out
run() { TOKENS=8192 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c 1 -f -o "gpurun_out/r2_k_$1_mb2" python tools/one_kernel.py "$3" 3 > "gpurun_out/ncu_$1.log" 2>&1; tail -1 "gpurun_out/ncu_$1.log"; }
run gemm_wgrad_acc gemm_bf16 gemm_wgrad_acc 2
run gemm_fwd gemm_bf16 gemm_fwd 2
run gemm_dgrad gemm_bf16 gemm_dgrad 2
ls -la gpurun_out/*_mb2.ncu-rep
