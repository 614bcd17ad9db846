#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c 1 -f -o "gpurun_out/k_$1" python tools/one_kernel.py "$3" 3 > "gpurun_out/ncu_$1.log" 2>&1; tail -1 "gpurun_out/ncu_$1.log"; }
run attn_fwd attn_fwd_kernel attn_fwd 2
run attn_dkdv attn_bwd_dkdv attn_bwd 2
run attn_dq attn_bwd_dq attn_bwd 2
