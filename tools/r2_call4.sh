#!/bin/bash
# 2 GPUs: multi-rank parity inside pytest, N=2 bench (overlapped bf16 per-matrix all-reduce), then the decode changes
set -u; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== multi-GPU pytest"; timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -12
echo "== bench N=2"; timeout 300 $T --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -c 700 gpurun_out/r2_bench_n2.json; grep -E "b200w:|bench.py:|Error" gpurun_out/bench_n2.err | tail -3
echo "== bench N=2, sharded optimiser state"; B200W_SHARD_STATE=1 timeout 300 $T --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_shard.json 2> gpurun_out/bench_n2s.err; echo "rc=$?"; tail -c 400 gpurun_out/r2_bench_n2_shard.json; grep -E "b200w:|bench.py:|Error" gpurun_out/bench_n2s.err | tail -3
echo "== infer tests after the cluster split-K + fused merge"; timeout 400 python -m pytest tests/test_infer.py tests/test_infer_round2.py tests/test_onchip_state.py tests/test_gemm.py tests/test_opt.py tests/test_server_round2.py tests/test_worker_opt.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "opt grad|passed|failed|FAILED|Error|rel_err|greedy|identical" | tail -25
echo "== decode"; timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v2.json 2> gpurun_out/decode.err; cat gpurun_out/r2_decode_v2.json; tail -2 gpurun_out/decode.err
