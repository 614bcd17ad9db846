"""Quick device-timed throughput probe of the two tensor-core kernels at Llama-2-7B sizes.
Development aid (bench.py is the contract); prints one line per kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from runbooks_b200.engine import Engine  # noqa: E402
from util import call  # noqa: E402
from bench import ClockSampler  # noqa: E402  (numbers from different boxes are only comparable with clocks)


def timeit(fn, iters=10, warm=3, e=None):
    """Median DEVICE time of fn(): CUDA events on the library's own stream (b200w_timer_*), L2
    flushed between iterations."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        ENGINE.timer_start()
        fn()
        ts.append(ENGINE.timer_stop() * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


ENGINE = None


def main():
    global ENGINE
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", choices=("all", "gemm", "attn"), default="all")
    ap.add_argument("--out", default="gpurun_out/perf_probe.json")
    args = ap.parse_args()
    e = ENGINE = Engine(0)
    out = {}
    clocks = ClockSampler(0)
    clocks.start()
    T, d, f, V = 4096, 4096, 11008, 32000
    shapes = {
        "fwd_qkv": (T, 3 * d, d, 0, 0), "fwd_gateup": (T, 2 * f, d, 0, 0), "fwd_down": (T, d, f, 0, 0),
        "fwd_lmhead": (T, V, d, 0, 0), "dgrad_gateup": (T, d, 2 * f, 0, 1), "dgrad_down": (T, f, d, 0, 1),
        "wgrad_gateup": (2 * f, d, T, 1, 1), "wgrad_gateup_acc": (2 * f, d, T, 1, 1), "wgrad_down": (d, f, T, 1, 1),
        "wgrad_down_acc": (d, f, T, 1, 1), "wgrad_lmhead_acc": (V, d, T, 1, 1), "dgrad_qkv": (T, d, 3 * d, 0, 1),
        "square_8k": (8192, 8192, 8192, 0, 0),
    }
    if args.only == "attn":
        shapes = {}
    for name, (M, N, K, a_mn, b_mn) in shapes.items():
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
        wg = name.startswith("wgrad")
        D = torch.empty(M, N, device="cuda", dtype=torch.float32 if wg else torch.bfloat16)
        for bn in (512,):
            Cacc = D if name.endswith("_acc") else None
            t = timeit(lambda: call(e, "b200w_op_gemm", A, a_mn, A.shape[1], B, b_mn, B.shape[1], D, Cacc,
                                    1 if wg else 0, N, M, N, K, bn))
            tf = 2.0 * M * N * K / t / 1e12
            out[f"gemm_{name}_bn{bn}"] = dict(ms=t * 1e3, tflops=tf)
            print(f"gemm {name:14s} bn{bn} M{M} N{N} K{K}: {t * 1e3:8.3f} ms  {tf:7.1f} TF/s", flush=True)
        del A, B, D
    if args.only == "gemm":
        return finish(out, clocks, args.out)
    # attention, one 4096-token sequence, 32 heads
    B_, S, H = 1, 4096, 32
    qkv = torch.randn(B_ * S, 3 * H * 128, device="cuda").bfloat16()
    o = torch.empty(B_ * S, H * 128, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(H, B_ * S, device="cuda", dtype=torch.float32)
    t = timeit(lambda: call(e, "b200w_op_attention_fwd", qkv, 3 * H * 128, H * 128, 2 * H * 128, o, H * 128,
                            lse, B_, S, H, H, 128 ** -0.5))
    fl = 4.0 * B_ * H * S * S * 128 / 2
    out["attn_fwd"] = dict(ms=t * 1e3, tflops=fl / t / 1e12)
    print(f"attention fwd S{S} H{H}: {t * 1e3:.3f} ms  {fl / t / 1e12:.1f} TF/s (causal flops)", flush=True)
    do = torch.randn_like(o)
    delta = torch.empty_like(lse)
    dqkv = torch.empty_like(qkv)
    t = timeit(lambda: call(e, "b200w_op_attention_bwd", qkv, 3 * H * 128, H * 128, 2 * H * 128, o, do,
                            H * 128, lse, delta, dqkv, B_, S, H, H, 128 ** -0.5))
    out["attn_bwd"] = dict(ms=t * 1e3, tflops=2.5 * fl / t / 1e12)
    print(f"attention bwd S{S} H{H}: {t * 1e3:.3f} ms  {2.5 * fl / t / 1e12:.1f} TF/s (causal flops)", flush=True)
    finish(out, clocks, args.out)


def finish(out, clocks, path):
    out["clocks"] = clocks.stop()
    print("clocks", out["clocks"], flush=True)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
