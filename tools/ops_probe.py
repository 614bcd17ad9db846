"""Times the bandwidth-bound kernels of the fine-tune step in isolation at Llama-2-7B micro-batch size (T x d =
8192 x 4096) with CUDA events on the launching stream (the engine's hooks run on the context's stream and
sync). Prints microseconds and achieved GB/s against the algorithmic bytes. Use for same-box A/B of a kernel
change:  python tools/ops_probe.py > gpurun_out/ops_probe.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from runbooks_b200.engine import Engine  # noqa: E402
from util import call  # noqa: E402

e = Engine(0)
T, d = 8192, 4096
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, d, device="cuda", generator=g).bfloat16()
dy = torch.randn(T, d, device="cuda", generator=g).bfloat16()
dres = torch.randn(T, d, device="cuda", generator=g).bfloat16()
w = torch.randn(d, device="cuda", generator=g).bfloat16()
y = torch.empty_like(x)
dx = torch.empty_like(x)
rstd = torch.empty(T, device="cuda", dtype=torch.float32)
dw = torch.zeros(d, device="cuda", dtype=torch.float32)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


def timed(fn, reps=20):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # evict L2 between repetitions
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()                               # the hook syncs the context stream before returning
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6


out = {}
call(e, "b200w_op_rmsnorm_fwd", x, w, y, rstd, T, d, 1e-5)
for name, fn, nbytes in (
        ("rmsnorm_fwd", lambda: call(e, "b200w_op_rmsnorm_fwd", x, w, y, rstd, T, d, 1e-5), 2 * T * d * 2),
        ("rmsnorm_bwd", lambda: call(e, "b200w_op_rmsnorm_bwd", dy, x, w, rstd, dres, dx, dw, T, d), 4 * T * d * 2),
        ("rmsnorm_bwd_noresid", lambda: call(e, "b200w_op_rmsnorm_bwd", dy, x, w, rstd, None, dx, dw, T, d), 3 * T * d * 2)):
    us = timed(fn)
    out[name] = {"us_incl_host_call": round(us, 1), "GB_per_s": round(nbytes / us / 1e3, 1), "bytes": nbytes}
print(json.dumps(out, indent=1))
