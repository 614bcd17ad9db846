"""Launches one kernel a few times at Llama-2-7B size, for `ncu --set full -k regex:... -s 2 -c 1`.
    python tools/one_kernel.py gemm_fwd|gemm_dgrad|gemm_wgrad|gemm_wgrad_acc|attn_fwd|attn_bwd
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from runbooks_b200.engine import Engine  # noqa: E402
from util import call  # noqa: E402

which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e = Engine(0)
T, d, f = int(os.environ.get("TOKENS", "4096")), 4096, 11008   # TOKENS=8192: the micro-batch-2 launches
if which.startswith("gemm"):
    M, N, K, a_mn, b_mn, f32 = {
        "gemm_fwd": (T, 2 * f, d, 0, 0, 0), "gemm_dgrad": (T, d, 2 * f, 0, 1, 0),
        "gemm_wgrad": (2 * f, d, T, 1, 1, 1), "gemm_wgrad_acc": (2 * f, d, T, 1, 1, 1),
        "gemm_wgrad_bf16": (2 * f, d, T, 1, 1, 0), "gemm_amn_only": (2 * f, d, T, 1, 0, 0),
    }[which]
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").bfloat16()
    D = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    C = D if which.endswith("acc") else None
    for _ in range(reps):
        call(e, "b200w_op_gemm", A, a_mn, A.shape[1], B, b_mn, B.shape[1], D, C, f32, N, M, N, K, 0)
else:
    B_, S, H = 1, 4096, 32
    qkv = torch.randn(B_ * S, 3 * H * 128, device="cuda").bfloat16()
    o = torch.empty(B_ * S, H * 128, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(H, B_ * S, device="cuda", dtype=torch.float32)
    for _ in range(reps):
        call(e, "b200w_op_attention_fwd", qkv, 3 * H * 128, H * 128, 2 * H * 128, o, H * 128, lse, B_, S, H,
             H, 128 ** -0.5)
    if which == "attn_bwd":
        do = torch.randn_like(o)
        delta = torch.empty_like(lse)
        dqkv = torch.empty_like(qkv)
        for _ in range(reps):
            call(e, "b200w_op_attention_bwd", qkv, 3 * H * 128, H * 128, 2 * H * 128, o, do, H * 128, lse,
                 delta, dqkv, B_, S, H, H, 128 ** -0.5)
print("done", which)
