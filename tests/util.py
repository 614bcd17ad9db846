"""Helpers shared by the GPU parity tests: device buffers via torch, calls via the C ABI."""
import ctypes as C

import numpy as np
import torch

from runbooks_b200 import _lib


def dev(x, dtype=torch.bfloat16):
    return torch.as_tensor(x).to(device="cuda", dtype=dtype).contiguous()


def call(engine, name, *args):
    torch.cuda.synchronize()
    fn = getattr(engine._lib, name)
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else a for a in args]
    st = fn(engine.handle, *conv)
    if st != 0:
        raise _lib.B200WError(st, (engine._lib.b200w_last_error(engine.handle) or b"").decode())


def rel_err(a, b):
    """||a-b||_F / ||b||_F in fp64."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> raw bf16 bit pattern (uint16), round-to-nearest-even."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) >> 16).astype(np.uint16)
