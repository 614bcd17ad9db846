"""Python host over the C ABI (include/b200w.h): one `Engine` = one b200w context on one GPU.

This is the layer the worker (runbooks_b200/worker.py — the container-contract entry point) and
bench.py drive. It holds no arithmetic; every number comes out of libb200w.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, asdict
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import Arch as _CArch, HParams as _CHParams, B200WError


@dataclass
class LlamaArch:
    """Subset of the HF LlamaConfig the kernels need (config.json keys in comments)."""
    vocab_size: int
    hidden_size: int
    intermediate_size: int
    num_layers: int            # num_hidden_layers
    num_heads: int             # num_attention_heads
    num_kv_heads: int          # num_key_value_heads
    head_dim: int = 128
    max_seq_len: int = 4096    # sequences are packed to this length
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0

    @classmethod
    def llama2_7b(cls, seq_len: int = 4096) -> "LlamaArch":
        return cls(32000, 4096, 11008, 32, 32, 32, 128, seq_len, 1e-5, 10000.0)

    @classmethod
    def from_hf_config(cls, cfg: dict, seq_len: Optional[int] = None) -> "LlamaArch":
        if cfg.get("model_type", "llama") != "llama":
            raise ValueError(f"unsupported model_type {cfg.get('model_type')!r} (llama only)")
        heads = cfg["num_attention_heads"]
        rope = cfg.get("rope_parameters") or {}
        return cls(
            vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"],
            intermediate_size=cfg["intermediate_size"], num_layers=cfg["num_hidden_layers"],
            num_heads=heads, num_kv_heads=cfg.get("num_key_value_heads") or heads,
            head_dim=cfg.get("head_dim") or cfg["hidden_size"] // heads,
            max_seq_len=seq_len or cfg.get("max_position_embeddings", 4096),
            rms_norm_eps=cfg.get("rms_norm_eps", 1e-6),
            rope_theta=float(cfg.get("rope_theta") or rope.get("rope_theta") or 10000.0),
        )

    def to_hf_config(self) -> dict:
        return {
            "architectures": ["LlamaForCausalLM"], "model_type": "llama",
            "vocab_size": self.vocab_size, "hidden_size": self.hidden_size,
            "intermediate_size": self.intermediate_size, "num_hidden_layers": self.num_layers,
            "num_attention_heads": self.num_heads, "num_key_value_heads": self.num_kv_heads,
            "head_dim": self.head_dim, "max_position_embeddings": self.max_seq_len,
            "rms_norm_eps": self.rms_norm_eps, "rope_theta": self.rope_theta,
            "hidden_act": "silu", "tie_word_embeddings": False, "attention_bias": False,
            "mlp_bias": False, "torch_dtype": "bfloat16",
        }


def _as_i32(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    """Owns a b200w context. Raises B200WError on any failure (no CPU fallback exists)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = _lib.c_ctx()
        st = self._lib.b200w_create(device, C.byref(h))
        if st != 0:
            raise B200WError(st, (self._lib.b200w_last_error(None) or b"").decode())
        self._h = h
        self.arch: Optional[LlamaArch] = None
        self.micro_batch = 0

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, st: int):
        if st != 0:
            raise B200WError(st, (self._lib.b200w_last_error(self._h) or b"").decode())

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200w_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self._lib.b200w_sync(self._h))

    # -- model --------------------------------------------------------------------------------
    def init_model(self, arch: LlamaArch, micro_batch: int = 1, training: bool = True,
                   max_grad_norm: float = 1.0, weight_decay: float = 0.0,
                   betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8):
        ca = _CArch(**asdict(arch))
        hp = _CHParams()
        self._lib.b200w_default_hparams(C.byref(hp))
        hp.max_grad_norm, hp.weight_decay = max_grad_norm, weight_decay
        hp.beta1, hp.beta2, hp.eps = betas[0], betas[1], eps
        self._check(self._lib.b200w_model_init(self._h, C.byref(ca), C.byref(hp), micro_batch,
                                               1 if training else 0))
        self.arch, self.micro_batch = arch, micro_batch

    def params(self) -> Iterable[Tuple[str, Tuple[int, ...]]]:
        n = C.c_int64()
        self._check(self._lib.b200w_param_count(self._h, C.byref(n), None))
        buf = C.create_string_buffer(256)
        r, c = C.c_int64(), C.c_int64()
        for i in range(n.value):
            self._check(self._lib.b200w_param_info(self._h, i, buf, 256, C.byref(r), C.byref(c)))
            name = buf.value.decode()
            yield name, ((c.value,) if name.endswith("norm.weight") else (r.value, c.value))

    def load_tensor(self, name: str, arr: np.ndarray):
        """arr: float32, or uint16 holding raw bf16 bits."""
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.uint16:
            dt = _lib.BF16
        else:
            arr = arr.astype(np.float32, copy=False)
            dt = _lib.F32
        self._check(self._lib.b200w_load_tensor(self._h, name.encode(), arr.ctypes.data, dt, arr.size))

    def load_state_dict(self, sd: Dict[str, np.ndarray]):
        names = {n for n, _ in self.params()}
        missing = names - set(sd)
        if missing:
            raise KeyError(f"state dict lacks {sorted(missing)[:4]}... ({len(missing)} tensors)")
        for n in names:
            self.load_tensor(n, sd[n])

    def read_tensor(self, name: str, shape, bf16_bits: bool = False) -> np.ndarray:
        out = np.empty(shape, dtype=np.uint16 if bf16_bits else np.float32)
        self._check(self._lib.b200w_read_tensor(self._h, name.encode(), out.ctypes.data,
                                                _lib.BF16 if bf16_bits else _lib.F32, out.size))
        return out

    def read_state(self, name: str, shape, kind: str) -> np.ndarray:
        out = np.empty(shape, dtype=np.float32)
        k = {"master": 0, "grad": 1, "m": 2, "v": 3}[kind]
        self._check(self._lib.b200w_read_state(self._h, name.encode(), k, out.ctypes.data, out.size))
        return out

    def state_dict(self, bf16_bits: bool = False) -> Dict[str, np.ndarray]:
        return {n: self.read_tensor(n, s, bf16_bits) for n, s in self.params()}

    def init_random(self, seed: int = 0, std: float = 0.02):
        self._check(self._lib.b200w_init_random(self._h, seed, std))

    # -- data parallel ------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        st = self._lib.b200w_comm_unique_id(buf)
        if st != 0:
            raise B200WError(st, (self._lib.b200w_last_error(None) or b"").decode())
        return buf.raw

    def comm_init(self, rank: int, nranks: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._check(self._lib.b200w_comm_init(self._h, rank, nranks, buf))

    # -- the hot path -------------------------------------------------------------------------
    def train_step(self, ids, labels, lr: float = 5e-5) -> Tuple[float, float]:
        """ids/labels: [n_seqs, seq_len] host integer arrays (labels unshifted, -100 ignored).
        Runs fwd + loss + bwd + all-reduce + clip + AdamW; returns (loss, grad_norm)."""
        ids, labels = _as_i32(ids), _as_i32(labels)
        assert ids.shape == labels.shape and ids.shape[1] == self.arch.max_seq_len
        loss, gn = C.c_float(), C.c_float()
        self._check(self._lib.b200w_train_step(self._h, ids.ctypes.data, labels.ctypes.data,
                                               ids.shape[0], lr, C.byref(loss), C.byref(gn)))
        return loss.value, gn.value

    def train_step_resident(self, ids_dev_ptr: int, labels_dev_ptr: int, n_seqs: int, n_valid: int,
                            lr: float = 5e-5):
        """The same step on a batch already in HBM (raw device addresses); no host sync."""
        self._check(self._lib.b200w_train_step_resident(self._h, ids_dev_ptr, labels_dev_ptr, n_seqs,
                                                        n_valid, lr))

    def read_scalars(self) -> Tuple[float, float]:
        loss, gn = C.c_float(), C.c_float()
        self._check(self._lib.b200w_read_scalars(self._h, C.byref(loss), C.byref(gn)))
        return loss.value, gn.value

    def timer_start(self):
        self._check(self._lib.b200w_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._check(self._lib.b200w_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def profile_gemm(self, enable: bool):
        self._check(self._lib.b200w_profile_gemm(self._h, 1 if enable else 0))

    def profile_read(self) -> Tuple[float, float, int]:
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self._lib.b200w_profile_read(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, n.value

    def forward_backward(self, ids, labels) -> float:
        ids, labels = _as_i32(ids), _as_i32(labels)
        loss = C.c_float()
        self._check(self._lib.b200w_forward_backward(self._h, ids.ctypes.data, labels.ctypes.data,
                                                     ids.shape[0], C.byref(loss)))
        return loss.value

    def forward(self, ids, labels=None, want_logits: bool = True):
        """Returns (logits [T,V] float32 or None, nll [T] or None, loss or None)."""
        ids = _as_i32(ids)
        T = ids.size
        logits = np.empty((T, self.arch.vocab_size), dtype=np.float32) if want_logits else None
        lab = _as_i32(labels) if labels is not None else None
        nll = np.empty(T, dtype=np.float32) if lab is not None else None
        loss = C.c_float()
        self._check(self._lib.b200w_forward(
            self._h, ids.ctypes.data, lab.ctypes.data if lab is not None else None, ids.shape[0],
            logits.ctypes.data if want_logits else None, nll.ctypes.data if nll is not None else None,
            C.byref(loss)))
        return logits, nll, (loss.value if lab is not None else None)

    def launch_count(self) -> int:
        return int(self._lib.b200w_launch_count(self._h))

    def device_bytes(self) -> int:
        return int(self._lib.b200w_device_bytes(self._h))
