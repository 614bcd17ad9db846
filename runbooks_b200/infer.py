"""Host side of the Server decode path: `InferEngine` wraps the b200w_infer_* C ABI; `Generator`
turns it into a continuously-batched greedy generator (one cache slot per request, one token per
slot per step — prompt ingestion and generation are the same step)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import B200WError, InferArch as _CArch
from .engine import Engine

FAMILY = {"llama": 0, "falcon": 1}


@dataclass
class ServeArch:
    family: str
    vocab_size: int
    hidden_size: int
    intermediate_size: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    max_ctx: int = 2048
    norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    tie_embeddings: bool = False

    @classmethod
    def falcon_7b(cls, max_ctx: int = 2048) -> "ServeArch":
        return cls("falcon", 65024, 4544, 18176, 32, 71, 1, 64, max_ctx, 1e-5, 10000.0, True)

    @classmethod
    def from_hf_config(cls, cfg: dict, max_ctx: Optional[int] = None) -> "ServeArch":
        mt = cfg.get("model_type")
        if mt == "falcon":
            if cfg.get("alibi") or cfg.get("bias") or cfg.get("new_decoder_architecture") or not cfg.get(
                    "parallel_attn", True):
                raise ValueError("only the falcon-7b layout (rotary, bias-free, parallel_attn, 1 layernorm) is supported")
            heads = cfg["num_attention_heads"]
            d = cfg["hidden_size"]
            kv = 1 if cfg.get("multi_query", True) else heads
            return cls("falcon", cfg["vocab_size"], d, cfg.get("ffn_hidden_size") or 4 * d,
                       cfg["num_hidden_layers"], heads, kv, d // heads,
                       max_ctx or min(2048, cfg.get("max_position_embeddings", 2048)),
                       cfg.get("layer_norm_epsilon", 1e-5), float(cfg.get("rope_theta", 10000.0)),
                       bool(cfg.get("tie_word_embeddings", True)))
        if mt == "llama":
            heads = cfg["num_attention_heads"]
            rope = cfg.get("rope_parameters") or {}
            return cls("llama", cfg["vocab_size"], cfg["hidden_size"], cfg["intermediate_size"],
                       cfg["num_hidden_layers"], heads, cfg.get("num_key_value_heads") or heads,
                       cfg.get("head_dim") or cfg["hidden_size"] // heads,
                       max_ctx or min(4096, cfg.get("max_position_embeddings", 4096)),
                       cfg.get("rms_norm_eps", 1e-6), float(cfg.get("rope_theta") or rope.get("rope_theta") or 1e4),
                       bool(cfg.get("tie_word_embeddings", False)))
        raise ValueError(f"unsupported model_type {mt!r} (falcon, llama)")


class InferEngine(Engine):
    def init_infer(self, arch: ServeArch, max_batch: int = 32):
        ca = _CArch(FAMILY[arch.family], arch.vocab_size, arch.hidden_size, arch.intermediate_size,
                    arch.num_layers, arch.num_heads, arch.num_kv_heads, arch.head_dim, arch.max_ctx,
                    arch.norm_eps, arch.rope_theta, 1 if arch.tie_embeddings else 0)
        self._check(self._lib.b200w_infer_init(self._h, C.byref(ca), max_batch))
        self.serve_arch, self.max_batch = arch, max_batch

    def infer_params(self) -> Iterable[Tuple[str, Tuple[int, ...]]]:
        n = C.c_int64()
        self._check(self._lib.b200w_infer_param_count(self._h, C.byref(n), None))
        buf = C.create_string_buffer(256)
        r, c = C.c_int64(), C.c_int64()
        for i in range(n.value):
            self._check(self._lib.b200w_infer_param_info(self._h, i, buf, 256, C.byref(r), C.byref(c)))
            name = buf.value.decode()
            one_d = r.value == 1 and ("norm" in name or "ln_f" in name)
            yield name, ((c.value,) if one_d else (r.value, c.value))

    def infer_load_tensor(self, name: str, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        dt = _lib.BF16 if arr.dtype == np.uint16 else _lib.F32
        if dt == _lib.F32:
            arr = arr.astype(np.float32, copy=False)
        self._check(self._lib.b200w_infer_load_tensor(self._h, name.encode(), arr.ctypes.data, dt, arr.size))

    def infer_load_state_dict(self, sd: Dict[str, np.ndarray]):
        for name, _ in self.infer_params():
            if name not in sd:
                raise KeyError(f"state dict lacks {name}")
            self.infer_load_tensor(name, sd[name])

    def infer_init_random(self, seed: int = 0, std: float = 0.02):
        self._check(self._lib.b200w_infer_init_random(self._h, seed, std))

    def step(self, tokens, positions, slots, want_logits: bool = False):
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        pos = np.ascontiguousarray(positions, dtype=np.int32)
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        n = tok.size
        nxt = np.empty(n, dtype=np.int32)
        logits = np.empty((n, self.serve_arch.vocab_size), dtype=np.float32) if want_logits else None
        self._check(self._lib.b200w_infer_step(self._h, tok.ctypes.data, pos.ctypes.data, sl.ctypes.data, n,
                                               nxt.ctypes.data, logits.ctypes.data if want_logits else None))
        return nxt, logits


@dataclass
class _Req:
    prompt: List[int]
    max_tokens: int
    out: List[int]
    fed: int = 0          # tokens fed so far (prompt + generated)
    slot: int = -1
    done: bool = False


class Generator:
    """Greedy continuous batching over the engine's cache slots."""

    def __init__(self, engine: InferEngine, eos_id: Optional[int] = None):
        self.e, self.eos = engine, eos_id
        self.free = list(range(engine.max_batch))
        self.active: List[_Req] = []

    def add(self, prompt: List[int], max_tokens: int) -> _Req:
        if not prompt:
            raise ValueError("empty prompt")
        if len(prompt) + max_tokens > self.e.serve_arch.max_ctx:
            raise ValueError("prompt + max_tokens exceeds the KV cache length")
        if not self.free:
            raise RuntimeError("no free cache slot")
        r = _Req(list(prompt), max_tokens, [], 0, self.free.pop())
        self.active.append(r)
        return r

    def step(self):
        """One engine step for every active request."""
        if not self.active:
            return
        toks = [(r.prompt[r.fed] if r.fed < len(r.prompt) else r.out[-1]) for r in self.active]
        nxt, _ = self.e.step(toks, [r.fed for r in self.active], [r.slot for r in self.active])
        for r, t in zip(self.active, nxt):
            r.fed += 1
            if r.fed >= len(r.prompt):        # the token just fed was the last known one
                r.out.append(int(t))
                if len(r.out) >= r.max_tokens or (self.eos is not None and int(t) == self.eos):
                    r.done = True
        for r in [r for r in self.active if r.done]:
            self.active.remove(r)
            self.free.append(r.slot)

    def generate(self, prompts: List[List[int]], max_tokens: int) -> List[List[int]]:
        reqs = [self.add(p, max_tokens) for p in prompts]
        while self.active:
            self.step()
        return [r.out for r in reqs]
