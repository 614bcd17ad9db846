"""CPU oracle for the Server decode path (falcon-7b layout) — TEST INFRASTRUCTURE ONLY.

Restates HF transformers 5.5.0 models/falcon/modeling_falcon.py for multi_query=True,
parallel_attn=True, new_decoder_architecture=False, bias=False, alibi=False (what
examples/falcon-7b-instruct/server.yaml serves):
  FalconDecoderLayer.forward (:580-650)  ln = input_layernorm(h); h + attn(ln) + mlp(ln)
  FalconAttention._split_heads (:259-281) fused qkv = [H q heads | 1 k head | 1 v head]
  FalconRotaryEmbedding / apply_rotary_pos_emb : rotate_half, theta 10000
  FalconMLP (:528-543)                   dense_4h_to_h(gelu(dense_h_to_4h(x))), exact (erf) GeLU
  lm_head tied to word_embeddings
Pinned against FalconForCausalLM outputs in tests/golden/falcon_tiny.npz, and the fine-tune step (train_step:
HF Trainer loss normaliser, clip, AdamW with the Trainer's decay groups) against two real optimiser steps in
tests/golden/falcon_tiny_train.npz (oracle/make_golden.py run_falcon / run_falcon_train).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .llama_oracle import (adamw_update, apply_rope, bf16_round, causal_attention, causal_lm_loss, clip_grad_norm,
                           rope_cos_sin, trainer_num_items)
from .llama_oracle import decays as _decays_by_name


@dataclass
class FalconArch:
    vocab_size: int
    hidden_size: int
    num_layers: int
    num_heads: int
    head_dim: int = 64
    ffn_hidden_size: int = 0
    layer_norm_epsilon: float = 1e-5
    rope_theta: float = 10000.0

    @property
    def ffn(self):
        return self.ffn_hidden_size or 4 * self.hidden_size


def param_shapes(a: FalconArch):
    d, dh, H = a.hidden_size, a.head_dim, a.num_heads
    s = {"transformer.word_embeddings.weight": (a.vocab_size, d)}
    for l in range(a.num_layers):
        p = f"transformer.h.{l}."
        s[p + "input_layernorm.weight"] = (d,)
        s[p + "input_layernorm.bias"] = (d,)
        s[p + "self_attention.query_key_value.weight"] = ((H + 2) * dh, d)
        s[p + "self_attention.dense.weight"] = (d, H * dh)
        s[p + "mlp.dense_h_to_4h.weight"] = (a.ffn, d)
        s[p + "mlp.dense_4h_to_h.weight"] = (d, a.ffn)
    s["transformer.ln_f.weight"] = (d,)
    s["transformer.ln_f.bias"] = (d,)
    return s


def seeded_params(a: FalconArch, seed: int = 0, std: float = 0.12) -> Dict[str, np.ndarray]:
    """Small embeddings, larger layer matrices: with tied embeddings a random model otherwise just
    repeats its last input token, which would make greedy-parity tests vacuous."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in param_shapes(a).items():
        if name.endswith("layernorm.weight") or name.endswith("ln_f.weight"):
            w = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            w = 0.05 * rng.standard_normal(shape)
        elif "word_embeddings" in name:
            w = 0.02 * rng.standard_normal(shape)
        else:
            w = std * rng.standard_normal(shape)
        out[name] = bf16_round(w.astype(np.float32))
    return out


def forward(params: Dict[str, torch.Tensor], ids: torch.Tensor, a: FalconArch) -> torch.Tensor:
    """ids [B,S] -> logits [B,S,V] (fp32, full causal forward, no cache)."""
    B, S = ids.shape
    H, dh, d = a.num_heads, a.head_dim, a.hidden_size
    cos, sin = rope_cos_sin(S, dh, a.rope_theta)
    h = params["transformer.word_embeddings.weight"][ids]
    for l in range(a.num_layers):
        p = f"transformer.h.{l}."
        ln = F.layer_norm(h, (d,), params[p + "input_layernorm.weight"], params[p + "input_layernorm.bias"],
                          a.layer_norm_epsilon)
        qkv = F.linear(ln, params[p + "self_attention.query_key_value.weight"]).view(B, S, H + 2, dh)
        q = qkv[:, :, :H].transpose(1, 2)
        k = qkv[:, :, H:H + 1].transpose(1, 2)
        v = qkv[:, :, H + 1:].transpose(1, 2)
        q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
        att = causal_attention(q, k, v).transpose(1, 2).reshape(B, S, H * dh)
        att = F.linear(att, params[p + "self_attention.dense.weight"])
        mlp = F.linear(F.gelu(F.linear(ln, params[p + "mlp.dense_h_to_4h.weight"])),
                       params[p + "mlp.dense_4h_to_h.weight"])
        h = h + att + mlp
    h = F.layer_norm(h, (d,), params["transformer.ln_f.weight"], params["transformer.ln_f.bias"],
                     a.layer_norm_epsilon)
    return F.linear(h, params["transformer.word_embeddings.weight"])


def decays(name: str) -> bool:
    """Trainer.get_decay_parameter_names (trainer.py:1280-1290) excludes nn.LayerNorm parameters by module TYPE
    as well as by name: Falcon's final `ln_f` is an nn.LayerNorm whose name matches none of the patterns (the
    golden's no_decay list holds it; found by the two-step pin, 8.6e-4 on ln_f.weight)."""
    return _decays_by_name(name) and not name.startswith("transformer.ln_f.")


def train_step(params_np, ids, labels, a: FalconArch, lr=5e-5, max_grad_norm=1.0, state=None, step=1,
               weight_decay=0.0):
    """fwd, HF causal-LM loss, bwd (the tied table receives lookup + head gradients through autograd; Falcon's
    nn.Embedding has no padding_idx, modeling_falcon.py:680), global-norm clip, AdamW. Same return layout as
    llama_oracle.train_step."""
    P = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in params_np.items()}
    logits = forward(P, torch.as_tensor(ids, dtype=torch.int64), a)
    lab = torch.as_tensor(labels, dtype=torch.int64)
    loss, nll = causal_lm_loss(logits, lab, trainer_num_items(lab))
    loss.backward()
    grads = {k: p.grad.detach() for k, p in P.items()}
    gnorm, coef = clip_grad_norm(grads, max_grad_norm)
    new_p, new_m, new_v = {}, {}, {}
    for k, p in P.items():
        m0 = torch.zeros_like(p) if state is None else torch.as_tensor(state["m"][k])
        v0 = torch.zeros_like(p) if state is None else torch.as_tensor(state["v"][k])
        pn, mn, vn = adamw_update(p.detach(), grads[k] * coef, m0, v0, step, lr,
                                  wd=weight_decay if decays(k) else 0.0)
        new_p[k], new_m[k], new_v[k] = pn.numpy(), mn.numpy(), vn.numpy()
    return dict(loss=float(loss.detach()), gnorm=gnorm, logits=logits.detach().numpy(), nll=nll.detach().numpy(),
                grads={k: g.numpy() for k, g in grads.items()}, params=new_p, m=new_m, v=new_v)


def greedy(params_np, prompt_ids, n_new: int, a: FalconArch):
    """Greedy continuation by full recomputation (tiny models only). Returns (tokens, per-step
    logits of the last position, top-2 margins)."""
    params = {k: torch.tensor(v) for k, v in params_np.items()}
    ids = torch.tensor(np.asarray(prompt_ids, dtype=np.int64))[None]
    out, logs = [], []
    with torch.no_grad():
        for _ in range(n_new):
            lg = forward(params, ids, a)[0, -1]
            t = int(lg.argmax())
            out.append(t)
            logs.append(lg.numpy())
            ids = torch.cat([ids, torch.tensor([[t]])], dim=1)
    return out, np.stack(logs)
