"""CPU oracle for the reference's config #1 (facebook/opt-125m fine-tune) — TEST INFRASTRUCTURE ONLY.

Nothing under runbooks_b200/ may import this file; tests, __graft_entry__.smoke() and bench.py's CPU
leg are the only users (tests/test_abi.py checks). SURVEY.md 8 row a15; selector in the reference:
examples/facebook-opt-125m/finetuned-model.yaml (the un-vendored trainer image wraps HF + torch).

Restates HF transformers 5.5.0 models/opt/modeling_opt.py for the opt-125m family
(do_layer_norm_before=True, word_embed_proj_dim == hidden_size, enable_bias=True, relu, tied head):
  OPTLearnedPositionalEmbedding (:45-70)  positions = cumsum(mask) * mask - 1, looked up at +2
                                          (the table has max_position_embeddings + 2 rows)
  OPTAttention.forward (:135-183)          q = q_proj(x) * head_dim**-0.5 BEFORE the attention call,
                                          which then runs with scaling 1.0; biases on q, k, v, out
  OPTDecoderLayer.forward (:203-249)       pre-LN: h + attn(LN1(h)); h + fc2(relu(fc1(LN2(h))))
  OPTDecoder (:286-312)                    embed_tokens + embed_positions, layers, final_layer_norm;
                                          embed_tokens = nn.Embedding(V, d, padding_idx=pad_token_id) (:289):
                                          the LOOKUP contributes no gradient to the pad row (the tied head
                                          still does) -- found by the golden, 7e-4 on that tensor's gradient
  OPTForCausalLM                            lm_head.weight tied to embed_tokens.weight, no bias
Loss, clip and AdamW are llama_oracle's restatements of loss_utils / clip_grad_norm_ / AdamW.
Pinned against the real OPTForCausalLM by tests/golden/opt_tiny.npz (oracle/make_golden.py run_opt):
two optimiser steps with -100 labels, logits, gradients, updated weights, greedy continuation.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from .llama_oracle import adamw_update, bf16_round, causal_lm_loss, clip_grad_norm, decays, trainer_num_items

POS_OFFSET = 2  # modeling_opt.py:53


@dataclass
class OptArch:
    vocab_size: int
    hidden_size: int
    ffn_dim: int
    num_layers: int
    num_heads: int
    max_position_embeddings: int = 2048
    pad_token_id: int = 1             # OPTConfig default; only its gradient row matters here
    layer_norm_eps: float = 1e-5      # nn.LayerNorm default: OPT passes no eps

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads


OPT_125M = OptArch(50272, 768, 3072, 12, 12, 2048)


def param_shapes(a: OptArch) -> Dict[str, tuple]:
    d, f = a.hidden_size, a.ffn_dim
    s = {"model.decoder.embed_tokens.weight": (a.vocab_size, d),
         "model.decoder.embed_positions.weight": (a.max_position_embeddings + POS_OFFSET, d)}
    for l in range(a.num_layers):
        p = f"model.decoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = (d, d)
            s[p + f"self_attn.{n}.bias"] = (d,)
        s[p + "self_attn_layer_norm.weight"] = (d,)
        s[p + "self_attn_layer_norm.bias"] = (d,)
        s[p + "fc1.weight"] = (f, d)
        s[p + "fc1.bias"] = (f,)
        s[p + "fc2.weight"] = (d, f)
        s[p + "fc2.bias"] = (d,)
        s[p + "final_layer_norm.weight"] = (d,)
        s[p + "final_layer_norm.bias"] = (d,)
    s["model.decoder.final_layer_norm.weight"] = (d,)
    s["model.decoder.final_layer_norm.bias"] = (d,)
    return s  # lm_head.weight is embed_tokens.weight


def seeded_params(a: OptArch, seed: int = 0, std: float = 0.08, bf16: bool = True) -> Dict[str, np.ndarray]:
    """Non-trivial biases and LayerNorm affine parameters (HF's init would make them 0 / 1 and hide
    a dropped bias); small embeddings so that a tied-head random model does not just echo its input."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in param_shapes(a).items():
        if k.endswith("layer_norm.weight"):
            v = 1.0 + 0.1 * rng.standard_normal(shp)
        elif k.endswith(".bias"):
            v = 0.05 * rng.standard_normal(shp)
        elif "embed_" in k:
            v = 0.03 * rng.standard_normal(shp)
        else:
            v = std * rng.standard_normal(shp)
        v = v.astype(np.float32)
        out[k] = bf16_round(v) if bf16 else v
    return out


def positions(attention_mask: torch.Tensor) -> torch.Tensor:
    """OPTLearnedPositionalEmbedding.forward (:64-70) without a KV cache."""
    pos = torch.cumsum(attention_mask, dim=1) * attention_mask - 1
    return pos.long() + POS_OFFSET


def attention(x: torch.Tensor, P: Dict[str, torch.Tensor], pre: str, a: OptArch) -> torch.Tensor:
    B, S, d = x.shape
    H, dh = a.num_heads, a.head_dim
    q = F.linear(x, P[pre + "q_proj.weight"], P[pre + "q_proj.bias"]) * dh ** -0.5     # :151
    k = F.linear(x, P[pre + "k_proj.weight"], P[pre + "k_proj.bias"])
    v = F.linear(x, P[pre + "v_proj.weight"], P[pre + "v_proj.bias"])
    q, k, v = (t.view(B, S, H, dh).transpose(1, 2) for t in (q, k, v))
    s = q @ k.transpose(-1, -2)                                                         # scaling = 1.0
    mask = torch.ones(S, S, dtype=torch.bool, device=x.device).tril()
    s = s.masked_fill(~mask, float("-inf"))
    o = torch.softmax(s.float(), dim=-1).to(v.dtype) @ v
    o = o.transpose(1, 2).reshape(B, S, d)
    return F.linear(o, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])


def forward(P: Dict[str, torch.Tensor], ids: torch.Tensor, a: OptArch) -> torch.Tensor:
    """ids [B, S] (no padding: attention_mask all ones) -> logits [B, S, V] fp32."""
    B, S = ids.shape
    eps = a.layer_norm_eps
    mask = torch.ones(B, S, dtype=torch.long, device=ids.device)
    h = (F.embedding(ids, P["model.decoder.embed_tokens.weight"], padding_idx=a.pad_token_id)
         + P["model.decoder.embed_positions.weight"][positions(mask)])
    for l in range(a.num_layers):
        p = f"model.decoder.layers.{l}."
        x = F.layer_norm(h, (a.hidden_size,), P[p + "self_attn_layer_norm.weight"], P[p + "self_attn_layer_norm.bias"], eps)
        h = h + attention(x, P, p + "self_attn.", a)
        x = F.layer_norm(h, (a.hidden_size,), P[p + "final_layer_norm.weight"], P[p + "final_layer_norm.bias"], eps)
        h = h + F.linear(F.relu(F.linear(x, P[p + "fc1.weight"], P[p + "fc1.bias"])), P[p + "fc2.weight"], P[p + "fc2.bias"])
    h = F.layer_norm(h, (a.hidden_size,), P["model.decoder.final_layer_norm.weight"],
                     P["model.decoder.final_layer_norm.bias"], eps)
    return F.linear(h, P["model.decoder.embed_tokens.weight"])                          # tied head, no bias


def train_step(params_np, ids, labels, a: OptArch, lr=5e-5, max_grad_norm=1.0, state=None, step=1,
               weight_decay=0.0):
    """fwd, HF causal-LM loss, bwd (the tied table receives embedding + head gradients through
    autograd), global-norm clip, AdamW. Same return layout as llama_oracle.train_step."""
    P = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in params_np.items()}
    logits = forward(P, torch.as_tensor(ids, dtype=torch.int64), a)
    lab = torch.as_tensor(labels, dtype=torch.int64)
    loss, nll = causal_lm_loss(logits, lab, trainer_num_items(lab))   # HF Trainer's normaliser
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


def greedy(params_np, prompts: np.ndarray, a: OptArch, max_new_tokens: int) -> np.ndarray:
    """model.generate(do_sample=False) by full recomputation (tiny models only)."""
    P = {k: torch.tensor(v, dtype=torch.float32) for k, v in params_np.items()}
    ids = torch.as_tensor(prompts, dtype=torch.int64)
    with torch.no_grad():
        for _ in range(max_new_tokens):
            nxt = forward(P, ids, a)[:, -1].argmax(-1, keepdim=True)
            ids = torch.cat([ids, nxt], dim=1)
    return ids[:, prompts.shape[1]:].numpy()
