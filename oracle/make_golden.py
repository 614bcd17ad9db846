"""Generates tests/golden/*.npz by running the REAL HuggingFace/PyTorch path — the libraries the
reference's trainer image wraps — on seeded tiny inputs. Run here (CPU):

    python oracle/make_golden.py

The fixtures pin oracle/llama_oracle.py (tests/test_oracle_golden.py) and are what the CUDA
parity tests are ultimately anchored to. transformers.Trainer itself cannot be imported in this
image (needs `accelerate`), so the Trainer step is written out with the same torch objects
Trainer uses: model(...).loss.backward(); clip_grad_norm_(params, 1.0); AdamW(lr=5e-5,
betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0).step()  (TrainingArguments defaults).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.llama_oracle import Arch, seeded_params  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (arch, batch, seed)
    "llama_tiny_mha": (Arch(256, 256, 384, 2, 2, 2, 128, 128, 1e-5, 10000.0), 2, 11),
    "llama_tiny_gqa": (Arch(320, 512, 256, 1, 4, 2, 128, 256, 1e-6, 10000.0), 1, 12),
}
SAMPLE_STRIDE = 61  # strided samples of big tensors keep the fixtures small


def make_batch(a: Arch, B: int, seed: int):
    rng = np.random.default_rng(seed + 1000)
    ids = rng.integers(0, a.vocab_size, size=(B, a.max_seq_len), dtype=np.int64)
    labels = ids.copy()
    # mask a prompt-like prefix and a few scattered tokens with the HF ignore index
    labels[:, : a.max_seq_len // 8] = -100
    labels[rng.random(labels.shape) < 0.05] = -100
    return ids, labels


def hf_model(a: Arch, params):
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(
        vocab_size=a.vocab_size, hidden_size=a.hidden_size, intermediate_size=a.intermediate_size,
        num_hidden_layers=a.num_layers, num_attention_heads=a.num_heads,
        num_key_value_heads=a.num_kv_heads, head_dim=a.head_dim,
        max_position_embeddings=a.max_seq_len, rms_norm_eps=a.rms_norm_eps,
        rope_parameters={"rope_type": "default", "rope_theta": a.rope_theta},
        tie_word_embeddings=False, attention_bias=False, mlp_bias=False, attention_dropout=0.0,
    )
    cfg._attn_implementation = "sdpa"
    model = LlamaForCausalLM(cfg).float()
    sd = {k: torch.tensor(v) for k, v in params.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    return model


def run_case(name, a: Arch, B: int, seed: int):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    params = seeded_params(a, seed)
    ids, labels = make_batch(a, B, seed)
    model = hf_model(a, params)
    model.train()
    out = model(input_ids=torch.tensor(ids), labels=torch.tensor(labels))
    loss = out.loss
    loss.backward()
    named = dict(model.named_parameters())
    grads = {k: p.grad.detach().clone() for k, p in named.items()}
    gnorm = float(torch.nn.utils.clip_grad_norm_(list(named.values()), 1.0))
    opt = torch.optim.AdamW(list(named.values()), lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    opt.step()
    # second step with a fresh batch, to exercise non-zero Adam moments and bias correction
    opt.zero_grad(set_to_none=True)
    ids2, labels2 = make_batch(a, B, seed + 7)
    out2 = model(input_ids=torch.tensor(ids2), labels=torch.tensor(labels2))
    out2.loss.backward()
    gnorm2 = float(torch.nn.utils.clip_grad_norm_(list(named.values()), 1.0))
    for g in opt.param_groups:
        g["lr"] = 2.5e-5
    opt.step()

    fx = dict(
        arch=np.array([a.vocab_size, a.hidden_size, a.intermediate_size, a.num_layers, a.num_heads,
                       a.num_kv_heads, a.head_dim, a.max_seq_len], dtype=np.int64),
        arch_f=np.array([a.rms_norm_eps, a.rope_theta], dtype=np.float64),
        batch=np.array([B, seed], dtype=np.int64),
        ids=ids, labels=labels, ids2=ids2, labels2=labels2,
        loss=np.float32(loss.item()), gnorm=np.float32(gnorm),
        loss2=np.float32(out2.loss.item()), gnorm2=np.float32(gnorm2),
        logits=out.logits.detach().numpy().astype(np.float32),
    )
    for k in named:
        fx["gradnorm/" + k] = np.float32(grads[k].norm().item())
        fx["grad/" + k] = grads[k].flatten()[::SAMPLE_STRIDE].numpy().copy()
        fx["param2/" + k] = named[k].detach().flatten()[::SAMPLE_STRIDE].numpy().copy()
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **fx)
    print(f"{name}: loss {loss.item():.6f} gnorm {gnorm:.6f} loss2 {out2.loss.item():.6f} -> {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


def run_ops():
    """Per-op fixtures from the HF modules themselves."""
    from transformers.loss.loss_utils import ForCausalLMLoss
    from transformers.models.llama.modeling_llama import (LlamaRMSNorm, LlamaRotaryEmbedding,
                                                          apply_rotary_pos_emb)
    from transformers import LlamaConfig

    rng = np.random.default_rng(5)
    fx = {}
    x = rng.standard_normal((6, 256)).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(256)).astype(np.float32)
    n = LlamaRMSNorm(256, eps=1e-5)
    n.weight.data = torch.tensor(w)
    fx["rms_x"], fx["rms_w"] = x, w
    fx["rms_y"] = n(torch.tensor(x)).detach().numpy()

    cfg = LlamaConfig(hidden_size=256, num_attention_heads=2, head_dim=128,
                      max_position_embeddings=512,
                      rope_parameters={"rope_type": "default", "rope_theta": 10000.0})
    rot = LlamaRotaryEmbedding(cfg)
    q = rng.standard_normal((1, 2, 512, 128)).astype(np.float32)
    k = rng.standard_normal((1, 2, 512, 128)).astype(np.float32)
    pos = torch.arange(512)[None]
    cos, sin = rot(torch.tensor(q), pos)
    qe, ke = apply_rotary_pos_emb(torch.tensor(q), torch.tensor(k), cos, sin)
    sel = np.array([0, 1, 2, 3, 64, 127, 128, 255, 300, 511])  # positions kept in the fixture
    fx["rope_pos"] = sel
    fx["rope_q"], fx["rope_k"] = q[:, :, sel], k[:, :, sel]
    fx["rope_qe"], fx["rope_ke"] = qe.numpy()[:, :, sel], ke.numpy()[:, :, sel]
    fx["rope_cos"], fx["rope_sin"] = cos[0].numpy()[sel], sin[0].numpy()[sel]

    qa = rng.standard_normal((1, 2, 128, 128)).astype(np.float32)
    ka = rng.standard_normal((1, 1, 128, 128)).astype(np.float32)
    va = rng.standard_normal((1, 1, 128, 128)).astype(np.float32)
    o = torch.nn.functional.scaled_dot_product_attention(
        torch.tensor(qa), torch.tensor(ka).repeat_interleave(2, 1), torch.tensor(va).repeat_interleave(2, 1),
        is_causal=True, scale=128 ** -0.5)
    fx["att_q"], fx["att_k"], fx["att_v"], fx["att_o"] = qa, ka, va, o.numpy()

    lg = (3 * rng.standard_normal((2, 16, 64))).astype(np.float32)
    lb = rng.integers(0, 64, size=(2, 16), dtype=np.int64)
    lb[0, :3] = -100
    lb[1, 7] = -100
    loss = ForCausalLMLoss(torch.tensor(lg), torch.tensor(lb), vocab_size=64)
    loss_n = ForCausalLMLoss(torch.tensor(lg), torch.tensor(lb), vocab_size=64,
                             num_items_in_batch=torch.tensor(40))
    fx["ce_logits"], fx["ce_labels"] = lg, lb
    fx["ce_loss"], fx["ce_loss_items40"] = np.float32(loss.item()), np.float32(loss_n.item())
    path = os.path.join(OUT, "llama_ops.npz")
    np.savez_compressed(path, **fx)
    print(f"ops -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_falcon():
    """Tiny falcon-7b-layout model: prompt logits + greedy generation from the real HF classes."""
    from transformers import FalconConfig, FalconForCausalLM
    from oracle import falcon_oracle as FO

    a = FO.FalconArch(vocab_size=512, hidden_size=256, num_layers=2, num_heads=4, head_dim=64)
    params = FO.seeded_params(a, 21)
    cfg = FalconConfig(vocab_size=a.vocab_size, hidden_size=a.hidden_size, num_hidden_layers=a.num_layers,
                       num_attention_heads=a.num_heads, multi_query=True, parallel_attn=True, bias=False,
                       new_decoder_architecture=False, alibi=False, layer_norm_epsilon=1e-5,
                       max_position_embeddings=256, tie_word_embeddings=True, hidden_dropout=0.0,
                       attention_dropout=0.0)
    cfg._attn_implementation = "sdpa"
    model = FalconForCausalLM(cfg).float().eval()
    sd = {k: torch.tensor(v) for k, v in params.items()}
    sd["lm_head.weight"] = sd["transformer.word_embeddings.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in m or "inv_freq" in m for m in missing), (missing, unexpected)
    rng = np.random.default_rng(77)
    prompts = rng.integers(0, a.vocab_size, size=(3, 24), dtype=np.int64)
    with torch.no_grad():
        logits = model(torch.tensor(prompts)).logits.numpy()
        gen = model.generate(torch.tensor(prompts), max_new_tokens=12, do_sample=False,
                             pad_token_id=0).numpy()[:, prompts.shape[1]:]
    path = os.path.join(OUT, "falcon_tiny.npz")
    np.savez_compressed(path, arch=np.array([a.vocab_size, a.hidden_size, a.num_layers, a.num_heads, a.head_dim]),
                        seed=np.int64(21), prompts=prompts, logits=logits.astype(np.float32), generated=gen)
    print(f"falcon -> {path} ({os.path.getsize(path) / 1024:.0f} KiB); generated[0] = {gen[0].tolist()}")


def run_opt():
    """Tiny opt-125m-family model through the real OPTForCausalLM: two optimiser
    steps with -100 labels (logits, gradients, updated weights) and a greedy continuation."""
    from transformers import OPTConfig, OPTForCausalLM
    from oracle import opt_oracle as OO

    a = OO.OptArch(vocab_size=192, hidden_size=128, ffn_dim=256, num_layers=2, num_heads=2, max_position_embeddings=64)
    params = OO.seeded_params(a, 31)
    cfg = OPTConfig(vocab_size=a.vocab_size, hidden_size=a.hidden_size, ffn_dim=a.ffn_dim, num_hidden_layers=a.num_layers,
                    num_attention_heads=a.num_heads, max_position_embeddings=a.max_position_embeddings,
                    word_embed_proj_dim=a.hidden_size, do_layer_norm_before=True, activation_function="relu",
                    enable_bias=True, dropout=0.0, attention_dropout=0.0, layerdrop=0.0, tie_word_embeddings=True)
    cfg._attn_implementation = "sdpa"
    torch.manual_seed(0)
    model = OPTForCausalLM(cfg).float()
    sd = {k: torch.tensor(v) for k, v in params.items()}
    sd["lm_head.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    assert model.lm_head.weight.data_ptr() == model.model.decoder.embed_tokens.weight.data_ptr(), "head must be tied"
    rng = np.random.default_rng(32)
    B, S = 2, 48

    def batch():
        ids = rng.integers(0, a.vocab_size, size=(B, S)).astype(np.int64)
        labels = ids.copy()
        labels[0, :9] = -100
        labels[1, 20:27] = -100
        return ids, labels

    ids, labels = batch()
    model.train()
    out = model(input_ids=torch.tensor(ids), labels=torch.tensor(labels))
    out.loss.backward()
    named = dict(model.named_parameters())          # tied weight appears once
    grads = {k: p.grad.detach().clone() for k, p in named.items()}
    gnorm = float(torch.nn.utils.clip_grad_norm_(list(named.values()), 1.0))
    opt = torch.optim.AdamW(list(named.values()), lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    opt.step()
    opt.zero_grad(set_to_none=True)
    ids2, labels2 = batch()
    out2 = model(input_ids=torch.tensor(ids2), labels=torch.tensor(labels2))
    out2.loss.backward()
    gnorm2 = float(torch.nn.utils.clip_grad_norm_(list(named.values()), 1.0))
    for g in opt.param_groups:
        g["lr"] = 2.5e-5
    opt.step()
    fx = dict(arch=np.array([a.vocab_size, a.hidden_size, a.ffn_dim, a.num_layers, a.num_heads, a.max_position_embeddings]),
              seed=np.int64(31), ids=ids, labels=labels, ids2=ids2, labels2=labels2,
              loss=np.float32(out.loss.item()), gnorm=np.float32(gnorm), loss2=np.float32(out2.loss.item()),
              gnorm2=np.float32(gnorm2), logits=out.logits.detach().numpy().astype(np.float32))
    for k in named:
        fx["gradnorm/" + k] = np.float32(grads[k].norm().item())
        fx["grad/" + k] = grads[k].flatten()[::17].numpy().copy()
        fx["param2/" + k] = named[k].detach().flatten()[::17].numpy().copy()
    # greedy continuation from the ORIGINAL weights
    model2 = OPTForCausalLM(cfg).float().eval()
    model2.load_state_dict(sd, strict=False)
    prompts = rng.integers(0, a.vocab_size, size=(3, 16)).astype(np.int64)
    with torch.no_grad():
        gen = model2.generate(torch.tensor(prompts), max_new_tokens=10, do_sample=False, pad_token_id=1).numpy()[:, 16:]
        fx["gen_logits"] = model2(torch.tensor(prompts)).logits.numpy().astype(np.float32)
    fx["prompts"], fx["generated"] = prompts, gen
    path = os.path.join(OUT, "opt_tiny.npz")
    np.savez_compressed(path, **fx)
    print(f"opt -> {path} ({os.path.getsize(path) / 1024:.0f} KiB): loss {out.loss.item():.6f} gnorm {gnorm:.6f} "
          f"loss2 {out2.loss.item():.6f}; generated[0] = {gen[0].tolist()}")


if __name__ == "__main__":
    if "--opt-only" in sys.argv:
        run_opt()
        sys.exit(0)
    run_ops()
    run_falcon()
    run_opt()
    for name, (a, B, seed) in CASES.items():
        run_case(name, a, B, seed)
