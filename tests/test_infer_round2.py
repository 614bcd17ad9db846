"""Round-2 Server-path tests on the GPU: config #4 at TRUE Falcon-7B layer width, the one-pass prefill
against the token-by-token decode path and the oracles, and the reference's system-test model family
(OPT, test/system.sh:46-78) through the serve engine.

Tolerances as tests/test_infer.py: logits 1.5e-2 relative Frobenius (bf16 compute vs fp32 golden); greedy
ids IDENTICAL wherever the fp32 top-2 margin exceeds twice the logit error bound (north_star: bit-exact
argmax)."""
import numpy as np
import pytest
import torch

from oracle import falcon_oracle as FO
from oracle import llama_oracle as LO
from oracle import opt_oracle as OO

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _falcon_width():
    fx = np.load("tests/golden/falcon_7b_width.npz")
    v, d, L, H, dh = (int(x) for x in fx["arch"])
    a = FO.FalconArch(vocab_size=v, hidden_size=d, num_layers=L, num_heads=H, head_dim=dh)
    return fx, a, FO.seeded_params(a, int(fx["seed"]), std=float(fx["std"]))


def _serve_falcon(a, params, max_ctx=256, max_batch=4):
    from runbooks_b200.infer import InferEngine, ServeArch
    arch = ServeArch("falcon", a.vocab_size, a.hidden_size, a.ffn, a.num_layers, a.num_heads, 1, a.head_dim,
                     max_ctx=max_ctx, norm_eps=a.layer_norm_epsilon, rope_theta=a.rope_theta, tie_embeddings=True)
    e = InferEngine(0)
    e.init_infer(arch, max_batch=max_batch)
    e.infer_load_state_dict(params)
    return e


def _check_greedy(outs, gen, margins, bound, tag):
    """ids identical to HF's wherever the fp32 margin is above the error bound; a flip is legitimate only at
    a near-tie, and everything after a flip is a different continuation (not compared)."""
    agree = total = 0
    for b in range(gen.shape[0]):
        for i in range(gen.shape[1]):
            total += 1
            if outs[b][i] != int(gen[b, i]):
                assert margins[b, i] < bound[b], (tag, b, i, outs[b], gen[b].tolist(), float(margins[b, i]))
                total += gen.shape[1] - i - 1
                break
            agree += 1
    return agree, total


def test_falcon_7b_width_decode_and_prefill_match_hf():
    """d 4544, 71 query heads + 1 kv head of 64, ffn 18176, V 65024 (2 layers): 71 heads = one tensor-core
    tile of the MQA decode attention, d not a multiple of 128, split-K at K = 22720 (fused [dense|4h_to_h]),
    fused [qkv|h_to_4h] N = 22848. Golden from the real FalconForCausalLM."""
    from runbooks_b200.infer import Generator
    fx, a, params = _falcon_width()
    prompts, gen, margins = fx["prompts"], fx["generated"], fx["margins"]
    B, P = prompts.shape
    stride = int(fx["logits_stride"])
    # (1) token-by-token decode of the prompt: logits of the last prompt position vs HF
    e = _serve_falcon(a, params)
    for t in range(P):
        _, lg = e.step(prompts[:, t], [t] * B, list(range(B)), want_logits=(t == P - 1))
    err_dec = max(float(np.linalg.norm(lg[b, ::stride] - fx["logits_last"][b]) / np.linalg.norm(fx["logits_last"][b]))
                  for b in range(B))
    # (2) one-pass prefill of the same prompts into fresh slots
    e2 = _serve_falcon(a, params)
    nxt, lg2 = e2.prefill([p.tolist() for p in prompts], list(range(B)), want_logits=True)
    err_pre = max(float(np.linalg.norm(lg2[b, ::stride] - fx["logits_last"][b]) / np.linalg.norm(fx["logits_last"][b]))
                  for b in range(B))
    print(f"falcon-7b width: last-position logits rel_err decode {err_dec:.3e}, prefill {err_pre:.3e}; "
          f"decode vs prefill {rel(lg, lg2):.3e}")
    assert err_dec < 1.5e-2 and err_pre < 1.5e-2
    # (3) greedy continuation, both ways, against HF's ids
    bound = 2 * 1.5e-2 * fx["logits_last_absmax"] * 2
    n_new = gen.shape[1]
    outs_pre = Generator(_serve_falcon(a, params)).generate([p.tolist() for p in prompts], n_new)
    outs_dec = Generator(_serve_falcon(a, params), use_prefill=False).generate([p.tolist() for p in prompts], n_new)
    ag_p, tot = _check_greedy(outs_pre, gen, margins, bound, "prefill")
    ag_d, _ = _check_greedy(outs_dec, gen, margins, bound, "decode")
    print(f"falcon-7b width greedy: prefill path {ag_p}/{tot}, decode-only path {ag_d}/{tot} ids identical to HF; "
          f"margins min {float(margins.min()):.3f}")
    assert ag_p >= 0.7 * tot and ag_d >= 0.7 * tot
    e.close(); e2.close()


def test_prefill_equals_token_by_token_decode_llama_gqa():
    """Llama family (GQA 4:2 -> CUDA-core decode attention, dh 128: no head padding in the prefill):
    the K/V cache written by the prefill must serve later decode steps exactly like one written token
    by token. Ragged prompt lengths, non-contiguous slots."""
    from runbooks_b200.infer import InferEngine, ServeArch
    oa = LO.Arch(320, 512, 256, 2, 4, 2, 128, 256, 1e-6, 10000.0)
    params = LO.seeded_params(oa, 5)
    arch = ServeArch("llama", oa.vocab_size, oa.hidden_size, oa.intermediate_size, oa.num_layers, oa.num_heads,
                     oa.num_kv_heads, oa.head_dim, max_ctx=256, norm_eps=oa.rms_norm_eps, tie_embeddings=False)
    rng = np.random.default_rng(9)
    lens, slots = [150, 37, 128], [5, 0, 2]
    seqs = [rng.integers(0, oa.vocab_size, size=n + 6) for n in lens]     # prompt + 6 teacher-forced tokens
    with torch.no_grad():
        ref = [LO.forward({k: torch.tensor(v) for k, v in params.items()}, torch.tensor(s[None]), oa)[0].numpy() for s in seqs]
    e = InferEngine(0)
    e.init_infer(arch, max_batch=8)
    e.infer_load_state_dict(params)
    nxt, lg = e.prefill([s[:n].tolist() for s, n in zip(seqs, lens)], slots, want_logits=True)
    worst = max(rel(lg[i], ref[i][lens[i] - 1]) for i in range(3))
    for i in range(3):
        assert int(nxt[i]) == int(lg[i].argmax())
    for t in range(6):                                                    # decode on top of the prefilled cache
        _, lg = e.step([s[n + t] for s, n in zip(seqs, lens)], [n + t for n in lens], slots, want_logits=True)
        worst = max(worst, max(rel(lg[i], ref[i][lens[i] + t]) for i in range(3)))
    print(f"llama prefill + decode: worst logits rel_err vs the fp32 oracle {worst:.3e}")
    assert worst < 1.5e-2
    e.close()


def _opt():
    fx = np.load("tests/golden/opt_tiny.npz")
    V, d, f, L, H, P = (int(x) for x in fx["arch"])
    oa = OO.OptArch(V, d, f, L, H, P)
    return fx, oa, OO.seeded_params(oa, int(fx["seed"]))


def test_opt_serve_greedy_matches_hf():
    """The reference's system-test family through the serve engine: prompt logits and the greedy
    continuation from the real OPTForCausalLM.generate (tests/golden/opt_tiny.npz)."""
    from runbooks_b200.infer import Generator, InferEngine, ServeArch
    fx, oa, params = _opt()
    arch = ServeArch("opt", oa.vocab_size, oa.hidden_size, oa.ffn_dim, oa.num_layers, oa.num_heads, oa.num_heads,
                     oa.head_dim, max_ctx=oa.max_position_embeddings, norm_eps=oa.layer_norm_eps, tie_embeddings=True,
                     max_positions=oa.max_position_embeddings)
    prompts, gen = fx["prompts"], fx["generated"]
    B, P = prompts.shape

    def fresh():
        e = InferEngine(0)
        e.init_infer(arch, max_batch=4)
        e.infer_load_state_dict(params)
        return e

    e = fresh()
    errs = []
    for t in range(P):
        _, lg = e.step(prompts[:, t], [t] * B, list(range(B)), want_logits=True)
        errs.append(rel(lg, fx["gen_logits"][:, t]))
    print(f"opt serve: per-position logits rel_err max {max(errs):.3e}")
    assert max(errs) < 1.5e-2
    _, lgp = fresh().prefill([p.tolist() for p in prompts], list(range(B)), want_logits=True)
    assert rel(lgp, fx["gen_logits"][:, -1]) < 1.5e-2
    n_new = gen.shape[1]
    for use_prefill in (True, False):
        outs = Generator(fresh(), use_prefill=use_prefill).generate([p.tolist() for p in prompts], n_new)
        agree = 0
        for b in range(B):
            ref_logits = []
            ids = prompts[b].tolist()
            P_ = {k: torch.tensor(v) for k, v in params.items()}
            with torch.no_grad():
                for i in range(n_new):
                    lgr = OO.forward(P_, torch.tensor([ids]), oa)[0, -1].numpy()
                    ref_logits.append(lgr)
                    ids.append(int(gen[b, i]))
            for i in range(n_new):
                if outs[b][i] != int(gen[b, i]):
                    top2 = np.sort(ref_logits[i])[-2:]
                    assert top2[1] - top2[0] < 4 * 1.5e-2 * np.abs(ref_logits[i]).max(), (b, i, outs[b], gen[b].tolist())
                    break
                agree += 1
        print(f"opt serve greedy (prefill={use_prefill}): {agree}/{B * n_new} ids identical to HF generate()")
        assert agree >= 0.8 * B * n_new
    e.close()


def test_decode_attention_long_context_many_blocks():
    """MQA tensor-core decode attention over several 128-key blocks with a ragged last block, and the
    CUDA-core kernel on the same data (G < 4 model) -- both against the oracle's causal forward at the last
    position. Context 300: blocks of 128, 128, 44 keys."""
    from runbooks_b200.infer import InferEngine, ServeArch
    a = FO.FalconArch(vocab_size=512, hidden_size=512, num_layers=1, num_heads=8, head_dim=64)
    params = FO.seeded_params(a, 77, std=0.06)
    arch = ServeArch("falcon", a.vocab_size, a.hidden_size, a.ffn, a.num_layers, a.num_heads, 1, a.head_dim,
                     max_ctx=384, norm_eps=a.layer_norm_epsilon, rope_theta=a.rope_theta, tie_embeddings=True)
    e = InferEngine(0)
    e.init_infer(arch, max_batch=2)
    e.infer_load_state_dict(params)
    rng = np.random.default_rng(4)
    ids = rng.integers(0, a.vocab_size, size=(2, 300))
    with torch.no_grad():
        ref = FO.forward({k: torch.tensor(v) for k, v in params.items()}, torch.tensor(ids), a).numpy()
    nxt, lg = e.prefill([ids[0, :299].tolist(), ids[1, :140].tolist()], [1, 0], want_logits=True)
    assert rel(lg[0], ref[0, 298]) < 1.5e-2 and rel(lg[1], ref[1, 139]) < 1.5e-2
    _, lg = e.step([ids[0, 299], ids[1, 140]], [299, 140], [1, 0], want_logits=True)
    err = max(rel(lg[0], ref[0, 299]), rel(lg[1], ref[1, 140]))
    print(f"decode attention over 3 / 2 key blocks: logits rel_err {err:.3e}")
    assert err < 1.5e-2
    e.close()
