"""Host side of the container contract (docs/container-contract.md; params_reconciler.go:28-68):
params.json parsing incl. string-typed numbers and PARAM_* env, prompt templating, packing,
LR schedule, HF-format checkpoint round trip. CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from runbooks_b200 import contract


def test_params_defaults_when_file_is_empty_object(tmp_path):
    p = tmp_path / "params.json"
    p.write_text("{}")          # what the controller writes when spec.params is empty
    tp = contract.load_params(str(p), environ={})
    assert tp.num_train_epochs == 3.0 and tp.learning_rate == 5e-5 and tp.save_steps == 500
    assert tp.per_device_train_batch_size == 8 and tp.max_grad_norm == 1.0 and tp.weight_decay == 0.0


def test_params_int_or_string_and_aliases(tmp_path):
    p = tmp_path / "params.json"
    # json.MarshalIndent of map[string]intstr.IntOrString: numbers may arrive as strings
    p.write_text(json.dumps({"num_train_epochs": 1, "save_steps": "5", "epochs": "2",
                             "learning_rate": "1e-4", "prompt_template": "Q: {prompt}\nA: {completion}",
                             "something_else": "kept"}, indent=2))
    tp = contract.load_params(str(p), environ={"PARAM_SAVE_STEPS": "7", "UNRELATED": "x"})
    assert tp.num_train_epochs == 2.0          # alias from examples/facebook-opt-125m
    assert tp.save_steps == 7                  # PARAM_* env overrides the file
    assert tp.learning_rate == 1e-4
    assert tp.extra == {"something_else": "kept"}


def test_params_errors(tmp_path):
    p = tmp_path / "params.json"
    p.write_text("[1,2]")
    with pytest.raises(ValueError):
        contract.load_params(str(p), environ={})
    p.write_text(json.dumps({"save_steps": "often"}))
    with pytest.raises(ValueError):
        contract.load_params(str(p), environ={})
    # a missing file behaves like {}
    assert contract.load_params(str(tmp_path / "nope.json"), environ={}).save_steps == 500


def test_linear_schedule_matches_transformers():
    from transformers import get_linear_schedule_with_warmup
    for warm in (0, 3):
        tp = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([tp], lr=5e-5)
        sch = get_linear_schedule_with_warmup(opt, warm, 12)
        for i in range(12):
            assert abs(opt.param_groups[0]["lr"] - contract.linear_lr(i, 12, 5e-5, warm)) < 1e-12
            opt.step()
            sch.step()


def test_render_and_records(tmp_path):
    d = tmp_path / "data"
    d.mkdir()
    (d / "a.jsonl").write_text('{"prompt": "p1", "completion": "c1"}\n\n{"prompt": "p2", "completion": "c2"}\n')
    (d / "b.json").write_text(json.dumps([{"text": "plain"}]))
    recs = list(contract.iter_records(str(d)))
    assert len(recs) == 3
    tpl = "## Instruction\n{prompt}\n## Response:\n{completion}"
    assert contract.render(recs[0], tpl) == "## Instruction\np1\n## Response:\nc1"
    assert contract.render({"text": "plain"}, tpl) == "plain"
    with pytest.raises(FileNotFoundError):
        list(contract.iter_records(str(tmp_path / "empty")))


def test_packing():
    docs = [[5, 6, 7], [8], list(range(10, 20))]
    ids, labels = contract.pack_sequences(docs, 8, bos_id=1, eos_id=2)
    stream = [1, 5, 6, 7, 2, 1, 8, 2, 1] + list(range(10, 20)) + [2]
    assert ids.shape == (3, 8) and ids.dtype == np.int32
    flat = ids.reshape(-1)
    assert list(flat[: len(stream)]) == stream
    assert (flat[len(stream):] == 2).all()                       # tail padded with eos
    assert (labels.reshape(-1)[: len(stream)] == flat[: len(stream)]).all()
    assert (labels.reshape(-1)[len(stream):] == -100).all()      # and ignored by the loss
    with pytest.raises(ValueError):
        contract.pack_sequences([], 8, None, None)


def _tiny_model_dir(tmp_path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from oracle import llama_oracle as O
    from runbooks_b200.engine import LlamaArch

    a = O.Arch(256, 256, 384, 2, 2, 2, 128, 128, 1e-5, 10000.0)
    params = O.seeded_params(a, 3)
    md = tmp_path / "model"
    md.mkdir()
    vocab = {"<s>": 0, "</s>": 1, "<unk>": 2, **{f"w{i}": i + 3 for i in range(253)}}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    tok.save(str(md / "tokenizer.json"))
    (md / "tokenizer_config.json").write_text(json.dumps({"bos_token": "<s>", "eos_token": "</s>"}))
    arch = LlamaArch(256, 256, 384, 2, 2, 2, 128, 128, 1e-5, 10000.0)
    from tests.util import bf16_bits
    contract.save_hf_checkpoint(str(md), arch.to_hf_config(), ((k, bf16_bits(v)) for k, v in params.items()))
    return md, a, params


def test_checkpoint_round_trip_loads_in_hf(tmp_path):
    """Our artifact layout must be loadable as a model input again (the Server later mounts it:
    server_controller.go:184-193) — checked with the real AutoModelForCausalLM."""
    from transformers import AutoModelForCausalLM
    from oracle import llama_oracle as O

    md, a, params = _tiny_model_dir(tmp_path)
    assert sorted(os.listdir(md)) == ["config.json", "model.safetensors", "tokenizer.json", "tokenizer_config.json"]
    back = dict(contract.iter_safetensors(str(md)))
    assert set(back) == set(params)
    for k, v in params.items():
        assert back[k].dtype == np.uint16 and back[k].shape == v.shape
        assert np.array_equal(torch.from_numpy(back[k]).view(torch.bfloat16).float().numpy(), v)
    model = AutoModelForCausalLM.from_pretrained(str(md), torch_dtype=torch.float32)
    ids = torch.randint(0, 256, (1, 128), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        hf = model(ids).logits
        mine = O.forward({k: torch.tensor(v) for k, v in params.items()}, ids, a)
    assert float((hf - mine).norm() / mine.norm()) < 2e-5


def test_checkpoint_sharding(tmp_path, monkeypatch):
    monkeypatch.setattr(contract, "MAX_SHARD_BYTES", 300_000)
    md, a, params = _tiny_model_dir(tmp_path)
    files = sorted(f for f in os.listdir(md) if f.endswith(".safetensors"))
    assert len(files) > 1 and files[0].startswith("model-00001-of-")
    idx = json.load(open(md / "model.safetensors.index.json"))
    assert set(idx["weight_map"]) == set(params)
    assert set(dict(contract.iter_safetensors(str(md)))) == set(params)


def test_tokenizer_and_dataset_build(tmp_path):
    from runbooks_b200 import worker
    md, a, params = _tiny_model_dir(tmp_path)
    d = tmp_path / "data"
    d.mkdir()
    (d / "x.jsonl").write_text("\n".join(json.dumps({"prompt": "w1 w2 w3", "completion": "w4 w5"}) for _ in range(40)))
    tp = contract.TrainParams(prompt_template="{prompt} w9 {completion}")
    ids, labels = worker.build_dataset(tp, str(md), str(d), 128)
    assert ids.shape[1] == 128 and ids.shape == labels.shape
    assert list(ids[0, :8]) == [0, 4, 5, 6, 12, 7, 8, 1]   # <s> w1 w2 w3 w9 w4 w5 </s>
    per_step, spe, total = worker.plan_steps(len(ids), contract.TrainParams(num_train_epochs=1, per_device_train_batch_size=1), 2)
    assert per_step == 2 and spe == len(ids) // 2 and total == spe


def test_serve_arch_from_hf_configs():
    from runbooks_b200.infer import ServeArch
    falcon7b = {"model_type": "falcon", "vocab_size": 65024, "hidden_size": 4544, "num_hidden_layers": 32,
                "num_attention_heads": 71, "multi_query": True, "parallel_attn": True, "bias": False,
                "alibi": False, "new_decoder_architecture": False, "layer_norm_epsilon": 1e-5}
    a = ServeArch.from_hf_config(falcon7b)
    assert (a.family, a.head_dim, a.num_kv_heads, a.intermediate_size, a.tie_embeddings) == ("falcon", 64, 1, 18176, True)
    assert a == ServeArch.falcon_7b(2048)
    with pytest.raises(ValueError):
        ServeArch.from_hf_config(dict(falcon7b, alibi=True))
    with pytest.raises(ValueError):
        ServeArch.from_hf_config({"model_type": "gpt2"})
    # the reference's system-test model (test/system.sh:46-78, examples/facebook-opt-125m/base-server.yaml)
    opt125m = {"model_type": "opt", "vocab_size": 50272, "hidden_size": 768, "ffn_dim": 3072, "num_hidden_layers": 12,
               "num_attention_heads": 12, "max_position_embeddings": 2048, "do_layer_norm_before": True,
               "word_embed_proj_dim": 768, "activation_function": "relu", "enable_bias": True, "pad_token_id": 1}
    o = ServeArch.from_hf_config(opt125m)
    assert (o.family, o.head_dim, o.num_kv_heads, o.intermediate_size, o.tie_embeddings, o.max_positions, o.max_ctx) == (
        "opt", 64, 12, 3072, True, 2048, 2048)
    with pytest.raises(ValueError):          # opt-350m's post-LN / projected-embedding layout is not built
        ServeArch.from_hf_config(dict(opt125m, do_layer_norm_before=False))
    with pytest.raises(ValueError):
        ServeArch.from_hf_config(dict(opt125m, word_embed_proj_dim=512))
    ll = ServeArch.from_hf_config({"model_type": "llama", "vocab_size": 32000, "hidden_size": 4096,
                                   "intermediate_size": 11008, "num_hidden_layers": 32, "num_attention_heads": 32,
                                   "rms_norm_eps": 1e-5, "max_position_embeddings": 4096})
    assert (ll.family, ll.num_kv_heads, ll.head_dim, ll.max_ctx, ll.tie_embeddings) == ("llama", 32, 128, 4096, False)


def test_unimplemented_checkpoint_variants_fail_instead_of_training_something_else():
    """a Llama-3.1 rope_scaling, a biased variant or another activation must raise, not be
    ignored (the Job would exit 0 with different arithmetic)."""
    from runbooks_b200.engine import LlamaArch, OptArch, arch_from_hf_config
    base = {"model_type": "llama", "vocab_size": 32000, "hidden_size": 4096, "intermediate_size": 11008,
            "num_hidden_layers": 32, "num_attention_heads": 32, "rms_norm_eps": 1e-5, "max_position_embeddings": 4096}
    a = arch_from_hf_config(dict(base, pad_token_id=0))
    assert isinstance(a, LlamaArch) and a.pad_token_id == 0 and arch_from_hf_config(base).pad_token_id == -1
    for bad in ({"rope_scaling": {"rope_type": "llama3", "factor": 8.0}}, {"rope_parameters": {"rope_type": "linear", "factor": 2}},
                {"attention_bias": True}, {"mlp_bias": True}, {"hidden_act": "gelu"}, {"tie_word_embeddings": True}):
        with pytest.raises(ValueError):
            arch_from_hf_config(dict(base, **bad))
    assert arch_from_hf_config(dict(base, rope_scaling=None, rope_parameters={"rope_type": "default", "rope_theta": 5e5})).rope_theta == 5e5
    o = arch_from_hf_config({"model_type": "opt", "vocab_size": 50272, "hidden_size": 768, "ffn_dim": 3072,
                             "num_hidden_layers": 12, "num_attention_heads": 12, "max_position_embeddings": 2048})
    assert isinstance(o, OptArch) and (o.head_dim, o.pad_token_id, o.max_seq_len) == (64, 1, 2048)
    with pytest.raises(ValueError):
        arch_from_hf_config({"model_type": "mistral"})
    from runbooks_b200.engine import FalconArch
    fbase = {"model_type": "falcon", "vocab_size": 65024, "hidden_size": 4544, "num_hidden_layers": 32,
             "num_attention_heads": 71, "multi_query": True, "parallel_attn": True, "bias": False, "alibi": False,
             "new_decoder_architecture": False, "pad_token_id": 11}
    fa = arch_from_hf_config(fbase, 1024)
    assert isinstance(fa, FalconArch) and (fa.head_dim, fa.num_kv_heads, fa.intermediate_size, fa.max_seq_len) == (64, 1, 18176, 1024)
    assert fa.pad_token_id == -1          # FalconModel's nn.Embedding has no padding_idx, whatever the config says
    assert fa == FalconArch.falcon_7b(1024)
    for bad in ({"alibi": True}, {"bias": True}, {"new_decoder_architecture": True}, {"parallel_attn": False},
                {"multi_query": False}, {"activation": "relu"}, {"tie_word_embeddings": False},
                {"rope_parameters": {"rope_type": "linear", "factor": 2.0}}):
        with pytest.raises(ValueError):
            arch_from_hf_config(dict(fbase, **bad))
    # tensors the engine would silently drop
    assert contract.is_ignorable_tensor("lm_head.weight", {"model_type": "opt"})
    assert contract.is_ignorable_tensor("model.layers.0.self_attn.rotary_emb.inv_freq", {"model_type": "llama"})
    assert not contract.is_ignorable_tensor("model.layers.0.self_attn.q_proj.bias", {"model_type": "llama"})
    assert not contract.is_ignorable_tensor("lm_head.weight", {"model_type": "llama"})
    assert contract.canonical_tensor_name("decoder.embed_tokens.weight", {"model_type": "opt"}) == "model.decoder.embed_tokens.weight"


def test_warmup_semantics_match_training_arguments(tmp_path):
    """warmup_steps >= 1: exact steps; in [0, 1): ratio of the total, rounded up; warmup_ratio: deprecated
    alias. Checked against the real TrainingArguments.get_warmup_steps (the class itself cannot be
    instantiated here -- it needs `accelerate` -- but the method only reads self.warmup_steps)."""
    from types import SimpleNamespace
    from transformers import TrainingArguments
    for w in (0, 0.03, 0.1, 0.25, 0.999, 1, 7, 250):
        for total in (1, 7, 33, 100, 1000):
            want = TrainingArguments.get_warmup_steps(SimpleNamespace(warmup_steps=w), total)
            assert contract.warmup_steps_for(total, w) == want, (w, total)
    f = TrainingArguments.__dataclass_fields__
    d = contract.TrainParams()
    assert f["warmup_steps"].default == d.warmup_steps == 0
    assert f["optim"].default == d.optim and d.optim in ("adamw_torch", "adamw_torch_fused")
    assert f["label_smoothing_factor"].default == d.label_smoothing_factor == 0.0
    assert f["average_tokens_across_devices"].default is True
    p = tmp_path / "params.json"
    p.write_text(json.dumps({"warmup_ratio": "0.1", "max_steps": 50}))
    q = contract.load_params(str(p), environ={})
    assert q.warmup_steps == 0.1 and contract.warmup_steps_for(50, q.warmup_steps) == 5
    p.write_text(json.dumps({"warmup_steps": 12}))
    assert contract.warmup_steps_for(50, contract.load_params(str(p), environ={}).warmup_steps) == 12
    # the schedule with warm-up, against transformers' own lambda
    import torch
    from transformers import get_linear_schedule_with_warmup
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=2e-4)
    sch = get_linear_schedule_with_warmup(opt, num_warmup_steps=5, num_training_steps=50)
    for i in range(50):
        assert abs(contract.linear_lr(i, 50, 2e-4, 5) - sch.get_last_lr()[0]) < 1e-12
        opt.step()
        sch.step()


def test_unimplemented_arithmetic_changing_params_fail_loudly(tmp_path):
    """The Job's exit code is the whole protocol: a TrainingArguments value that would change the arithmetic and
    is not implemented must not be ignored. Unknown harmless keys are kept in `extra` (and logged by the worker)."""
    p = tmp_path / "params.json"
    for bad in ({"optim": "adafactor"}, {"lr_scheduler_type": "cosine"}, {"label_smoothing_factor": 0.1},
                {"average_tokens_across_devices": False}, {"average_tokens_across_devices": "false"}, {"warmup_steps": -1}):
        p.write_text(json.dumps(bad))
        with pytest.raises(ValueError):
            contract.load_params(str(p), environ={})
    for ok in ({"optim": "adamw_torch_fused"}, {"average_tokens_across_devices": True}, {"average_tokens_across_devices": "True"},
               {"bf16": True, "gradient_checkpointing": "true", "report_to": "none"}):
        p.write_text(json.dumps(ok))
        q = contract.load_params(str(p), environ={})
        assert set(q.extra) == set(ok) - set(contract.TrainParams.__dataclass_fields__)
