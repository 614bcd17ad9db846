"""worker.train_rank's host logic on CPU, with the CUDA engine replaced by a recording stub (the product has
no CPU engine; this tests everything AROUND b200w_train_step): params -> steps, the linear schedule, the
N-rank sharding rule of SURVEY.md 8e (same permutation on every rank, rank r takes [r::N]), checkpoints
every save_steps, the final artifacts directory, and a non-finite loss turning into a failure (the Job's
exit code is the whole protocol: internal/controller/utils.go:37-49)."""
import json
import os

import numpy as np
import pytest

from runbooks_b200 import contract, worker
from tests.test_contract import _tiny_model_dir

CALLS = {}


class StubEngine:
    """records what each rank is asked to do; weights stay what was loaded"""

    def __init__(self, device=0):
        self.rank_log = CALLS.setdefault(device, dict(steps=[], comm=None, loaded=set()))
        self.weights, self.arch = {}, None

    def init_model(self, arch, **kw):
        self.arch, self.rank_log["hparams"] = arch, kw

    def params(self):
        from oracle import llama_oracle as O
        a = self.arch
        return list(O.param_shapes(O.Arch(a.vocab_size, a.hidden_size, a.intermediate_size, a.num_layers, a.num_heads,
                                          a.num_kv_heads, a.head_dim, a.max_seq_len)).items())

    def load_tensor(self, name, arr):
        self.weights[name] = np.array(arr)
        self.rank_log["loaded"].add(name)

    def comm_init(self, rank, world, uid):
        self.rank_log["comm"] = (rank, world, bytes(uid))

    def train_step(self, ids, labels, lr=0.0):
        step = len(self.rank_log["steps"])
        self.rank_log["steps"].append(dict(ids=np.array(ids), labels=np.array(labels), lr=lr))
        loss = float("nan") if CALLS.get("nan_at") == step else 3.0 - 0.1 * step
        return loss, 1.5

    def read_tensor(self, name, shape, bf16_bits=False):
        return self.weights[name].reshape(shape)

    def device_bytes(self):
        return 0

    def close(self):
        self.rank_log["closed"] = True


@pytest.fixture()
def content(tmp_path, monkeypatch):
    CALLS.clear()
    md, a, params = _tiny_model_dir(tmp_path)
    (tmp_path / "data").mkdir()
    rng = np.random.default_rng(1)
    with open(tmp_path / "data" / "train.jsonl", "w") as f:
        for _ in range(120):
            w = [f"w{i}" for i in rng.integers(0, 250, size=30)]
            f.write(json.dumps({"prompt": " ".join(w[:18]), "completion": " ".join(w[18:])}) + "\n")
    import runbooks_b200.engine as eng_mod
    monkeypatch.setattr(eng_mod, "Engine", StubEngine)
    return tmp_path


def _write_params(root, **kw):
    (root / "params.json").write_text(json.dumps(kw))


def test_two_ranks_plan_the_same_steps_and_partition_every_global_batch(content, capsys):
    _write_params(content, max_steps=5, per_device_train_batch_size=2, max_seq_length=128, learning_rate="2e-4",
                  save_steps=2, logging_steps=1)
    uid = bytes(range(128))
    worker.train_rank(1, 2, uid, str(content))           # rank 1 first: it must not write anything
    assert not (content / "artifacts").exists() or not os.listdir(content / "artifacts")
    worker.train_rank(0, 2, uid, str(content))
    r0, r1 = CALLS[0], CALLS[1]
    assert r0["comm"] == (0, 2, uid) and r1["comm"] == (1, 2, uid)
    assert len(r0["steps"]) == len(r1["steps"]) == 5
    # the single-process run on the same global batch: world 1, per-device batch 4
    CALLS.clear()
    _write_params(content, max_steps=5, per_device_train_batch_size=4, max_seq_length=128, learning_rate="2e-4", save_steps=0)
    worker.train_rank(0, 1, b"", str(content))
    one = CALLS[0]["steps"]
    for s0, s1, g in zip(r0["steps"], r1["steps"], one):
        assert s0["lr"] == s1["lr"] == g["lr"]                                  # same schedule everywhere
        assert s0["ids"].shape == s1["ids"].shape == (2, 128)
        assert np.array_equal(g["ids"][0::2], s0["ids"]) and np.array_equal(g["ids"][1::2], s1["ids"])   # [r::N]
        assert np.array_equal(g["labels"][0::2], s0["labels"]) and np.array_equal(g["labels"][1::2], s1["labels"])
    lrs = [s["lr"] for s in r0["steps"]]
    assert lrs == [contract.linear_lr(i, 5, 2e-4, 0) for i in range(5)] and lrs[0] == 2e-4 and lrs[-1] < lrs[0]
    # prompt tokens are masked, completions are targets
    lab = r0["steps"][0]["labels"]
    assert (lab == -100).any() and (lab != -100).any()


def test_gradient_checkpointing_param_selects_activation_recomputation(content, capsys):
    """TrainingArguments.gradient_checkpointing (string or bool in params.json) reaches the engine as recompute."""
    for value, want in (("true", True), (True, True), ("false", False), (None, False)):
        CALLS.clear()
        kw = dict(max_steps=1, per_device_train_batch_size=2, max_seq_length=128, save_steps=0)
        if value is not None:
            kw["gradient_checkpointing"] = value
        _write_params(content, **kw)
        worker.train_rank(0, 1, b"", str(content))
        assert CALLS[0]["hparams"]["recompute"] is want, (value, CALLS[0]["hparams"])
    _write_params(content, max_steps=1, gradient_checkpointing="sometimes")
    with pytest.raises(ValueError):
        worker.train_rank(0, 1, b"", str(content))


def test_checkpoints_final_artifacts_and_log_lines(content, capsys):
    _write_params(content, max_steps=5, per_device_train_batch_size=2, max_seq_length=128, save_steps=2, logging_steps=1)
    worker.train_rank(0, 1, b"", str(content))
    art = content / "artifacts"
    assert sorted(p for p in os.listdir(art) if p.startswith("checkpoint-")) == ["checkpoint-2", "checkpoint-4"]
    for d in (art, art / "checkpoint-2"):
        assert {"config.json", "model.safetensors", "tokenizer.json", "trainer_state.json"} <= set(os.listdir(d))
    assert json.loads((art / "trainer_state.json").read_text())["global_step"] == 5
    back = dict(contract.iter_safetensors(str(art)))                          # loadable as the next /content/model
    assert set(back) == CALLS[0]["loaded"]
    lines = [json.loads(l) for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert [l["step"] for l in lines if "step" in l] == [1, 2, 3, 4, 5]
    assert lines[0]["event"] == "start" and lines[0]["world_size"] == 1 and lines[-1]["event"] == "done"
    assert CALLS[0]["closed"]


def test_non_finite_loss_fails_the_job(content):
    _write_params(content, max_steps=4, per_device_train_batch_size=2, max_seq_length=128, save_steps=0)
    CALLS["nan_at"] = 2
    with pytest.raises(FloatingPointError):
        worker.train_rank(0, 1, b"", str(content))
    assert len(CALLS[0]["steps"]) == 3
    # through the real entry point the exception becomes exit code 1
    CALLS.clear()
    CALLS["nan_at"] = 0
    os.environ["B200W_NUM_GPUS"] = "1"
    try:
        assert worker.main(["train", "--content", str(content)]) == 1
    finally:
        del os.environ["B200W_NUM_GPUS"]


def test_sequence_length_must_suit_the_kernels(content):
    _write_params(content, max_steps=1, max_seq_length=100)
    with pytest.raises(ValueError, match="multiple of 128"):
        worker.train_rank(0, 1, b"", str(content))
