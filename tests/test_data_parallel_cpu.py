"""World-size-2 data parallelism on CPU (gloo): the arithmetic of the N-rank step.

Three formulations of the same gradient must agree on a batch whose ranks hold DIFFERENT numbers
of target tokens (prompt-masked rows):
  engine  what engine.cu fwd_bwd_device does: n_global = all-reduce(sum) of the per-rank target
          counts, every rank back-propagates sum(nll_rank) / n_global, gradients all-reduced (sum);
  hf      HF Trainer under DDP with the TrainingArguments default average_tokens_across_devices=True
          (transformers 5.5 trainer.py:2140-2143 gathers and sums num_items_in_batch, :2013-2018
          multiplies the loss by the world size, DDP averages the gradients), using HF's own
          ForCausalLMLoss on the oracle's logits;
  single  one process on the whole batch (mean over all target tokens).
The formulation the engine had before round 1's fix -- per-rank token mean, then the mean over
ranks -- must NOT agree here, or the test would not discriminate.

Also the host-side sharding rule of SURVEY.md 8e (rank r takes sequences [r::N])."""
import multiprocessing as mp
import os
import socket

import numpy as np
import torch
import torch.distributed as dist

from oracle import llama_oracle as O

ARCH = O.Arch(vocab_size=96, hidden_size=32, intermediate_size=64, num_layers=2, num_heads=2, num_kv_heads=1,
              head_dim=16, max_seq_len=24)


def _batch():
    rng = np.random.default_rng(5)
    ids = rng.integers(0, ARCH.vocab_size, size=(4, ARCH.max_seq_len)).astype(np.int64)
    labels = ids.copy()
    for i, n_prompt in enumerate((3, 17, 5, 11)):   # ragged: rank 0 gets rows 0,2 (few masked), rank 1 rows 1,3
        labels[i, :n_prompt] = -100
    return ids, labels


def _grads(params_np, ids, labels, scale_fn):
    """gradient dict of scale_fn(logits, labels) over the oracle forward"""
    params = {k: torch.tensor(v, dtype=torch.float32, requires_grad=True) for k, v in params_np.items()}
    logits = O.forward(params, torch.as_tensor(ids), ARCH)
    loss = scale_fn(logits, torch.as_tensor(labels))
    loss.backward()
    return float(loss.detach()), {k: p.grad.detach().clone() for k, p in params.items()}


def _n_targets(labels):
    return int((labels[:, 1:] != -100).sum())   # shifted, like engine.cu count_valid


def _rank_main(rank, world, port, q):
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    from transformers.loss.loss_utils import ForCausalLMLoss
    params_np = O.seeded_params(ARCH, seed=3, std=0.08)
    ids, labels = _batch()
    my_ids, my_lab = ids[rank::world], labels[rank::world]
    n_local = torch.tensor([_n_targets(my_lab)], dtype=torch.int64)
    n_global = n_local.clone()
    dist.all_reduce(n_global)                                         # engine.cu global_valid
    n_global = int(n_global)

    def allreduce(gr, mean=False):
        out = {}
        for k in sorted(gr):
            t = gr[k].clone()
            dist.all_reduce(t)
            out[k] = t / world if mean else t
        return out

    # engine: sum(nll_rank) / n_global, SUM all-reduce
    loss_e, g_e = _grads(params_np, my_ids, my_lab, lambda lg, lb: O.causal_lm_loss(lg, lb)[1].sum() / n_global)
    g_e = allreduce(g_e)
    le = torch.tensor([loss_e]); dist.all_reduce(le)                  # the logged loss: sum of the partials
    # hf: ForCausalLMLoss(num_items_in_batch = global count) * world, DDP MEAN
    loss_h, g_h = _grads(params_np, my_ids, my_lab,
                         lambda lg, lb: ForCausalLMLoss(lg, lb, ARCH.vocab_size, num_items_in_batch=n_global) * world)
    g_h = allreduce(g_h, mean=True)
    # old behaviour: per-rank token mean, then mean over ranks
    _, g_o = _grads(params_np, my_ids, my_lab, lambda lg, lb: O.causal_lm_loss(lg, lb)[0])
    g_o = allreduce(g_o, mean=True)
    # single process, whole batch
    loss_s, g_s = _grads(params_np, ids, labels, lambda lg, lb: O.causal_lm_loss(lg, lb)[0])

    def rel(a, b):
        num = sum(float(((a[k] - b[k]).double() ** 2).sum()) for k in a)
        den = sum(float((b[k].double() ** 2).sum()) for k in a)
        return (num / den) ** 0.5

    q.put(dict(rank=rank, n_local=int(n_local), n_global=n_global, engine_vs_single=rel(g_e, g_s),
               hf_vs_single=rel(g_h, g_s), engine_vs_hf=rel(g_e, g_h), old_vs_single=rel(g_o, g_s),
               loss_engine=float(le), loss_single=loss_s))
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_step_equals_the_global_batch_step_and_hf_ddp():
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in rows:
        print(r)
        assert r["n_global"] == sum(x["n_local"] for x in rows)
        assert rows[0]["n_local"] != rows[1]["n_local"], "the batch must be ragged across ranks"
        assert r["engine_vs_single"] < 1e-5          # fp32 reassociation only
        assert r["hf_vs_single"] < 1e-5
        assert r["engine_vs_hf"] < 1e-5
        assert abs(r["loss_engine"] - r["loss_single"]) < 1e-5
        assert r["old_vs_single"] > 1e-2, "per-rank normalisation must be distinguishable on this batch"


def test_rank_sharding_partitions_the_batch():
    """worker.train_rank: `mine = batch[rank::world]` -- every sequence on exactly one rank, equal counts."""
    batch = list(range(16))
    for world in (1, 2, 4, 8):
        shards = [batch[r::world] for r in range(world)]
        assert sorted(sum(shards, [])) == batch
        assert len({len(s) for s in shards}) == 1
