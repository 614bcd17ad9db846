"""N-rank data-parallel diagnostic (torchrun, one rank per GPU): real-width Llama layers, a few
steps, prints per-rank loss / grad-norm per step and checks that (a) every value is finite,
(b) the ranks hold bit-identical weights after every step, (c) the N-rank step equals a 1-rank step
on the concatenated batch (rank 0 recomputes it on a second engine context when --check-global).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/n2_debug.py --layers 2
"""
import argparse
import hashlib
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from runbooks_b200.engine import Engine, LlamaArch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--nseq", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--ragged", action="store_true", help="different numbers of target tokens per rank")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(local)
    arch = LlamaArch.llama2_7b(args.seq)
    arch.num_layers = args.layers
    e = Engine(local)
    e.init_model(arch, micro_batch=1, training=True)
    e.init_random(seed=0, std=0.02)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(e.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)
    if world > 1:
        e.comm_init(rank, world, bytes(uid.numpy().tobytes()))
    names = [(n, s) for n, s in e.params()]
    probe = [names[0], names[len(names) // 2], names[-1]]
    ok = True
    for step in range(args.steps):
        g = torch.Generator().manual_seed(100 * step + 7)
        ids_all = torch.randint(0, arch.vocab_size, (world * args.nseq, args.seq), generator=g, dtype=torch.int32).numpy()
        lab_all = ids_all.copy()
        if args.ragged:  # rank r's rows lose their first (r+1)*S/4 targets, like prompt masking
            for r in range(world):
                lab_all[r::world, : (r + 1) * args.seq // 4] = -100
        ids, lab = ids_all[rank::world], lab_all[rank::world]
        loss, gn = e.train_step(np.ascontiguousarray(ids), np.ascontiguousarray(lab), lr=1e-3)
        digest = hashlib.sha1(b"".join(e.read_tensor(n, s, bf16_bits=True).tobytes() for n, s in probe)).hexdigest()[:12]
        report = []
        if not (math.isfinite(loss) and math.isfinite(gn)):
            # where are the non-finite gradients? (g survives the optimizer step until the next backward)
            for n, shp in names:
                gr = e.read_state(n, shp, "grad").reshape(-1)
                bad = np.flatnonzero(~np.isfinite(gr))
                if bad.size:
                    cols = shp[-1] if len(shp) > 1 else 1
                    runs = int(np.count_nonzero(np.diff(bad) != 1)) + 1
                    report.append(f"{n} shape={tuple(shp)} bad={bad.size} first={int(bad[0])} (row {int(bad[0]) // cols}, col {int(bad[0]) % cols}) "
                                  f"last={int(bad[-1])} contiguous_runs={runs} nan={int(np.isnan(gr[bad]).sum())} inf={int(np.isinf(gr[bad]).sum())}")
        if report:
            # one line per layer: which of its gradient tensors are non-finite, in BACKWARD order of
            # production (down, gate|up, ln2, o, q|k|v, ln1): the first bad letter from the top layer
            # down names the op where the activation-gradient stream went bad
            badnames = {r.split(" ")[0] for r in report}
            order = [("D", "mlp.down_proj"), ("G", "mlp.gate_proj"), ("U", "mlp.up_proj"), ("2", "post_attention_layernorm"),
                     ("O", "self_attn.o_proj"), ("Q", "self_attn.q_proj"), ("K", "self_attn.k_proj"), ("V", "self_attn.v_proj"),
                     ("1", "input_layernorm")]
            sig = []
            for l in range(arch.num_layers - 1, -1, -1):
                sig.append(f"{l}:" + "".join(c if f"model.layers.{l}.{n}.weight" in badnames else "." for c, n in order))
            extra = [n for n in ("lm_head.weight", "model.norm.weight", "model.embed_tokens.weight") if n in badnames]
            report.insert(0, "SIGNATURE (layer:DGU2OQKV1, top layer first) " + " ".join(sig) + f" | also: {extra}")
        rows = [None] * world
        dist.all_gather_object(rows, (rank, loss, gn, digest, report))
        if rank == 0:
            for r in rows:
                print(f"step {step} rank {r[0]}: loss {r[1]:.6f} grad_norm {r[2]:.6f} weights {r[3]}", flush=True)
            same = len({r[3] for r in rows}) == 1
            finite = all(math.isfinite(r[1]) and math.isfinite(r[2]) for r in rows)
            same_scal = len({(r[1], r[2]) for r in rows}) == 1
            print(f"   finite={finite} ranks_identical_weights={same} ranks_identical_scalars={same_scal}", flush=True)
            ok = ok and same and finite and same_scal
            for r in rows:
                for line in r[4][:6]:
                    print(f"   NONFINITE rank {r[0]}: {line}", flush=True)
                if len(r[4]) > 6:
                    print(f"   NONFINITE rank {r[0]}: ... {len(r[4])} tensors in total", flush=True)
        # every rank holds the gathered rows: stop together (a lone break would hang the peer in NCCL)
        if not all(math.isfinite(r[1]) and math.isfinite(r[2]) for r in rows):
            break
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    if rank == 0:
        print("N2_DEBUG", "OK" if ok else "FAILED", f"mode={os.environ.get('B200W_AR_MODE', 'overlap')}", flush=True)
    e.close()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
