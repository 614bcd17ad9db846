"""Runs the container-contract trainer on N GPUs of this box with a tiny Llama directory and
checks: exit code 0, artifacts written, and that the N-rank result equals the 1-rank result on
the same global batch (data parallelism must not change the arithmetic beyond fp32 reassociation).
    python tools/worker_multigpu_check.py 2
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_content(root):
    from pathlib import Path
    from tests.test_contract import _tiny_model_dir
    root = Path(root)
    md, a, params = _tiny_model_dir(root)
    (root / "data").mkdir()
    rng = np.random.default_rng(0)
    with open(root / "data" / "train.jsonl", "w") as f:
        for _ in range(96):
            w = [f"w{i}" for i in rng.integers(0, 250, size=40)]
            f.write(json.dumps({"prompt": " ".join(w[:25]), "completion": " ".join(w[25:])}) + "\n")
    return params


def run(n_gpus, per_device):
    root = tempfile.mkdtemp(prefix=f"b200w_{n_gpus}gpu_")
    make_content(root)
    with open(os.path.join(root, "params.json"), "w") as f:
        json.dump({"max_steps": 3, "per_device_train_batch_size": per_device, "max_seq_length": 128,
                   "learning_rate": "1e-3", "save_steps": 0}, f)
    env = dict(os.environ, B200W_NUM_GPUS=str(n_gpus), PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, "-m", "runbooks_b200.worker", "train", "--content", root],
                       env=env, capture_output=True, text=True, timeout=600)
    print(f"--- {n_gpus} GPU(s): exit {p.returncode}")
    print("\n".join(l for l in p.stdout.splitlines() if '"loss"' in l or "event" in l)[-1500:])
    if p.returncode != 0:
        print(p.stderr[-3000:])
        sys.exit(1)
    from runbooks_b200 import contract
    w = dict(contract.iter_safetensors(os.path.join(root, "artifacts")))
    losses = [json.loads(l)["loss"] for l in p.stdout.splitlines() if l.startswith("{") and '"loss"' in l]
    shutil.rmtree(root, ignore_errors=True)
    return w, losses


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    w1, l1 = run(1, 2 * n)      # same global batch on one GPU
    wn, ln = run(n, 2)
    print("losses 1 GPU:", l1, f"\nlosses {n} GPU:", ln)
    import torch
    worst = 0.0
    for k in w1:
        a = torch.from_numpy(w1[k].astype(np.int32)).to(torch.int16).view(torch.bfloat16).float()
        b = torch.from_numpy(wn[k].astype(np.int32)).to(torch.int16).view(torch.bfloat16).float()
        worst = max(worst, float((a - b).norm() / a.norm()))
    print(f"worst relative weight difference 1 vs {n} GPUs: {worst:.3e}")
    # the logged loss is the global token mean on every rank count (HF average_tokens_across_devices)
    worst_loss = max(abs(a - b) for a, b in zip(l1, ln))
    print(f"worst |loss difference|: {worst_loss:.3e}")
    assert len(l1) == len(ln) and worst_loss < 2e-3, (l1, ln)
    assert worst < 2e-3, worst
    print("OK")
