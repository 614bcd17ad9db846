"""Falcon-7B greedy decode throughput at batch 32 (BASELINE.json configs[3]) against the HBM
roofline of SURVEY.md §8d: bytes per step = bf16 weights + KV read, peak from MEASURED_PEAKS.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from runbooks_b200.infer import InferEngine, ServeArch  # noqa: E402

B, CTX, STEPS = 32, 1024, 64
arch = ServeArch.falcon_7b(max_ctx=CTX + STEPS + 8)
e = InferEngine(0)
e.init_infer(arch, max_batch=B)
e.infer_init_random(0, 0.02)
rng = np.random.default_rng(0)
slots = list(range(B))
# fill the cache up to CTX with arbitrary tokens at sparse positions is not valid: positions must be
# contiguous per slot, so ingest a short prompt and then time decode steps at a given depth by
# starting positions at CTX (the attention kernel reads [0, pos] whatever was written there)
tok = rng.integers(0, arch.vocab_size, size=B)
for w in range(3):
    tok, _ = e.step(tok, [CTX + w] * B, slots)
e.sync()
e.timer_start()
t0 = time.perf_counter()
for i in range(STEPS):
    tok, _ = e.step(tok, [CTX + 3 + i] * B, slots)
ms = e.timer_stop()
wall = (time.perf_counter() - t0) * 1e3
n_params = sum(int(np.prod(s)) for _, s in e.infer_params())
bytes_step = 2 * n_params + B * (CTX + STEPS // 2) * arch.num_layers * 2 * arch.num_kv_heads * arch.head_dim * 2
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
per = max(ms, wall) / STEPS
print(json.dumps({"metric": "Falcon-7B greedy decode tokens/s, batch 32, ctx ~1024, 1xB200",
                  "value": round(B / (per / 1e3), 1), "ms_per_step": round(per, 3), "device_ms_per_step": round(ms / STEPS, 3),
                  "params": n_params, "bytes_per_step": bytes_step,
                  "roofline": {"bound": "hbm", "achieved": round(bytes_step / (per / 1e3) / 1e9, 1), "peak": peak,
                               "unit": "GB/s", "frac": round(bytes_step / (per / 1e3) / 1e9 / peak, 4)}}))
