"""One FULL fine-tune step (Llama-2-7B, per-device batch 8 x 4096 tokens as MICRO_BATCH-sequence micro-steps --
default 2, what bench.py runs --, clip, AdamW) for an ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/r2_step_launches.csv python tools/step_timeline.py

Two warm-up steps run outside the profiled range (cudaProfilerStart/Stop bracket exactly one step), so the
capture costs ~6.7k serialised launches instead of bench.py's warm-up + timed regions. The un-profiled step
time of the same process is printed for the share-of-step comparison (profiles/r02_step_timeline.txt)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from runbooks_b200.engine import Engine, LlamaArch  # noqa: E402

layers = int(os.environ.get("LAYERS", "32"))
arch = LlamaArch.llama2_7b(4096)
arch.num_layers = layers
e = Engine(0)
micro = int(os.environ.get("MICRO_BATCH", "2"))
e.init_model(arch, micro_batch=micro, training=True)
e.init_random(0, 0.02)
S, nseq = 4096, 8
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, arch.vocab_size, (4, nseq, S), generator=g, dtype=torch.int32).cuda()
for i in range(2):
    e.train_step_resident(ids[i].data_ptr(), ids[i].data_ptr(), nseq, nseq * S, lr=5e-5)
e.sync()
e.timer_start()
e.train_step_resident(ids[2].data_ptr(), ids[2].data_ptr(), nseq, nseq * S, lr=5e-5)
ms = e.timer_stop()
l0 = e.launch_count()
torch.cuda.synchronize()
torch.cuda.profiler.start()
e.train_step_resident(ids[3].data_ptr(), ids[3].data_ptr(), nseq, nseq * S, lr=5e-5)
e.sync()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
loss, gn = e.read_scalars()
print(f"STEP_TIMELINE layers={layers} unprofiled_ms_per_step={ms:.2f} launches_in_profiled_step={e.launch_count() - l0} "
      f"loss={loss:.4f} grad_norm={gn:.4f}", flush=True)
e.close()
