"""The reference arm of bench.py as the driver launches it (`--impl reference`, alone and under torchrun with N
ranks): exactly ONE JSON line on stdout, from rank 0, exit code 0 on every rank, the base contract's keys plus
`impl`, `cpu_baseline` and a zero-copy `e2e`. Runs the real oracle-port sample, shrunk to 256 tokens through the
test-only variable so that the CPU suite stays short (the measured number is then not the workload's)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _check(stdout: str, n_gpus: int):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, stdout[:500]
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["cores"] >= 1 and "256-token" in d["cpu_baseline"]["sample"]
    assert "model" not in d["config"] and "workload" in d["config"]


@pytest.mark.parametrize("nproc", [1, 2])
def test_reference_arm_prints_one_line_from_rank_zero(nproc):
    env = dict(os.environ, B200W_BENCH_CPU_SAMPLE_TOKENS="256", CUDA_VISIBLE_DEVICES="")
    args = ["bench.py", "--impl", "reference", "--gpus", str(nproc), "--steps", "3", "--warmup", "1"]
    if nproc == 1:
        cmd = [sys.executable, *args]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", "29791", *args]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-1500:]
    _check(p.stdout, nproc)
