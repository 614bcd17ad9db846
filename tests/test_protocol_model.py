"""The warp / mbarrier protocols of the attention kernels under adversarial schedules (CPU model,
tools/protocol_model.py). The model must FIND the bug round 1 shipped (dQ kernel, one "dS ready"
barrier for two TMEM stages -- wrong in 1.2 % of launches on hardware while every parity test
passed) and find nothing in the protocols attention.cu ships now. It is a model of the
synchronisation structure, not of the CUDA code: change one, change the other.
A clean result over random schedules is evidence, not proof."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import protocol_model as pm  # noqa: E402

TRIALS = 300


def _clean(kernel, **kw):
    for seed in (0, 1):
        ok, first, other = pm.explore(kernel, TRIALS, seed=seed, **kw)
        assert first is None, first
        assert not other, other
        assert ok == TRIALS


def _broken(kernel, needle, **kw):
    ok, first, other = pm.explore(kernel, TRIALS, seed=0, **kw)
    assert first is not None and needle in first, (ok, first, other)


def test_model_finds_the_v9_dq_race():
    _broken(pm.dq_kernel, "dQ MMA", njb=8, per_stage_bar_p=False)


def test_model_finds_the_same_hazard_in_the_v3_to_v8_dq_protocol():
    _broken(pm.dq_kernel_v8, "dQ MMA", njb=8)


def test_shipped_dq_protocol_is_clean():
    _clean(pm.dq_kernel, njb=8, per_stage_bar_p=True)
    _clean(pm.dq_kernel, njb=2, per_stage_bar_p=True)      # the shortest loop a CTA can have
    _clean(pm.dq_kernel, njb=3, per_stage_bar_p=True)


def test_round1_dkdv_protocol_had_no_stale_reads_but_an_aba_deadlock():
    """The single-bar_p protocol shipped in round 1 (no longer in attention.cu): never a stale read (the block-wide
    barrier keeps the compute warps together), but if the MMA warp is held up for a whole compute
    iteration the single bar_p flips twice and it waits for ever (attention.cu). The
    model must keep seeing that, or it has lost the sensitivity that makes its clean verdicts mean
    something."""
    deadlocks = 0
    for n_iter in (2, 3, 9):
        ok, first, other = pm.explore(pm.dkdv_kernel, 2000, seed=3, n_iter=n_iter)
        assert first is None, first                       # no stale data in any schedule
        assert all(k.startswith("deadlock") for k in other), other
        deadlocks += sum(other.values())
    assert deadlocks > 0
    _broken(pm.dkdv_kernel, "dV/dK MMA", n_iter=9, block_barrier=False)   # the block barrier is load-bearing


def test_shipped_dkdv_protocol_is_clean():
    """One bar_p per staging buffer: the only dK/dV protocol in attention.cu since round 2."""
    for n_iter in (2, 3, 4, 9):                              # fewer / more blocks than Q/dO buffers
        _clean(pm.dkdv_kernel, n_iter=n_iter, per_stage_bar_p=True)
    ok, first, other = pm.explore(pm.dkdv_kernel, 2000, seed=3, n_iter=2, per_stage_bar_p=True)
    assert (ok, first, other) == (2000, None, {})


def test_shipped_forward_protocol_is_clean_and_needs_its_bar_o_wait():
    for njb in (1, 2, 8):
        _clean(pm.fwd_kernel, njb=njb)
    _broken(pm.fwd_kernel, "PV MMA", njb=8, wait_bar_o=False)


def test_pair_gemm_ring_protocol_is_clean():
    """gemm_bf16_pair_kernel: a full/empty pair per pipeline stage and a tfull/tempty pair per accumulator
    stage -- per-stage barriers, so no ABA; checked anyway, including fewer k-blocks than stages, one
    tile, and the 8-arrival accumulator release collected from both CTAs."""
    for kw in (dict(num_tiles=5, num_kb=7), dict(num_tiles=1, num_kb=1), dict(num_tiles=3, num_kb=2, stages=6),
               dict(num_tiles=4, num_kb=13, stages=6)):
        ok, first, other = pm.explore(pm.gemm_pair_kernel, 300, seed=2, **kw)
        assert (ok, first, other) == (300, None, {}), (kw, first, other)
