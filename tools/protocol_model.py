"""Executable model of the attention kernels' warp / mbarrier protocols (CPU, no GPU).

Why: round 1 shipped a dQ kernel whose lane quarters were coupled through ONE "dS ready" mbarrier
shared by two TMEM stages; a quarter running a block ahead completed the phase for a slower one and
the dQ MMA consumed stale data in 1.2 % of launches. Every parity test passed. The
protocol, not the arithmetic, was wrong -- and a protocol can be checked on a CPU by running it
under adversarial schedules with tagged buffer contents.

Model: agents are generators that yield blocking conditions; a random scheduler picks any runnable
agent, with arbitrary stalls. `Mbar` follows PTX mbarrier semantics (pending-arrival count, phase
bit, try_wait on parity). The tensor pipe is an in-order queue of issued ops (MMAs and commits)
that an extra agent retires one at a time, so `tcgen05.commit` arrives only after everything issued
before it has executed. Buffers carry tags ("what is in here"); every consumer asserts the tag it
needs. A run ends in OK, a Violation (stale / overwritten data) or a deadlock.

attention.cu's dQ kernel is modelled in its v9 (single bar_p) and v10 (bar_p per stage) forms, the
dK/dV and forward kernels in their shipped forms plus variants. tests/test_protocol_model.py requires
that the model FINDS the v9 bug. It also found something nobody had seen on hardware: the shipped
dK/dV protocol (one bar_p) is free of stale reads but not of an ABA deadlock -- if the MMA warp is
held up for a whole compute iteration right after issuing the next block's score MMAs, the compute
warps finish two blocks, bar_p flips twice and the MMA warp waits for a phase that is already gone.
Real warps do not stall like the model's scheduler, so this needs an external stall to happen
(DESIGN.md 7 discusses whether the 8-GPU failure was one); bar_p per stage removes it."""
import random


class Violation(Exception):
    pass


class Mbar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.pending, self.phase = self.count, self.phase ^ 1

    def passed(self, parity):          # mbarrier.try_wait.parity: true once the phase `parity` completed
        return self.phase != parity


class TensorPipe:
    """in-order execution of issued tcgen05 ops"""

    def __init__(self):
        self.q = []

    def issue(self, fn):
        self.q.append(fn)

    def commit(self, bar):
        self.q.append(bar.arrive)

    def agent(self, done):
        while True:
            if self.q:
                self.q.pop(0)()
                yield None
            elif done():
                return
            else:
                yield (lambda: bool(self.q) or done())


def run(agents, rng, max_steps=200000, stall_p=0.15):
    """agents: dict name -> generator. Generators yield None (runnable again) or a predicate to
    wait on. Random scheduling with random multi-step stalls."""
    waiting = {n: None for n in agents}
    stalled = {n: 0 for n in agents}
    live = dict(agents)
    for _ in range(max_steps):
        if not live:
            return "ok"
        runnable = [n for n in live if stalled[n] == 0 and (waiting[n] is None or waiting[n]())]
        any_stalled = any(stalled[n] for n in live)
        for n in live:
            if stalled[n]:
                stalled[n] -= 1
        if not runnable:
            if not any_stalled:      # nobody can move and nobody is merely descheduled
                return "deadlock: " + ", ".join(sorted(live))
            continue
        n = rng.choice(runnable)
        if rng.random() < stall_p:                      # this agent loses its SM sub-partition for a while
            stalled[n] = rng.randint(1, 40)
        try:
            waiting[n] = next(live[n])
        except StopIteration:
            del live[n]
    return "step limit"


# ---------------------------------------------------------------------------------------------
# dQ kernel (attention.cu attn_bwd_dq_kernel): MMA warp + 4 lane quarters (x hc collapsed) + tensor pipe
# ---------------------------------------------------------------------------------------------
def dq_kernel(njb, per_stage_bar_p, rng, quarters=4, threads_per_quarter=2):
    n_thr = quarters * threads_per_quarter
    pipe = TensorPipe()
    bar_s = [Mbar(1), Mbar(1)]
    bar_p = [Mbar(n_thr), Mbar(n_thr)] if per_stage_bar_p else [Mbar(n_thr)]
    bar_dq = Mbar(1)
    # tags: S[tb] = block whose scores are there; dS[tb][q][t] = block whose dS that thread wrote
    S = [None, None]
    dS = [[[None] * threads_per_quarter for _ in range(quarters)] for _ in range(2)]
    state = dict(mma_done=False, dq_blocks=[])

    def scores(j):
        def ex():
            S[j & 1] = j
            for q in range(quarters):                  # the dP MMA overwrites the columns dS lives in
                for t in range(threads_per_quarter):
                    dS[j & 1][q][t] = ("dP", j)
        return ex

    def dq_mma(j):
        def ex():
            for q in range(quarters):
                for t in range(threads_per_quarter):
                    if dS[j & 1][q][t] != ("dS", j):
                        raise Violation(f"dQ MMA of block {j} read {dS[j & 1][q][t]} from quarter {q}")
            state["dq_blocks"].append(j)
        return ex

    def mma_warp():
        pipe.issue(scores(0)); pipe.commit(bar_s[0])
        yield None
        for j in range(njb):
            if j + 1 < njb:
                pipe.issue(scores(j + 1)); pipe.commit(bar_s[(j + 1) & 1])
                yield None
            if per_stage_bar_p:
                b, par = bar_p[j & 1], (j >> 1) & 1
            else:
                b, par = bar_p[0], j & 1
            yield (lambda b=b, par=par: b.passed(par))
            pipe.issue(dq_mma(j))
            if j + 1 == njb:
                pipe.commit(bar_dq)
            yield None
        state["mma_done"] = True

    def compute(q, t):
        for j in range(njb):
            tb = j & 1
            yield (lambda tb=tb, j=j: bar_s[tb].passed((j >> 1) & 1))
            if S[tb] != j:
                raise Violation(f"quarter {q} read scores of block {S[tb]} while working on block {j}")
            yield None                                  # exp / dS math
            dS[tb][q][t] = ("dS", j)                    # tcgen05.st over the dP columns
            yield None
            (bar_p[tb] if per_stage_bar_p else bar_p[0]).arrive()
            yield None
        yield (lambda: bar_dq.passed(0))

    agents = {"mma": mma_warp(), "pipe": pipe.agent(lambda: state["mma_done"])}
    for q in range(quarters):
        for t in range(threads_per_quarter):
            agents[f"q{q}t{t}"] = compute(q, t)
    res = run(agents, rng)
    if res == "ok" and state["dq_blocks"] != list(range(njb)):
        raise Violation(f"dQ blocks executed: {state['dq_blocks']}")
    return res


# ---------------------------------------------------------------------------------------------
# dK/dV kernel (attn_bwd_dkdv_kernel): TMA warp (3 Q/dO buffers), MMA warp, compute warps with a
# block-wide barrier every iteration, staging tiles double-buffered behind bar_d
# ---------------------------------------------------------------------------------------------
def dkdv_kernel(n_iter, rng, n_thr=6, block_barrier=True, per_stage_bar_p=False):
    pipe = TensorPipe()
    bar_q = [Mbar(1) for _ in range(3)]
    bar_qfree = [Mbar(1) for _ in range(3)]
    bar_s, bar_d = [Mbar(1), Mbar(1)], [Mbar(1), Mbar(1)]
    bar_p = [Mbar(n_thr), Mbar(n_thr)] if per_stage_bar_p else [Mbar(n_thr)]
    Q = [None] * 3                       # which block's Q/dO is in buffer b
    S = [None, None]
    stage = [[None] * n_thr for _ in range(2)]     # P/dS staging tiles
    state = dict(mma_done=False, done_blocks=[], sync_count=0, sync_gen=0)

    def tma_warp():
        def load(it, buf):
            Q[buf] = it
            bar_q[buf].arrive()                        # complete_tx
        for it in range(min(3, n_iter)):
            load(it, it)
            yield None
        buf, par = 0, 0
        for it in range(n_iter - 3):
            yield (lambda buf=buf, par=par: bar_qfree[buf].passed(par))
            load(it + 3, buf)
            buf += 1
            if buf == 3:
                buf, par = 0, par ^ 1
            yield None

    def scores(it, qb):
        def ex():
            if Q[qb] != it:
                raise Violation(f"score MMA of block {it} read Q/dO of block {Q[qb]}")
            S[it & 1] = it
        return ex

    def dvdk(it, qb):
        def ex():
            if Q[qb] != it:
                raise Violation(f"dV/dK MMA of block {it} read Q/dO of block {Q[qb]}")
            for t in range(n_thr):
                if stage[it & 1][t] != it:
                    raise Violation(f"dV/dK MMA of block {it} read staging written for block {stage[it & 1][t]} (thread {t})")
            state["done_blocks"].append(it)
        return ex

    def mma_warp():
        yield (lambda: bar_q[0].passed(0))
        pipe.issue(scores(0, 0)); pipe.commit(bar_s[0])
        qb, qpar = 0, 0
        for it in range(n_iter):
            nqb, npar = qb + 1, qpar
            if nqb == 3:
                nqb, npar = 0, npar ^ 1
            if it + 1 < n_iter:
                yield (lambda nqb=nqb, npar=npar: bar_q[nqb].passed(npar))
                pipe.issue(scores(it + 1, nqb)); pipe.commit(bar_s[(it + 1) & 1])
            if per_stage_bar_p:
                b, par = bar_p[it & 1], (it >> 1) & 1
            else:
                b, par = bar_p[0], it & 1
            yield (lambda b=b, par=par: b.passed(par))
            pipe.issue(dvdk(it, qb)); pipe.commit(bar_d[it & 1]); pipe.commit(bar_qfree[qb])
            qb, qpar = nqb, npar
            yield None
        state["mma_done"] = True

    def compute(t):
        for it in range(n_iter):
            tb = it & 1
            yield (lambda tb=tb, it=it: bar_s[tb].passed((it >> 1) & 1))
            if S[tb] != it:
                raise Violation(f"thread {t} read scores of block {S[tb]} while working on block {it}")
            if it >= 2:
                yield (lambda tb=tb, it=it: bar_d[tb].passed(((it >> 1) - 1) & 1))
            yield None
            stage[tb][t] = it
            yield None
            (bar_p[tb] if per_stage_bar_p else bar_p[0]).arrive()
            if block_barrier:                            # bwd_compute_bar_sync()
                gen = state["sync_gen"]
                state["sync_count"] += 1
                if state["sync_count"] == n_thr:
                    state["sync_count"], state["sync_gen"] = 0, gen + 1
                yield (lambda gen=gen: state["sync_gen"] != gen)
            else:
                yield None
        yield (lambda: bar_d[(n_iter - 1) & 1].passed(((n_iter - 1) >> 1) & 1))

    agents = {"tma": tma_warp(), "mma": mma_warp(), "pipe": pipe.agent(lambda: state["mma_done"])}
    for t in range(n_thr):
        agents[f"c{t}"] = compute(t)
    res = run(agents, rng)
    if res == "ok" and state["done_blocks"] != list(range(n_iter)):
        raise Violation(f"blocks executed: {state['done_blocks']}")
    return res


# ---------------------------------------------------------------------------------------------
# dQ kernel as it was in builds v3-v8: dS goes to a double-buffered SHARED-MEMORY staging tile, a
# thread waits for bar_dq[tb] (dQ MMA of block j-2 retired) before overwriting it; one bar_p.
# ---------------------------------------------------------------------------------------------
def dq_kernel_v8(njb, rng, n_thr=8):
    pipe = TensorPipe()
    bar_s, bar_dq = [Mbar(1), Mbar(1)], [Mbar(1), Mbar(1)]
    bar_p = Mbar(n_thr)
    S = [None, None]
    stage = [[None] * n_thr for _ in range(2)]
    state = dict(mma_done=False, blocks=[])

    def scores(j):
        def ex():
            S[j & 1] = j
        return ex

    def dq_mma(j):
        def ex():
            for t in range(n_thr):
                if stage[j & 1][t] != j:
                    raise Violation(f"dQ MMA of block {j} read staging written for block {stage[j & 1][t]} (thread {t})")
            state["blocks"].append(j)
        return ex

    def mma_warp():
        pipe.issue(scores(0)); pipe.commit(bar_s[0])
        yield None
        for j in range(njb):
            if j + 1 < njb:
                pipe.issue(scores(j + 1)); pipe.commit(bar_s[(j + 1) & 1])
                yield None
            yield (lambda j=j: bar_p.passed(j & 1))
            pipe.issue(dq_mma(j)); pipe.commit(bar_dq[j & 1])
            yield None
        state["mma_done"] = True

    def compute(t):
        for j in range(njb):
            tb = j & 1
            yield (lambda tb=tb, j=j: bar_s[tb].passed((j >> 1) & 1))
            if S[tb] != j:
                raise Violation(f"thread {t} read scores of block {S[tb]} while working on block {j}")
            if j >= 2:
                yield (lambda tb=tb, j=j: bar_dq[tb].passed(((j >> 1) - 1) & 1))
            yield None
            stage[tb][t] = j
            yield None
            bar_p.arrive()
            yield None
        yield (lambda: bar_dq[(njb - 1) & 1].passed(((njb - 1) >> 1) & 1))

    agents = {"mma": mma_warp(), "pipe": pipe.agent(lambda: state["mma_done"])}
    for t in range(n_thr):
        agents[f"c{t}"] = compute(t)
    res = run(agents, rng)
    if res == "ok" and state["blocks"] != list(range(njb)):
        raise Violation(f"blocks executed: {state['blocks']}")
    return res


# ---------------------------------------------------------------------------------------------
# forward kernel (attn_fwd_kernel): S double-buffered in TMEM, ONE P staging tile and O in TMEM,
# both released by bar_o (PV of the previous block retired); one bar_p. V double-buffered by the
# TMA warp behind bar_vfree (not modelled: it has the dK/dV kernel's Q/dO ring structure).
# ---------------------------------------------------------------------------------------------
def fwd_kernel(njb, rng, n_thr=8, wait_bar_o=True):
    pipe = TensorPipe()
    bar_s = [Mbar(1), Mbar(1)]
    bar_p, bar_o = Mbar(n_thr), Mbar(1)
    S = [None, None]
    P = [None] * n_thr
    state = dict(mma_done=False, blocks=[])

    def scores(j):
        def ex():
            S[j & 1] = j
        return ex

    def pv(j):
        def ex():
            for t in range(n_thr):
                if P[t] != j:
                    raise Violation(f"PV MMA of block {j} read P written for block {P[t]} (thread {t})")
            state["blocks"].append(j)
        return ex

    def mma_warp():
        pipe.issue(scores(0)); pipe.commit(bar_s[0])
        yield None
        for j in range(njb):
            if j + 1 < njb:
                pipe.issue(scores(j + 1)); pipe.commit(bar_s[(j + 1) & 1])
                yield None
            yield (lambda j=j: bar_p.passed(j & 1))
            pipe.issue(pv(j)); pipe.commit(bar_o)
            yield None
        state["mma_done"] = True

    def compute(t):
        for j in range(njb):
            tb = j & 1
            yield (lambda tb=tb, j=j: bar_s[tb].passed((j >> 1) & 1))
            if S[tb] != j:
                raise Violation(f"thread {t} read scores of block {S[tb]} while working on block {j}")
            yield None                                   # softmax math
            if j > 0 and wait_bar_o:
                yield (lambda j=j: bar_o.passed((j - 1) & 1))   # PV(j-1) retired: P tile and O are free
            P[t] = j
            yield None
            bar_p.arrive()
            yield None
        yield (lambda: bar_o.passed((njb - 1) & 1))

    agents = {"mma": mma_warp(), "pipe": pipe.agent(lambda: state["mma_done"])}
    for t in range(n_thr):
        agents[f"c{t}"] = compute(t)
    res = run(agents, rng)
    if res == "ok" and state["blocks"] != list(range(njb)):
        raise Violation(f"blocks executed: {state['blocks']}")
    return res


# ---------------------------------------------------------------------------------------------
# CTA-pair persistent GEMM (gemm.cu gemm_bf16_pair_kernel): TMA producers in both CTAs fill a ring of
# STAGES slots (one full / empty barrier pair per slot, the leader's full barrier collects both CTAs'
# bytes), the leader's MMA warp consumes them into one of two TMEM accumulator stages and commits
# empty (multicast to both CTAs) / tfull (multicast); 4 epilogue warps per CTA drain the
# accumulator and arrive on the LEADER's tempty barrier (count 8).
# ---------------------------------------------------------------------------------------------
def gemm_pair_kernel(num_tiles, num_kb, rng, stages=3, epi_warps=4):
    pipe = TensorPipe()
    full = [Mbar(2) for _ in range(stages)]               # leader's: one complete_tx per CTA (modelled as 2 arrivals)
    empty = [[Mbar(1) for _ in range(stages)] for _ in range(2)]   # per CTA (the commit is multicast)
    tfull = [[Mbar(1), Mbar(1)] for _ in range(2)]        # per CTA (multicast)
    tempty = [Mbar(2 * epi_warps), Mbar(2 * epi_warps)]   # leader's
    slot = [[None] * stages for _ in range(2)]            # per CTA: (tile, kb) whose operand halves are in the slot
    acc = [None, None]                                    # which tile's sum is in accumulator stage a: (tile, kb_count)
    state = dict(mma_done=False, out=[[], []])

    def producer(cta):
        stage, phase = 0, 0
        for tile in range(num_tiles):
            for kb in range(num_kb):
                yield (lambda s=stage, ph=phase: empty[cta][s].passed(ph ^ 1))
                slot[cta][stage] = (tile, kb)
                full[stage].arrive()
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield None

    def mma(tile, kb, stage, a):
        def ex():
            for cta in range(2):
                if slot[cta][stage] != (tile, kb):
                    raise Violation(f"MMA of tile {tile} k-block {kb} read slot holding {slot[cta][stage]} (CTA {cta})")
            if kb == 0:
                acc[a] = [tile, 1]
            else:
                if acc[a][0] != tile:
                    raise Violation(f"MMA of tile {tile} accumulated onto tile {acc[a][0]}")
                acc[a][1] += 1
        return ex

    def commit_multicast(bars):
        def ex():
            for b in bars:
                b.arrive()
        return ex

    def mma_warp():
        stage, phase, a, aph = 0, 0, 0, 0
        for tile in range(num_tiles):
            yield (lambda a=a, aph=aph: tempty[a].passed(aph ^ 1))
            for kb in range(num_kb):
                yield (lambda s=stage, ph=phase: full[s].passed(ph))
                pipe.issue(mma(tile, kb, stage, a))
                pipe.issue(commit_multicast([empty[0][stage], empty[1][stage]]))
                stage += 1
                if stage == stages:
                    stage, phase = 0, phase ^ 1
                yield None
            pipe.issue(commit_multicast([tfull[0][a], tfull[1][a]]))
            a += 1
            if a == 2:
                a, aph = 0, aph ^ 1
            yield None
        state["mma_done"] = True

    def epilogue(cta, w):
        a, aph = 0, 0
        for tile in range(num_tiles):
            yield (lambda a=a, aph=aph: tfull[cta][a].passed(aph))
            if acc[a] != [tile, num_kb]:
                raise Violation(f"epilogue of tile {tile} read accumulator holding {acc[a]}")
            yield None                                   # tcgen05.ld + stores
            if w == 0:
                state["out"][cta].append(tile)
            tempty[a].arrive()
            a += 1
            if a == 2:
                a, aph = 0, aph ^ 1
            yield None

    agents = {"mma": mma_warp(), "pipe": pipe.agent(lambda: state["mma_done"]), "tma0": producer(0), "tma1": producer(1)}
    for cta in range(2):
        for w in range(epi_warps):
            agents[f"epi{cta}{w}"] = epilogue(cta, w)
    res = run(agents, rng)
    if res == "ok" and state["out"] != [list(range(num_tiles))] * 2:
        raise Violation(f"tiles written: {state['out']}")
    return res


def explore(kernel, trials, seed=0, **kw):
    """-> (#ok, first violation or None, other outcomes)"""
    ok, first, other = 0, None, {}
    for i in range(trials):
        rng = random.Random(seed * 100003 + i)
        try:
            r = kernel(rng=rng, **kw)
        except Violation as v:
            first = first or f"trial {i}: {v}"
            continue
        if r == "ok":
            ok += 1
        else:
            other[r] = other.get(r, 0) + 1
    return ok, first, other


if __name__ == "__main__":
    for name, fn, kw in [("dQ v9 (one bar_p)", dq_kernel, dict(njb=8, per_stage_bar_p=False)),
                         ("dQ v10 (bar_p per stage)", dq_kernel, dict(njb=8, per_stage_bar_p=True)),
                         ("dQ v3-v8 (smem staging, one bar_p)", dq_kernel_v8, dict(njb=8)),
                         ("forward shipped", fwd_kernel, dict(njb=8)),
                         ("forward without the bar_o wait", fwd_kernel, dict(njb=8, wait_bar_o=False)),
                         ("dK/dV shipped (one bar_p)", dkdv_kernel, dict(n_iter=2)),
                         ("dK/dV with bar_p per stage", dkdv_kernel, dict(n_iter=2, per_stage_bar_p=True)),
                         ("dK/dV per stage, no block barrier", dkdv_kernel, dict(n_iter=9, per_stage_bar_p=True, block_barrier=False)),
                         ("dK/dV without the block barrier", dkdv_kernel, dict(n_iter=9, block_barrier=False)),
                         ("pair GEMM, 5 tiles x 7 k-blocks", gemm_pair_kernel, dict(num_tiles=5, num_kb=7)),
                         ("pair GEMM, 1 tile x 1 k-block", gemm_pair_kernel, dict(num_tiles=1, num_kb=1))]:
        ok, first, other = explore(fn, 2000, **kw)
        print(f"{name:36s} ok {ok:4d}/2000   first violation: {first}   other: {other}")
