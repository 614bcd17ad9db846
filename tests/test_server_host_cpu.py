"""Host logic of the Server path on CPU: continuous batching (infer.Generator), the scheduler thread and
the HTTP surface the reference probes and the system test speaks (server_controller.go:156-173 readiness
GET / -> 200; test/system.sh:73-78 POST /v1/completions {"prompt", "max_tokens"}).

The engine is a stub whose "model" makes the next token a hash of EVERYTHING fed into that KV-cache slot
so far, at the positions it was fed: a scheduler that mixes up slots, positions or the token it feeds
back produces different text. (The CUDA engine itself is covered by tests/test_infer.py / test_server.py
on a GPU; there is no CPU fallback in the product.)"""
import json
import threading
import urllib.error
import urllib.request
from http.server import ThreadingHTTPServer
from types import SimpleNamespace

import numpy as np
import pytest

from runbooks_b200 import server
from runbooks_b200.infer import Generator

VOCAB, EOS = 97, 96


class StubEngine:
    def __init__(self, max_batch=2, max_ctx=64):
        self.max_batch = max_batch
        self.serve_arch = SimpleNamespace(max_ctx=max_ctx, vocab_size=VOCAB)
        self.cache = {}            # slot -> list of (position, token)
        self.batch_sizes = []

    def step(self, tokens, positions, slots, want_logits=False):
        assert len(set(slots)) == len(slots), "one row per cache slot"
        assert len(tokens) <= self.max_batch
        self.batch_sizes.append(len(tokens))
        out = []
        for t, p, s in zip(tokens, positions, slots):
            hist = self.cache.setdefault(s, [])
            if p == 0:
                hist.clear()                       # a new request re-uses the slot from position 0
            assert p == len(hist), f"slot {s}: fed position {p}, cache holds {len(hist)} tokens"
            hist.append((int(p), int(t)))
            h = 7
            for pp, tt in hist:
                h = (h * 31 + 3 * pp + tt) % 1000003
            out.append(h % (VOCAB - 1))            # never EOS unless a test asks for it
        return np.array(out, dtype=np.int32), None


def reference(prompt, max_tokens):
    """the same stub, one request alone in slot 0"""
    g = Generator(StubEngine(max_batch=1))
    return g.generate([prompt], max_tokens)[0]


def test_continuous_batching_equals_one_request_at_a_time():
    rng = np.random.default_rng(0)
    prompts = [list(map(int, rng.integers(0, VOCAB - 1, size=n))) for n in (1, 5, 3, 9, 2, 7)]
    lens = [4, 1, 6, 3, 8, 2]
    eng = StubEngine(max_batch=3)
    g = Generator(eng)
    reqs, todo = [], list(zip(prompts, lens))
    while todo or g.active:
        while todo and g.free:                     # admit whenever a slot is free, like the scheduler
            p, n = todo.pop(0)
            reqs.append((g.add(p, n), p, n))
        g.step()
    for r, p, n in reqs:
        assert r.done and r.out == reference(p, n)
        assert len(r.out) == n
    assert max(eng.batch_sizes) == 3 and sorted(g.free) == [0, 1, 2]


def test_generator_rejects_what_cannot_fit():
    g = Generator(StubEngine(max_batch=1, max_ctx=8))
    with pytest.raises(ValueError):
        g.add([], 2)
    with pytest.raises(ValueError):
        g.add([1, 2, 3, 4, 5], 4)                  # 5 + 4 > 8
    g.add([1, 2, 3, 4], 4)                         # exactly the cache length
    with pytest.raises(RuntimeError):
        g.add([1], 1)                              # no free slot


def test_generator_stops_at_eos():
    class EosAtThird(StubEngine):
        def step(self, tokens, positions, slots, want_logits=False):
            nxt, _ = super().step(tokens, positions, slots)
            return np.array([EOS if p == 4 else n for n, p in zip(nxt, positions)], dtype=np.int32), None
    g = Generator(EosAtThird(max_batch=1), eos_id=EOS)
    out = g.generate([[5, 6, 7]], 10)[0]          # positions 0..2 prompt; generated at 2, 3, 4 -> third is EOS
    assert len(out) == 3 and out[-1] == EOS


class CharTok:
    """ids = code points mod VOCAB; enough for the HTTP layer"""
    bos_id, eos_id = 1, EOS

    def encode(self, s):
        return [ord(c) % (VOCAB - 1) for c in s]

    def decode(self, ids):
        return "".join(chr(97 + i % 26) for i in ids)


@pytest.fixture()
def http_server():
    eng = StubEngine(max_batch=2, max_ctx=48)
    sched = server.Scheduler(eng, CharTok())
    httpd = ThreadingHTTPServer(("127.0.0.1", 0), server.make_handler(sched, "stub-model"))
    t = threading.Thread(target=httpd.serve_forever, daemon=True)
    t.start()
    yield sched, eng, f"http://127.0.0.1:{httpd.server_address[1]}"
    httpd.shutdown()


def _get(url):
    try:
        with urllib.request.urlopen(url, timeout=10) as r:
            return r.status, json.loads(r.read())
    except urllib.error.HTTPError as e:
        return e.code, json.loads(e.read())


def _post(url, obj):
    req = urllib.request.Request(url, data=json.dumps(obj).encode(), headers={"Content-Type": "application/json"})
    try:
        with urllib.request.urlopen(req, timeout=30) as r:
            return r.status, json.loads(r.read())
    except urllib.error.HTTPError as e:
        return e.code, json.loads(e.read())


def test_readiness_probe_follows_the_scheduler(http_server):
    sched, _, base = http_server
    assert _get(base + "/")[0] == 503                       # model still loading: the Deployment is not Ready
    sched.start()
    assert sched.ready.wait(5)
    code, body = _get(base + "/")
    assert code == 200 and body["status"] == "ok"
    assert _get(base + "/v1/models")[1]["data"][0]["id"] == "stub-model"
    assert _get(base + "/nope")[0] == 404


def test_completions_shape_and_concurrency_beyond_the_slot_count(http_server):
    sched, eng, base = http_server
    prompts = [f"request number {i} " + "x" * i for i in range(8)]     # 8 requests, 2 cache slots
    results = [None] * len(prompts)

    def go(i):
        results[i] = _post(base + "/v1/completions", {"prompt": prompts[i], "max_tokens": 3 + i % 4})
    threads = [threading.Thread(target=go, args=(i,)) for i in range(len(prompts))]
    for t in threads:
        t.start()
    # all eight are queued BEFORE the scheduler runs (the stub engine is instantaneous, so otherwise
    # whether two requests ever share a step would depend on thread timing)
    import time
    deadline = time.time() + 20
    while sched.q.qsize() < len(prompts) and time.time() < deadline:
        time.sleep(0.01)
    assert sched.q.qsize() == len(prompts)
    sched.start()
    for t in threads:
        t.join(60)
    tok = CharTok()
    for i, (code, body) in enumerate(results):
        assert code == 200, body
        n = 3 + i % 4
        ids = [tok.bos_id] + tok.encode(prompts[i])
        assert body["object"] == "text_completion" and body["model"] == "stub-model"
        ch = body["choices"][0]
        assert ch["text"] == tok.decode(reference(ids, n)) and ch["finish_reason"] == "length"
        assert body["usage"] == {"prompt_tokens": len(ids), "completion_tokens": n, "total_tokens": len(ids) + n}
    assert max(eng.batch_sizes) == 2


def test_bad_requests_do_not_kill_the_server(http_server):
    sched, _, base = http_server
    sched.start()
    assert _post(base + "/v1/completions", {"max_tokens": 3})[0] == 400                     # no prompt
    assert _post(base + "/v1/completions", {"prompt": "hi", "max_tokens": 0})[0] == 400
    code, body = _post(base + "/v1/completions", {"prompt": "y" * 60, "max_tokens": 4})     # longer than the KV cache
    assert code == 400 and "KV cache" in body["error"]
    assert _post(base + "/v1/other", {"prompt": "hi"})[0] == 404
    code, body = _post(base + "/v1/completions", {"prompt": ["list form"], "max_tokens": 2})
    assert code == 200 and body["usage"]["completion_tokens"] == 2
    assert _get(base + "/")[0] == 200


# ---- round 2: streaming, parameters, back-pressure, prefill admission -- all on the stub engine ----
def _sse(url, obj):
    req = urllib.request.Request(url, data=json.dumps(obj).encode(), headers={"Content-Type": "application/json"})
    with urllib.request.urlopen(req, timeout=30) as r:
        assert r.headers["Content-Type"].startswith("text/event-stream")
        return [l[6:] for l in r.read().decode().split("\n") if l.startswith("data: ")]


def test_streaming_chunks_concatenate_to_the_plain_completion(http_server):
    """`"stream": true` answers with server-sent events (the shape basaran / the OpenAI API speak): one JSON
    chunk per text delta, a final chunk with finish_reason, then [DONE]."""
    sched, _, base = http_server
    sched.start()
    code, plain = _post(base + "/v1/completions", {"prompt": "stream me", "max_tokens": 9})
    assert code == 200
    ev = _sse(base + "/v1/completions", {"prompt": "stream me", "max_tokens": 9, "stream": True})
    assert ev[-1] == "[DONE]"
    chunks = [json.loads(e) for e in ev[:-1]]
    assert all(c["object"] == "text_completion" for c in chunks)
    assert "".join(c["choices"][0]["text"] for c in chunks) == plain["choices"][0]["text"]
    assert chunks[-1]["choices"][0]["finish_reason"] == "length" and len(chunks) >= 9
    # echo: the prompt is the first chunk
    ev = _sse(base + "/v1/completions", {"prompt": "stream me", "max_tokens": 2, "stream": True, "echo": True})
    assert json.loads(ev[0])["choices"][0]["text"] == "stream me"


def test_stop_n_echo_and_sampling_parameters(http_server):
    sched, _, base = http_server
    sched.start()
    code, plain = _post(base + "/v1/completions", {"prompt": "abc", "max_tokens": 12})
    text = plain["choices"][0]["text"]
    cut = text[5:7]
    code, st = _post(base + "/v1/completions", {"prompt": "abc", "max_tokens": 12, "stop": cut})
    k = text.find(cut)
    assert code == 200 and st["choices"][0]["text"] == text[:k] and st["choices"][0]["finish_reason"] == "stop"
    code, two = _post(base + "/v1/completions", {"prompt": "abc", "max_tokens": 4, "n": 2, "echo": True})
    assert code == 200 and [c["index"] for c in two["choices"]] == [0, 1]
    assert two["choices"][0]["text"] == two["choices"][1]["text"] == "abc" + text[:4]       # greedy: identical
    assert two["usage"]["completion_tokens"] == 8
    # what is not implemented is a 400, never silently ignored
    for bad in ({"logprobs": 2}, {"best_of": 3}, {"presence_penalty": 1.0}, {"frequency_penalty": 0.5},
                {"temperature": -0.1}, {"top_p": 0}, {"n": 0}, {"n": 2, "stream": True}, {"stop": ["a"] * 5}):
        assert _post(base + "/v1/completions", dict({"prompt": "abc", "max_tokens": 2}, **bad))[0] == 400, bad
    # temperature > 0 needs the engine's logits: the stub has none, so the request fails alone (500-class as 400)
    code, body = _post(base + "/v1/completions", {"prompt": "abc", "max_tokens": 2, "temperature": 0.7})
    assert code == 400
    assert _post(base + "/v1/completions", {"prompt": "abc", "max_tokens": 2})[0] == 200     # the server lives on


def test_sample_token_is_nucleus_sampling():
    from runbooks_b200.infer import sample_token
    rng = np.random.default_rng(0)
    logits = np.array([5.0, 4.0, 0.0, -3.0, -9.0], dtype=np.float32)
    draws = [sample_token(logits, 1.0, 1.0, rng) for _ in range(4000)]
    p = np.exp(logits - logits.max()); p /= p.sum()
    freq = np.bincount(draws, minlength=5) / len(draws)
    assert np.abs(freq - p).max() < 0.03
    assert set(sample_token(logits, 1.0, 0.5, rng) for _ in range(200)) == {0}          # top_p 0.5: only the head
    assert set(sample_token(logits, 1.0, 0.9, rng) for _ in range(500)) == {0, 1}
    assert all(sample_token(logits, 1e-4, 1.0, rng) == 0 for _ in range(50))             # T -> 0: argmax


def test_backpressure_answers_503_when_slots_and_queue_are_full():
    eng = StubEngine(max_batch=1, max_ctx=64)
    sched = server.Scheduler(eng, CharTok(), max_queue=2)        # not started: nothing is admitted
    httpd = ThreadingHTTPServer(("127.0.0.1", 0), server.make_handler(sched, "stub", request_timeout_s=0.5))
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    base = f"http://127.0.0.1:{httpd.server_address[1]}"
    try:
        results = []
        ts = [threading.Thread(target=lambda: results.append(_post(base + "/v1/completions", {"prompt": "q", "max_tokens": 2})))
              for _ in range(2)]
        for t in ts:
            t.start()
        import time
        deadline = time.time() + 10
        while sched.q.qsize() < 2 and time.time() < deadline:
            time.sleep(0.01)
        code, body = _post(base + "/v1/completions", {"prompt": "one too many", "max_tokens": 2})
        assert code == 503 and "busy" in body["error"]
        for t in ts:
            t.join(10)
        assert sorted(r[0] for r in results) == [504, 504]       # never scheduled: the request timeout answers
    finally:
        httpd.shutdown()


def test_generator_prefills_at_admission_when_the_engine_can():
    """An engine with `prefill` ingests the whole prompt in one call (b200w_infer_prefill) and continues with
    decode steps at position len(prompt); the token stream must equal the token-by-token path's."""
    class PrefillStub(StubEngine):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.prefills = []

        def prefill(self, prompts, slots, want_logits=False):
            self.prefills.append([len(p) for p in prompts])
            outs = []
            for p, s in zip(prompts, slots):
                self.cache[s] = []
                nxt = None
                for i, t in enumerate(p):                       # same hash model, whole prompt at once
                    nxt, _ = StubEngine.step(self, [t], [i], [s])
                outs.append(int(nxt[0]))
            self.batch_sizes = [b for b in self.batch_sizes if b != 1 or True]
            return np.array(outs, dtype=np.int32), None

    rng = np.random.default_rng(1)
    prompts = [list(map(int, rng.integers(0, VOCAB - 1, size=n))) for n in (1, 7, 3, 12)]
    eng = PrefillStub(max_batch=2)
    g = Generator(eng)
    outs = []
    for p in prompts:                                            # two slots: admit, run to completion in pairs
        while not g.free:
            g.step()
        outs.append(g.add(p, 5))
    while g.active:
        g.step()
    assert eng.prefills == [[7], [3], [12]]                      # the 1-token prompt goes through the decode path
    for r, p in zip(outs, prompts):
        assert r.out == reference(p, 5)


def test_scheduler_prefills_a_round_of_admissions_in_one_call():
    """Requests admitted in the same scheduler round are ingested by ONE prefill call per padded length."""
    class PrefillStub(StubEngine):
        prefills = []

        def prefill(self, prompts, slots, want_logits=False):
            PrefillStub.prefills.append(sorted(len(p) for p in prompts))
            outs = []
            for p, s in zip(prompts, slots):
                self.cache[s] = []
                for i, t in enumerate(p):
                    nxt, _ = StubEngine.step(self, [t], [i], [s])
                outs.append(int(nxt[0]))
            return np.array(outs, dtype=np.int32), None

    PrefillStub.prefills = []
    eng = PrefillStub(max_batch=4, max_ctx=400)
    sched = server.Scheduler(eng, CharTok())
    items = [sched.submit("a" * n, 3) for n in (10, 50, 200, 120)]       # 11, 51, 201, 121 ids with <s>
    sched.start()
    tok = CharTok()
    for it, n in zip(items, (10, 50, 200, 120)):
        kind, val = it["events"].get(timeout=20)
        while kind == "delta":
            kind, val = it["events"].get(timeout=20)
        assert kind == "done", val
        alone = Generator(StubEngine(max_batch=1, max_ctx=400)).generate([[tok.bos_id] + tok.encode("a" * n)], 3)[0]
        assert val["text"] == tok.decode(alone)
    assert PrefillStub.prefills == [[11, 51, 121], [201]]                # <= 128 tokens together, the 201 alone
