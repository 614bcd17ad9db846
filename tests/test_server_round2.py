"""HTTP surface of the Server on a GPU with the reference's system-test model family (OPT): what
test/system.sh:46-78 does -- load the model directory, wait for readiness on GET /, POST /v1/completions --
plus the streaming shape basaran speaks (SSE), sampling parameters and back-pressure."""
import json
import os
import threading
import urllib.error
import urllib.request
from http.server import ThreadingHTTPServer

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _opt_model_dir(tmp_path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from oracle import opt_oracle as OO
    from runbooks_b200 import contract
    from runbooks_b200.engine import OptArch
    from util import bf16_bits

    oa = OO.OptArch(256, 128, 256, 2, 2, 256)
    params = OO.seeded_params(oa, 8)
    md = tmp_path / "model"
    md.mkdir()
    vocab = {"<s>": 0, "<pad>": 1, "</s>": 2, "<unk>": 3, **{f"w{i}": i + 4 for i in range(252)}}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    tok.save(str(md / "tokenizer.json"))
    (md / "tokenizer_config.json").write_text(json.dumps({"bos_token": "</s>", "eos_token": "</s>", "pad_token": "<pad>"}))
    arch = OptArch(oa.vocab_size, oa.hidden_size, oa.ffn_dim, oa.num_layers, oa.num_heads, max_positions=256,
                   max_seq_len=128)
    contract.save_hf_checkpoint(str(md), arch.to_hf_config(), ((k, bf16_bits(v)) for k, v in params.items()))
    return md, oa, params


def _post(url, body, timeout=120):
    req = urllib.request.Request(url, data=json.dumps(body).encode(), headers={"Content-Type": "application/json"})
    return urllib.request.urlopen(req, timeout=timeout)


def test_opt_server_completions_streaming_and_parameters(tmp_path):
    from runbooks_b200 import contract, server
    md, oa, params = _opt_model_dir(tmp_path)
    engine, cfg = server.load_engine(str(md), max_batch=2, max_ctx=256)
    assert cfg["model_type"] == "opt"
    tok = contract.Tokenizer(str(md))
    sched = server.Scheduler(engine, tok, max_queue=2)
    sched.start()
    httpd = ThreadingHTTPServer(("127.0.0.1", 0), server.make_handler(sched, "opt-tiny"))
    port = httpd.server_address[1]
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    base = f"http://127.0.0.1:{port}"
    try:
        assert json.load(urllib.request.urlopen(base + "/", timeout=30))["status"] == "ok"     # readiness probe
        prompt = "w5 w9 w33 w7 w120 w8"
        full = json.load(_post(base + "/v1/completions", {"prompt": prompt, "max_tokens": 12}))   # test/system.sh:73-78
        text = full["choices"][0]["text"]
        assert full["object"] == "text_completion" and full["usage"]["completion_tokens"] <= 12 and text
        # the same request streamed: SSE chunks whose deltas concatenate to the same text, then [DONE]
        resp = _post(base + "/v1/completions", {"prompt": prompt, "max_tokens": 12, "stream": True})
        assert resp.headers["Content-Type"].startswith("text/event-stream")
        events = [l[6:] for l in resp.read().decode().split("\n") if l.startswith("data: ")]
        assert events[-1] == "[DONE]"
        chunks = [json.loads(e) for e in events[:-1]]
        assert "".join(c["choices"][0]["text"] for c in chunks) == text
        assert chunks[-1]["choices"][0]["finish_reason"] in ("length", "stop") and len(chunks) >= 3
        # greedy is deterministic; n = 2 gives two identical choices; echo prepends the prompt
        two = json.load(_post(base + "/v1/completions", {"prompt": prompt, "max_tokens": 12, "n": 2, "echo": True}))
        assert [c["text"] for c in two["choices"]] == [prompt + text] * 2
        # a stop string cuts the completion before it
        words = text.split()
        if len(words) >= 3:
            st = json.load(_post(base + "/v1/completions", {"prompt": prompt, "max_tokens": 12, "stop": [words[2]]}))
            assert st["choices"][0]["finish_reason"] == "stop" and words[2] not in st["choices"][0]["text"]
            assert text.startswith(st["choices"][0]["text"])
        # sampling: honoured (seeded, reproducible, and different from greedy for a hot temperature)
        s1 = json.load(_post(base + "/v1/completions", {"prompt": prompt, "max_tokens": 12, "temperature": 5.0, "top_p": 0.95, "seed": 7}))
        s2 = json.load(_post(base + "/v1/completions", {"prompt": prompt, "max_tokens": 12, "temperature": 5.0, "top_p": 0.95, "seed": 7}))
        assert s1["choices"][0]["text"] == s2["choices"][0]["text"] != text
        # what is not implemented is refused, never silently ignored
        for bad in ({"logprobs": 3}, {"presence_penalty": 0.5}, {"temperature": -1}, {"n": 2, "stream": True}, {"best_of": 4}):
            with pytest.raises(urllib.error.HTTPError) as ei:
                _post(base + "/v1/completions", dict({"prompt": prompt, "max_tokens": 4}, **bad))
            assert ei.value.code == 400, bad
        # a token outside the model's vocabulary is a per-request 400 and the server stays up
        with pytest.raises(urllib.error.HTTPError) as ei:
            _post(base + "/v1/completions", {"prompt": "w1 " * 300, "max_tokens": 4})      # longer than the KV cache
        assert ei.value.code == 400
        assert json.load(_post(base + "/v1/completions", {"prompt": prompt, "max_tokens": 3}))["choices"][0]["text"]
    finally:
        httpd.shutdown()
