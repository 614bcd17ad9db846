"""The Server container contract end to end (server_controller.go:156-173, test/system.sh:73-78):
load /content/model, listen, GET / -> 200, POST /v1/completions -> tokens identical to the oracle's
greedy continuation."""
import json
import socket
import threading
import time
import urllib.request

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_serve_falcon_over_http(tmp_path):
    from tokenizers import Tokenizer, models, pre_tokenizers
    from oracle import falcon_oracle as FO
    from runbooks_b200 import contract, server
    from util import bf16_bits

    a = FO.FalconArch(vocab_size=512, hidden_size=256, num_layers=2, num_heads=4, head_dim=64)
    params = FO.seeded_params(a, 21)
    md = tmp_path / "model"
    md.mkdir()
    vocab = {"<s>": 0, "</s>": 1, "<unk>": 2, **{f"w{i}": i + 3 for i in range(509)}}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    tok.save(str(md / "tokenizer.json"))
    (md / "tokenizer_config.json").write_text(json.dumps({"eos_token": "</s>"}))
    cfg = {"model_type": "falcon", "architectures": ["FalconForCausalLM"], "vocab_size": 512, "hidden_size": 256,
           "num_hidden_layers": 2, "num_attention_heads": 4, "multi_query": True, "parallel_attn": True,
           "bias": False, "alibi": False, "new_decoder_architecture": False, "layer_norm_epsilon": 1e-5,
           "max_position_embeddings": 128, "tie_word_embeddings": True}
    contract.save_hf_checkpoint(str(md), cfg, ((k, bf16_bits(v)) for k, v in params.items()))
    port = _free_port()
    th = threading.Thread(target=server.serve, args=(str(tmp_path), port, 4, 128), daemon=True)
    th.start()
    url = f"http://127.0.0.1:{port}"
    for _ in range(300):                                  # readiness probe: GET / -> 200
        try:
            with urllib.request.urlopen(url + "/", timeout=1) as r:
                if r.status == 200:
                    break
        except Exception:
            time.sleep(0.2)
    else:
        pytest.fail("server never became ready")
    words = [f"w{i}" for i in np.random.default_rng(1).integers(0, 500, size=10)]
    prompt_ids = [0] + [vocab[w] for w in words]          # the server prepends the tokenizer's <s> (id 0)
    ref, ref_logits = FO.greedy(params, prompt_ids, 8, a)

    def complete(prompt, n):
        req = urllib.request.Request(url + "/v1/completions", data=json.dumps({"prompt": prompt, "max_tokens": n}).encode(),
                                     headers={"Content-Type": "application/json"})
        with urllib.request.urlopen(req, timeout=60) as r:
            assert r.status == 200
            return json.loads(r.read())

    # several concurrent requests exercise the continuous batching path
    outs = [None] * 3
    ths = [threading.Thread(target=lambda i=i: outs.__setitem__(i, complete(" ".join(words), 8))) for i in range(3)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for o in outs:
        assert o["object"] == "text_completion" and o["usage"]["prompt_tokens"] == 11
        got = [vocab.get(w, 2) for w in o["choices"][0]["text"].split()]
        stop = ref.index(1) if 1 in ref else len(ref)     # eos ends generation and is not echoed
        margin = np.diff(np.sort(ref_logits, axis=-1)[:, -2:], axis=-1)[:, 0]
        n_ok = 0
        for g, r_ in zip(got, ref[:stop]):
            if g != r_:
                break
            n_ok += 1
        assert n_ok == len(ref[:stop]) or margin[n_ok] < 0.05, (got, ref)
    req = urllib.request.Request(url + "/v1/completions", data=b"{not json", headers={"Content-Type": "application/json"})
    with pytest.raises(urllib.error.HTTPError) as ei:
        urllib.request.urlopen(req, timeout=10)
    assert ei.value.code == 400
