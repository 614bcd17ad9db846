"""Container entry point of the Server Deployment's container "serve"
(internal/controller/server_controller.go:114-205):

    python -m runbooks_b200.server            # ENTRYPOINT of the server image

Contract (docs/container-contract.md:50-55; server_controller.go:156-173): listen on :8080,
`GET /` answers 200 once the model in /content/model is loaded (readiness probe), and — what
test/system.sh:73-78 and the basaran image speak — `POST /v1/completions {"prompt", "max_tokens"}`
returns an OpenAI-style completion. Greedy decoding, continuous batching over the engine's cache
slots (one scheduler thread owns the GPU; HTTP handler threads only queue requests).
"""
from __future__ import annotations

import argparse
import json
import os
import queue
import sys
import threading
import time
import traceback
import uuid
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

from . import contract


class Scheduler(threading.Thread):
    def __init__(self, engine, tokenizer, max_wait_s: float = 0.002):
        super().__init__(daemon=True)
        from .infer import Generator
        self.gen = Generator(engine, eos_id=tokenizer.eos_id)
        self.tok, self.q, self.max_wait = tokenizer, queue.Queue(), max_wait_s
        self.ready = threading.Event()
        self.failed: str | None = None

    def submit(self, prompt: str, max_tokens: int):
        done = threading.Event()
        item = {"prompt": prompt, "max_tokens": max_tokens, "done": done, "result": None, "error": None}
        self.q.put(item)
        return item

    def run(self):
        self.ready.set()
        pending = {}
        try:
            while True:
                # admit as many queued requests as there are free slots
                block = not self.gen.active
                while self.gen.free:
                    try:
                        item = self.q.get(timeout=0.5 if block else 0)
                    except queue.Empty:
                        break
                    block = False
                    try:
                        ids = self.tok.encode(item["prompt"])
                        if self.tok.bos_id is not None:
                            ids = [self.tok.bos_id] + ids
                        req = self.gen.add(ids, item["max_tokens"])
                        pending[id(req)] = (req, item, len(ids))
                    except Exception as e:  # noqa: BLE001 — per-request failure, the server lives on
                        item["error"] = str(e)
                        item["done"].set()
                self.gen.step()
                for key in [k for k, (r, _, _) in pending.items() if r.done]:
                    req, item, n_prompt = pending.pop(key)
                    out = req.out[:-1] if (self.tok.eos_id is not None and req.out and req.out[-1] == self.tok.eos_id) else req.out
                    item["result"] = {"text": self.tok.decode(out), "prompt_tokens": n_prompt,
                                      "completion_tokens": len(req.out),
                                      "finish_reason": "stop" if len(out) != len(req.out) else "length"}
                    item["done"].set()
        except BaseException:  # noqa: BLE001 — a CUDA failure is fatal: fail readiness, exit non-zero
            self.failed = traceback.format_exc()
            sys.stderr.write(self.failed)
            os._exit(1)


def make_handler(sched: Scheduler, model_name: str):
    class H(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def _send(self, code, obj):
            body = json.dumps(obj).encode()
            self.send_response(code)
            self.send_header("Content-Type", "application/json")
            self.send_header("Content-Length", str(len(body)))
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, fmt, *args):  # one JSON line per request on stdout
            print(json.dumps({"event": "http", "msg": fmt % args}), flush=True)

        def do_GET(self):
            if self.path in ("/", "/healthz"):
                ok = sched.ready.is_set() and not sched.failed
                self._send(200 if ok else 503, {"status": "ok" if ok else "loading", "model": model_name})
            elif self.path == "/v1/models":
                self._send(200, {"object": "list", "data": [{"id": model_name, "object": "model"}]})
            else:
                self._send(404, {"error": "not found"})

        def do_POST(self):
            if self.path != "/v1/completions":
                return self._send(404, {"error": "not found"})
            try:
                n = int(self.headers.get("Content-Length", "0"))
                req = json.loads(self.rfile.read(n) or b"{}")
                prompt = req["prompt"]
                if isinstance(prompt, list):
                    prompt = prompt[0]
                max_tokens = int(req.get("max_tokens", 16))
                if not isinstance(prompt, str) or max_tokens < 1:
                    raise ValueError("prompt must be a string and max_tokens >= 1")
            except Exception as e:  # noqa: BLE001
                return self._send(400, {"error": f"bad request: {e}"})
            item = sched.submit(prompt, max_tokens)
            item["done"].wait()
            if item["error"]:
                return self._send(400, {"error": item["error"]})
            r = item["result"]
            self._send(200, {
                "id": "cmpl-" + uuid.uuid4().hex[:24], "object": "text_completion", "created": int(time.time()),
                "model": model_name,
                "choices": [{"index": 0, "text": r["text"], "logprobs": None, "finish_reason": r["finish_reason"]}],
                "usage": {"prompt_tokens": r["prompt_tokens"], "completion_tokens": r["completion_tokens"],
                          "total_tokens": r["prompt_tokens"] + r["completion_tokens"]}})

    return H


def load_engine(model_dir: str, max_batch: int, max_ctx: int | None):
    from .infer import InferEngine, ServeArch

    cfg = contract.read_hf_config(model_dir)
    arch = ServeArch.from_hf_config(cfg, max_ctx)
    e = InferEngine(int(os.environ.get("B200W_DEVICE", "0")))
    e.init_infer(arch, max_batch=max_batch)
    wanted = {n for n, _ in e.infer_params()}
    seen = set()
    for name, arr in contract.iter_safetensors(model_dir):
        if name in wanted:
            e.infer_load_tensor(name, arr)
            seen.add(name)
    if wanted - seen:
        raise KeyError(f"checkpoint lacks {sorted(wanted - seen)[:3]} ... ({len(wanted - seen)} tensors)")
    return e, cfg


def serve(content: str, port: int, max_batch: int, max_ctx: int | None):
    model_dir = os.path.join(content, "model")
    t0 = time.time()
    engine, cfg = load_engine(model_dir, max_batch, max_ctx)
    tok = contract.Tokenizer(model_dir)
    sched = Scheduler(engine, tok)
    sched.start()
    name = cfg.get("_name_or_path") or cfg.get("model_type", "model")
    httpd = ThreadingHTTPServer(("0.0.0.0", port), make_handler(sched, name))
    print(json.dumps({"event": "ready", "port": port, "model": name, "load_seconds": round(time.time() - t0, 2),
                      "max_batch": max_batch}), flush=True)
    httpd.serve_forever()


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="runbooks_b200.server")
    ap.add_argument("--content", default=contract.CONTENT)
    ap.add_argument("--port", type=int, default=8080)       # server_controller.go:156-161
    ap.add_argument("--max-batch", type=int, default=32)
    ap.add_argument("--max-ctx", type=int, default=None)
    a = ap.parse_args(argv)
    try:
        serve(a.content, a.port, a.max_batch, a.max_ctx)
        return 0
    except BaseException:  # noqa: BLE001
        traceback.print_exc()
        return 1


if __name__ == "__main__":
    sys.exit(main())
