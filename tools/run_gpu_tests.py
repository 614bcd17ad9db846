"""Runs the -m gpu tests one test FUNCTION per process, each under a timeout, so that a kernel
trap (which poisons the CUDA context) or a hang in one test cannot hide the results of the others.
Writes gpurun_out/gpu_tests.json + per-function logs. Used for development probes; the driver's
own round-end run is the plain `pytest -m gpu`.
  --no-x   do not stop a function at its first failing parametrisation (with -x, the default, one bad
           parameter hides the outcome of all the later ones)."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
NO_X = "--no-x" in sys.argv


def main():
    os.makedirs(OUT, exist_ok=True)
    sel = [a for a in sys.argv[1:] if a != "--no-x"] or ["tests"]
    r = subprocess.run([sys.executable, "-m", "pytest", *sel, "-m", "gpu", "--collect-only", "-q"],
                       cwd=ROOT, capture_output=True, text=True)
    funcs = []
    for line in r.stdout.splitlines():
        if "::" in line:
            f = line.split("[")[0].strip()
            if f not in funcs:
                funcs.append(f)
    results = {}
    t_all = time.time()
    for f in funcs:
        t0 = time.time()
        log = os.path.join(OUT, "test_" + f.replace("/", "_").replace("::", "__") + ".log")
        try:
            p = subprocess.run([sys.executable, "-m", "pytest", f, "-m", "gpu", "-q", "-s", *([] if NO_X else ["-x"]),
                                "--no-header", "-p", "no:cacheprovider"],
                               cwd=ROOT, capture_output=True, text=True, timeout=420)
            out, code = p.stdout + p.stderr, p.returncode
        except subprocess.TimeoutExpired as e:
            out = ((e.stdout or b"").decode(errors="ignore") if isinstance(e.stdout, bytes) else (e.stdout or ""))
            out += "\nTIMEOUT"
            code = -9
        with open(log, "w") as fh:
            fh.write(out)
        tail = [l for l in out.splitlines() if l.strip()][-1:] or [""]
        results[f] = dict(code=code, secs=round(time.time() - t0, 1), tail=tail[0][-200:])
        print(f"[{code:>3}] {time.time() - t0:6.1f}s {f}  {tail[0][-120:]}", flush=True)
        # echo the measured error lines so one gpurun tail shows them
        for l in out.splitlines():
            if "rel_err" in l or "b200w:" in l or l.startswith("E  ") or "attention B" in l:
                print("      " + l[:220], flush=True)
    results["_total_secs"] = round(time.time() - t_all, 1)
    with open(os.path.join(OUT, "gpu_tests.json"), "w") as fh:
        json.dump(results, fh, indent=1)
    bad = [f for f, v in results.items() if isinstance(v, dict) and v["code"] != 0]
    print(f"{len(funcs) - len(bad)}/{len(funcs)} test functions green; failing: {bad}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
