"""Falcon-7B greedy decode throughput at batch 32 (BASELINE.json configs[3]) against the HBM roofline of
SURVEY.md 8d -- the `decode` object of bench.py's N=1 line, alone."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.exit(subprocess.call([sys.executable, os.path.join(ROOT, "bench.py"), "--decode-only"]))
