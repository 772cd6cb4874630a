#!/usr/bin/env python
"""Runs every GPU test in its own process with a timeout (a deadlocked kernel must not take the box down) and
writes a summary to gpurun_out/probe.log.  Usage: python tests/gpu_probe.py [pytest node ids...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
ids = sys.argv[1:]
if not ids:
    out = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "--collect-only", "-q"], cwd=ROOT,
                         capture_output=True, text=True).stdout
    ids = [l.strip() for l in out.splitlines() if "::" in l]
log = open(os.path.join(ROOT, "gpurun_out", "probe.log"), "w")
for nid in ids:
    t0 = time.time()
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", nid, "-x", "-q", "-s", "--no-header", "-p", "no:cacheprovider"],
                           cwd=ROOT, capture_output=True, text=True, timeout=240)
        status = "PASS" if r.returncode == 0 else f"FAIL({r.returncode})"
        tail = (r.stdout + r.stderr)[-3000:] if r.returncode != 0 else "\n".join(
            l for l in r.stdout.splitlines() if "engine-vs" in l or "TFLOP" in l)
    except subprocess.TimeoutExpired as e:
        status, tail = "TIMEOUT", ((e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace"))[-2000:]
    line = f"[{status}] {nid} ({time.time() - t0:.1f}s)"
    print(line, flush=True)
    log.write(line + "\n" + (tail + "\n" if tail else ""))
    log.flush()
