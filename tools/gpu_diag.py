#!/usr/bin/env python
"""Run the GPU test-suite in isolated sub-processes (a trapping kernel poisons its CUDA context,
so every section gets a fresh process and its own timeout) and append everything to
gpurun_out/diag.log.  Usage (on the GPU box):  python tools/gpu_diag.py [section ...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

PYTEST = [sys.executable, "-m", "pytest", "-q", "-x", "--timeout", "600", "-p", "no:cacheprovider", "-s"]
SECTIONS = {
    "elementwise": PYTEST + ["tests/test_gpu_ops.py", "-k", "embed or layernorm or vocab or error"],
    "ops_f32": PYTEST[:4] + PYTEST[5:] + ["tests/test_gpu_ops.py", "-k", "f32"],
    "linear_bf16": PYTEST[:4] + PYTEST[5:] + ["tests/test_gpu_ops.py", "-k", "linear and bf16"],
    "row_bf16": PYTEST[:4] + PYTEST[5:] + ["tests/test_gpu_ops.py", "-k", "row_attention and bf16"],
    "col_bf16": PYTEST[:4] + PYTEST[5:] + ["tests/test_gpu_ops.py", "-k", "column_attention and bf16"],
    "model_f32": PYTEST[:4] + PYTEST[5:] + ["tests/test_gpu_model.py", "-k", "fp32 or deep or contacts"],
    "model_bf16": PYTEST[:4] + PYTEST[5:] + ["tests/test_gpu_model.py", "-k", "bf16"],
    "smoke": [sys.executable, "__graft_entry__.py", "smoke"],
    "bench_cfg2": [sys.executable, "bench.py", "--steps", "5", "--warmup", "3"],
    "bench_cfg2_f32": [sys.executable, "bench.py", "--steps", "2", "--warmup", "1", "--precision", "fp32"],
    "bench_cfg1": [sys.executable, "bench.py", "--steps", "10", "--warmup", "3", "--workload", "cfg1"],
    "bench_cfg5": [sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--workload", "cfg5"],
    "bench_cfg4": [sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--workload", "cfg4"],
}
TIMEOUTS = {"bench_cfg2": 900, "bench_cfg5": 900, "bench_cfg4": 900}


def main():
    names = sys.argv[1:] or list(SECTIONS)
    log = open(os.path.join(OUT, "diag.log"), "a")
    summary = []
    for name in names:
        cmd = SECTIONS[name]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=TIMEOUTS.get(name, 600))
            rc, out = r.returncode, r.stdout + r.stderr
        except subprocess.TimeoutExpired as e:
            rc = "TIMEOUT"
            out = (e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace") \
                if isinstance(e.stdout, bytes) else str(e.stdout) + str(e.stderr)
        dt = time.time() - t0
        log.write(f"\n===== {name} rc={rc} {dt:.1f}s =====\n{out}\n")
        log.flush()
        tail = "\n".join(out.strip().splitlines()[-6:])
        summary.append(f"[{name}] rc={rc} {dt:.1f}s\n{tail}")
        print(summary[-1], flush=True)
    open(os.path.join(OUT, "diag_summary.txt"), "w").write("\n\n".join(summary) + "\n")


if __name__ == "__main__":
    main()
