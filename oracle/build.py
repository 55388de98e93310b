"""Build the CPU oracle shared library (test infrastructure; see gstex_oracle.c).

Tries an OpenMP build first (the system gcc has libgomp; the wrapper on PATH may not),
falling back to a single-threaded build.  Output: oracle/_build/liboracle.so
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gstex_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")

BASE = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c11", "-Wall", "-Wextra"]


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    attempts = []
    for cc in ("/usr/bin/gcc", shutil.which("gcc") or "gcc", "cc"):
        attempts.append([cc, *BASE, "-fopenmp", "-o", OUT, SRC, "-lm"])
    for cc in ("/usr/bin/gcc", shutil.which("gcc") or "gcc", "cc"):
        attempts.append([cc, *BASE, "-Wno-unknown-pragmas", "-o", OUT, SRC, "-lm"])
    last = None
    for cmd in attempts:
        try:
            r = subprocess.run(cmd, capture_output=True, text=True)
        except FileNotFoundError as e:  # compiler missing
            last = str(e)
            continue
        if r.returncode == 0:
            return OUT
        last = r.stderr
    raise RuntimeError(f"could not build the CPU oracle:\n{last}")


if __name__ == "__main__":
    print(build(force=True))
