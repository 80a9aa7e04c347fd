"""Build libddpm.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "libddpm.cu")
OUT = os.path.join(HERE, "libddpm.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo", "-Xptxas=-v",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(HERE, "..", "include", "libddpm.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", OUT, SRC, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    elif verbose:  # ptxas -v statistics go to csrc/.ptxas.log; show only real diagnostics
        keep = [l for l in res.stderr.splitlines()
                if l.strip() and not l.startswith("ptxas info") and "bytes stack frame" not in l]
        sys.stderr.write("\n".join(keep) + ("\n" if keep else ""))
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libddpm.so")
    with open(os.path.join(HERE, "csrc", ".ptxas.log"), "w") as fh:
        fh.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
