"""Tiny driver for ncu: times one named libddpm kernel on synthetic device-resident data.
usage: python profiles/run_kernel.py <kernel-name> <n_images> [iters] [precision]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa: E402,F401
from igdm_b200 import api, capi, tables  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
prec = {"fp32": 0, "fp16": 1, "bf16": 2, "tf32": 3}[sys.argv[4] if len(sys.argv) > 4 else "fp16"]
h = capi.Handle(T=500, precision=prec)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
ms, by, fl = h.time_kernel(name, n, iters)
print(f"{name} n={n}: {ms*1e3:.1f} us  {fl/ms/1e9 if fl else 0:.1f} TFLOP/s  {by/ms/1e6 if by else 0:.1f} GB/s")
