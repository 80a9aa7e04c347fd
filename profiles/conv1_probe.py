"""Tensor-core first conv (hi/lo split operands) against the CUDA-core FP32 kernel inside the sampler.
Note: after a full forward the a1 buffer is aliased by later layers -- the layer-level comparison lives in
tests/test_gpu_tc.py (ddpm_time_kernel("conv1") leaves the first conv's output in place); this probe reports the
sampler-level drift and the timings."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa: E402,F401
from igdm_b200 import api, capi, tables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
h = capi.Handle(T=500, precision=1)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
rng = np.random.default_rng(0)
for B, tstart in ((5, 2), (37, 3), (3, 500)):
    xT = (3.0 * rng.standard_normal((B, 1, 32, 32))).astype(np.float32)
    z = rng.standard_normal((tstart - 1, B, 1, 32, 32)).astype(np.float32)
    out = {}
    for m in (0, 1):
        h.set_option("conv1_tc", m)
        img = h.sample(B, x_T=xT, z=z, t_start=tstart)
        out[m] = (img, h.debug_fetch("infer:a1"))
    d = np.abs(out[0][1] - out[1][1])
    print(f"B={B} t_start={tstart}: a1 max abs diff {d.max():.3e} (max |a1| {np.abs(out[0][1]).max():.3f}), "
          f"mismatching elements {int((d > 0).sum())} of {d.size}; sample max diff {np.abs(out[0][0] - out[1][0]).max():.3e}", flush=True)
for m in (0, 1):
    h.set_option("conv1_tc", m)
    ms, by, fl = h.time_kernel("conv1", n, 20)
    print(f"conv1 tc={m} n={n}: {ms*1e3:.1f} us {by/ms/1e6:.0f} GB/s", flush=True)
    ms, by, fl = h.time_kernel("forward_infer", n, 20)
    print(f"forward_infer conv1_tc={m} n={n}: {ms*1e3:.1f} us", flush=True)
