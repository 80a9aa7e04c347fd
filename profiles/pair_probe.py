"""CTA-pair (cta_group::2) convolution against the single-CTA kernels: bitwise layer comparison on a ragged batch,
then timings.  usage: python profiles/pair_probe.py <mask> [n_images]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa: E402,F401
from igdm_b200 import api, capi, tables  # noqa: E402

mask = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
NAMES = ["y2", "y3", "y4", "y5", "y6", "u", "y7", "y8", "y9", "y10"]
h = capi.Handle(T=500, precision=1)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
rng = np.random.default_rng(0)
for B in (9, 2, 37):
    xt = rng.standard_normal((B, 1, 32, 32)).astype(np.float32)
    ts = rng.integers(1, 501, B).astype(np.int32)
    h.set_option("tc_pair", 0)
    e0 = h.predict_eps(xt, ts, train_mode=True)
    ref = {nm: h.debug_fetch(nm) for nm in NAMES}
    h.set_option("tc_pair", mask)
    e1 = h.predict_eps(xt, ts, train_mode=True)
    got = {nm: h.debug_fetch(nm) for nm in NAMES}
    print("B", B, {nm: float(np.abs(got[nm] - ref[nm]).max()) for nm in NAMES}, "eps", float(np.abs(e1 - e0).max()), flush=True)
    e0 = h.predict_eps(xt, ts, train_mode=False)
    h.set_option("tc_pair", 0)
    e1 = h.predict_eps(xt, ts, train_mode=False)
    print("   test-mode eps max diff", float(np.abs(e1 - e0).max()), flush=True)
for name in ("conv_l2", "conv_l3", "conv_l4", "conv_l9"):
    for m in (0, mask):
        h.set_option("tc_pair", m)
        ms, by, fl = h.time_kernel(name, n, 20)
        print(f"{name} pair={m} n={n}: {ms*1e3:.1f} us  {fl/ms/1e9 if fl else 0:.1f} TFLOP/s", flush=True)
for m in (0, 1, 2, 4, 8, 3, 7, 15):
    h.set_option("tc_pair", m)
    ms, by, fl = h.time_kernel("forward_infer", n, 20)
    print(f"forward_infer pair={m} n={n}: {ms*1e3:.1f} us  {fl/ms/1e9 if fl else 0:.1f} TFLOP/s", flush=True)
xT = rng.standard_normal((5, 1, 32, 32)).astype(np.float32)
z = rng.standard_normal((9, 5, 1, 32, 32)).astype(np.float32)
h.set_option("tc_pair", 0)
a = h.sample(5, x_T=xT, z=z, t_start=10)
h.set_option("tc_pair", mask)
b = h.sample(5, x_T=xT, z=z, t_start=10)
print("sampler max diff", float(np.abs(a - b).max()), flush=True)
h.set_option("tc_pair", mask)
img = h.sample_device(1024, seed=1, t_start=60)
print("sample ok", flush=True)
