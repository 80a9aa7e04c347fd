import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa
from igdm_b200 import api, capi, tables
h = capi.Handle(T=500, precision=1)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
B = 4
h32 = capi.Handle(T=500, precision=0)
h32.set_tables(beta, acum, tables.embedding_table(500))
h32.set_weights(api.SimpleUNet.load().arrays)
rng = np.random.default_rng(0)
cases = {"zeros": np.zeros((B, 1, 32, 32), np.float32), "ones": np.ones((B, 1, 32, 32), np.float32),
         "1+2^-12": np.full((B, 1, 32, 32), 1 + 2.0 ** -12, np.float32),
         "bf16-exact randn": (np.round(rng.standard_normal((B, 1, 32, 32)) * 16) / 16).astype(np.float32),
         "randn": rng.standard_normal((B, 1, 32, 32)).astype(np.float32)}
z = np.zeros((1, B, 1, 32, 32), np.float32)
for name, xT in cases.items():
    out = {}
    for m in (0, 1):
        h.set_option("conv1_tc", m)
        h.sample(B, x_T=xT, z=z, t_start=2)
        out[m] = h.debug_fetch("infer:a1")
    h32.sample(B, x_T=xT, z=z, t_start=2)
    ref = h32.debug_fetch("infer:a1")
    ulp = np.maximum(np.abs(ref), 6.1e-5)
    ulp = 2.0 ** (np.floor(np.log2(ulp)) - 10)
    for m in (0, 1):
        e = np.abs(out[m] - ref) / ulp
        print(f"   conv1_tc={m} vs fp32-mode a1: max err {e.max():.3f} ulp, frac > 0.5001 ulp {np.mean(e > 0.5001):.4f}", flush=True)
    d = np.abs(out[0] - out[1])
    print(f"{name}: mismatch {np.mean(d > 0):.4f} max abs {d.max():.3e} max|a1| {np.abs(out[0]).max():.3f}", flush=True)
