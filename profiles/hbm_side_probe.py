"""HBM-side kernels of one sampler evaluation (first conv, 2x2 max-pool, ConvTranspose) against the pure write / read
stream rates of the same buffer size, plus the parity of the first-conv variants (option conv1_tc = 0 CUDA cores,
1 tcgen05 with the timestep constants added in the epilogue, 2 constants folded into the contraction).
usage: python profiles/hbm_side_probe.py [n_images]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa: E402,F401
from igdm_b200 import api, capi, tables  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1300
h = capi.Handle(T=500, precision=capi.PREC_FP16)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
rep = {"n_images": n}

# parity of the variants on a ragged batch (the layer alone: time_kernel leaves a1 in place)
out = {}
for m in (0, 1, 2):
    h.set_option("conv1_tc", m)
    h.time_kernel("conv1", 37, 1)
    out[m] = h.debug_fetch("infer:a1")
ref = out[0]
ulp = np.maximum(2.0 ** (np.floor(np.log2(np.maximum(np.abs(ref), 2.0 ** -14))) - 10), 2e-5)
for m in (1, 2):
    d = np.abs(out[m] - ref)
    rep[f"conv1_tc{m}_vs_simt"] = {"frac_mismatch": float(np.mean(d > 0)), "max_err_ulp": float((d / ulp).max()),
                                   "max_abs": float(np.abs(ref).max()), "halo_or_nan": bool(~np.isfinite(out[m]).all())}
rep["conv1_tc2_vs_tc1_frac_mismatch"] = float(np.mean(out[2] != out[1]))
print(json.dumps(rep), flush=True)

for name in ("probe_fill", "probe_read", "pool", "up2"):
    ms, by, fl = h.time_kernel(name, n, 20)
    rep[name] = {"us": ms * 1e3, "GBps": by / ms / 1e6}
    print(name, rep[name], flush=True)
for m in (0, 1, 2):
    h.set_option("conv1_tc", m)
    ms, by, fl = h.time_kernel("conv1", n, 20)
    rep[f"conv1_tc{m}"] = {"us": ms * 1e3, "GBps": by / ms / 1e6}
    print("conv1", m, rep[f"conv1_tc{m}"], flush=True)
    ms, by, fl = h.time_kernel("forward_infer", n, 20)
    rep[f"forward_infer_conv1_tc{m}"] = {"us": ms * 1e3}
    print("forward_infer", m, ms * 1e3, flush=True)
h.set_option("conv1_tc", 2)
print(json.dumps(rep))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "hbm_side_probe.json"), "w") as fh:
    json.dump(rep, fh, indent=1)
