"""Smallest run that launches every cluster / mbarrier / TMEM kernel of the library once (driver for compute-sanitizer):
inference forward, train-mode forward + backward + Adam (CTA-pair convs with fused BatchNorm reductions, wgrad, ConvTranspose,
bulk-staged first-conv backward), three sampler steps (tensor-core first conv, fused final epilogue).
usage: compute-sanitizer --tool racecheck|synccheck|memcheck python profiles/sanitize_once.py [B]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa
from igdm_b200 import api, capi, tables
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
h = capi.Handle(T=500, precision=1)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
h.set_adam(1e-4)
h.set_option("train_graph", 0)
h.set_option("use_graph", 0)
rng = np.random.default_rng(0)
x = rng.standard_normal((B, 1, 32, 32)).astype(np.float32)
ts = rng.integers(1, 501, B)
e = h.predict_eps(x, ts)
l = h.train_step(x, ts, rng.standard_normal(x.shape).astype(np.float32))
s = h.sample(B, seed=1, t_start=4)
print("ok", float(np.abs(e).mean()), l, float(np.abs(s).mean()), h.counter("launches"))
