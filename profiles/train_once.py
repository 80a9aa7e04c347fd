"""One training step at a given batch (driver for ncu). usage: python profiles/train_once.py <B> [steps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import igdm_b200  # noqa
from igdm_b200 import api, capi, tables
B = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = capi.Handle(T=500, precision=1)
beta, _, acum = tables.beta_schedule(500)
h.set_tables(beta, acum, tables.embedding_table(500))
h.set_weights(api.SimpleUNet.load().arrays)
h.set_adam(1e-4)
data = np.random.default_rng(0).uniform(-1, 1, (B, 1, 32, 32)).astype(np.float32)
h.upload_dataset(data)
for k in range(steps):
    print("loss", h.train_step_device(B, 1, k))
