"""Generates tests/golden/oracle_golden.npz from the CPU oracle (oracle/ddpm_oracle.py) on the
shipped fixtures.  Run from the repo root:  python tests/golden/make_golden.py
The file pins the oracle against drift; the independent pins are the SURVEY.md Appendix F anchors
hard-coded in tests/test_oracle_*.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import igdm_b200  # noqa
from igdm_b200 import api, bson_io
import ddpm_oracle as O

torch.set_num_threads(1)  # deterministic summation order
arrs, _ = bson_io.load_checkpoint(os.path.join(ROOT, "fixtures", "trained_model.bson"))
arrays = [a.flat for a in arrs]
data = (api.load_dataset() * np.float32(2) - np.float32(1)).astype(np.float32)
beta, alpha, acum = O.schedule(500)
pe = O.embedding_table(500)
out = {"beta_bits": beta.view(np.uint32), "acum_bits": acum.view(np.uint32),
       "pe_bits_t1_t250_t500": pe[[0, 249, 499]].view(np.uint32),
       "samp_bits": O.sampler_table(acum).view(np.uint32)}

B = 8
x0 = data[:B]
ts = np.random.default_rng(1).integers(1, 501, B)
eps = np.random.default_rng(2).standard_normal(x0.shape).astype(np.float32)
out["b8_ts"] = ts
xt = O.q_sample(x0, ts, eps, acum)
out["b8_xt_bits"] = xt.view(np.uint32)
net = O.Net(arrays)
with torch.no_grad():
    out["b8_eps_test"] = O.unet_forward(net, torch.tensor(xt), torch.tensor(pe[ts - 1])).numpy()
    out["b8_eps_train"] = O.unet_forward(net, torch.tensor(xt), torch.tensor(pe[ts - 1]), train=True).numpy()
# 3 training steps (Adam 1e-4) from the shipped weights
net = O.Net(arrays)
opt = O.Adam(net.trainable(), eta=1e-4)
losses, gnorms = [], None
for k in range(3):
    ts_k = np.random.default_rng(100 + k).integers(1, 501, B)
    eps_k = np.random.default_rng(200 + k).standard_normal(x0.shape).astype(np.float32)
    loss, grads = O.train_step(net, opt, data[k * B:(k + 1) * B], ts_k, eps_k, acum, pe)
    losses.append(loss)
    if k == 0:
        gnorms = np.array([np.linalg.norm(g) for g in grads])
out["train3_losses"] = np.array(losses, np.float32)
out["train3_grad_norms_step0"] = gnorms.astype(np.float32)
out["train3_final_w62"] = net.arrays()[62]
# sampler: N=2, t_start=6
xT = np.random.default_rng(5).standard_normal((2, 1, 32, 32)).astype(np.float32)
z = np.random.default_rng(6).standard_normal((5, 2, 1, 32, 32)).astype(np.float32)
out["samp_t6"] = O.generate_image(O.Net(arrays), xT, z, acum, pe, t_start=6)
# apply_noise
img = np.full((64, 64), 0.7)
e = np.random.default_rng(7).standard_normal((64, 64))
out["apply_noise_64"] = O.apply_noise_f64(img, e)
# philox known answer: counter=0,key=0 and the device_normal head
out["philox_zero"] = O.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))
out["devnormal_head"] = O.device_normal(3, np.array([0, 1]), 0)[:, :8]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
