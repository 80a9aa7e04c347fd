"""Probe: does the tcgen05 wgrad kernel run in a given precision mode? usage: wgrad_probe.py fp16|bf16"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import igdm_b200  # noqa
from igdm_b200 import api, capi, tables
import ddpm_oracle as O
import torch
mode = sys.argv[1]
prec = {"fp16": 1, "bf16": 2}[mode]
h = capi.Handle(T=500, precision=prec)
beta, _, acum = tables.beta_schedule(500); pe = tables.embedding_table(500)
h.set_tables(beta, acum, pe)
arrays = api.SimpleUNet.load().arrays
h.set_weights(arrays)
data = (api.load_dataset() * np.float32(2) - np.float32(1)).astype(np.float32)
B = 16
x0 = data[:B]; ts = np.random.default_rng(1).integers(1, 501, B); eps = np.random.default_rng(2).standard_normal(x0.shape).astype(np.float32)
net = O.Net(arrays)
l, _, _ = O.train_step_loss(net, x0, ts, eps, acum, pe, update_stats=False); l.backward()
h.set_option("conv_impl", 1)
loss_s, g_s = h.loss_and_grad(x0, ts, eps)
h.set_option("conv_impl", 0)
loss, g = h.loss_and_grad(x0, ts, eps)
print(mode, "loss", loss, float(l))
for k in (6, 12, 18, 24, 30, 38, 44, 50, 56):
    want = net.flat[k].grad.numpy().ravel()
    r = lambda a: float(np.linalg.norm(a - want) / np.linalg.norm(want))
    print(f"  W{k}: tc rel {r(g[k]):.4f}   simt rel {r(g_s[k]):.4f}")
