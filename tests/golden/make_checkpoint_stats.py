"""All 20 checkpoints the reference ships (/root/reference/ddpm_epoch_{5..95}.bson, trained_model.bson) against the oracle:
(1) the BatchNorm running statistics Flux wrote into each file vs the batch statistics of the oracle's train-mode forward at
that file's weights (the pin of tests/test_oracle_checkpoint_stats.py, here for every checkpoint), (2) the oracle's
train-mode eps-MSE at that file's weights (must fall along the epochs like the published training_loss.png).
Runs in the build container only (reads /root/reference); writes tests/golden/checkpoint_stats_all.json.
usage: python tests/golden/make_checkpoint_stats.py"""
import glob
import json
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import igdm_b200  # noqa: E402,F401
import ddpm_oracle as O  # noqa: E402
from igdm_b200 import api, tables  # noqa: E402

REF = "/root/reference"
NB = 8


def run(path, T):
    model = api.SimpleUNet.load(path)
    net = O.Net([a.copy() for a in model.arrays])
    data = api.load_dataset() * np.float32(2) - np.float32(1)
    pe = tables.embedding_table(T)
    _, _, acum = tables.beta_schedule(T)
    rng = np.random.default_rng(0)
    acc, losses = None, []
    for _ in range(NB):
        idx = rng.permutation(500)[:64]
        x0 = data[idx]
        ts = rng.integers(1, T + 1, 64)
        eps = rng.standard_normal(x0.shape).astype(np.float32)
        xt = O.q_sample(x0, ts, eps, acum)
        taps = {}
        with torch.no_grad():
            out = O.unet_forward(net, torch.tensor(xt), torch.tensor(pe[ts - 1]), train=True, update_stats=False, taps=taps)
        losses.append(float(((out.numpy() - eps) ** 2).mean()))
        cur = []
        for l in range(1, 11):
            y = taps["y%d" % l]
            cur += [y.mean(dim=(0, 2, 3)).numpy(), y.var(dim=(0, 2, 3), unbiased=True).numpy()]
        acc = cur if acc is None else [a + c for a, c in zip(acc, cur)]
    stored = [model.arrays[i] for i, t in enumerate(O.trainable_mask()) if not t]
    r = [float(np.linalg.norm(a / NB - s) / np.linalg.norm(s)) for a, s in zip(acc, stored)]
    return {"T": T, "max_rel_l2": max(r), "median_rel_l2": float(np.median(r)), "train_mode_eps_mse": float(np.mean(losses))}


if __name__ == "__main__":
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    out = {}
    files = sorted(glob.glob(os.path.join(REF, "ddpm_epoch_*.bson")), key=lambda p: int(re.findall(r"(\d+)\.bson", p)[0]))
    for p in files:
        out[os.path.basename(p)] = run(p, 5)          # the epoch checkpoints were trained with num_timesteps = 5
        print(os.path.basename(p), out[os.path.basename(p)], flush=True)
    out["trained_model.bson"] = run(os.path.join(REF, "trained_model.bson"), 500)
    print("trained_model.bson", out["trained_model.bson"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "checkpoint_stats_all.json"), "w") as fh:
        json.dump(out, fh, indent=1)
