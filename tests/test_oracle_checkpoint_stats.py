"""A pin of the oracle against numbers the REFERENCE ITSELF produced: the BatchNorm running statistics inside the shipped
checkpoints.

Flux's BatchNorm updates its running mean / variance in place during every train-mode forward of the reference's own
training run (/root/reference/src/train_brain.jl:112-140,263-272; Flux BatchNorm: momentum 0.1, unbiased running variance),
and `@save ... model opt [epoch]` (:295-300) wrote them into `trained_model.bson` and `ddpm_epoch_*.bson`.  They are the
per-channel mean and variance of all ten pre-BatchNorm convolution outputs under the reference's real arithmetic, at the
checkpoint's weights (moving average over the last ~20 batches of 64).  The oracle's train-mode forward, fed the same
dataset with fresh random timesteps and noise, must therefore reproduce all twenty vectors -- and it does, to 1-7 %, for
three independently written checkpoints, down to the deepest layers behind the skip concatenation.

This is a statistical pin (not a bit-exact one; Julia cannot run here), so the test also shows what it can and cannot see:
mutations of the oracle's semantics (data not rescaled to [-1,1], embedding halves swapped, concat order swapped, 3x3
kernels transposed, a different schedule length) move the statistics by 25-130 %.  Not identifiable this way: the
180-degree kernel flip (the dataset's statistics are symmetric under it; pinned by the delta-input test in
test_oracle_net.py) and off-by-one timestep rows (neighbouring embeddings are nearly equal; pinned by the bit-exact tables)."""
import os

import numpy as np
import pytest
import torch

from conftest import FIX

NB = 8            # batches of 64 averaged for a positive check (noise of the estimate ~1 %)
NB_MUT = 4


def _layer_stats(oracle, path, T, nb, seed=0, mutate=None):
    """rel-L2 distance between the stored running (mean, var) of the 10 BatchNorm layers and the batch statistics of the
    oracle's train-mode forward at the checkpoint's weights, averaged over nb batches of 64 (train_brain.jl:197-206,225-241)."""
    from igdm_b200 import api, tables

    torch.set_num_threads(min(16, os.cpu_count() or 1))
    model = api.SimpleUNet.load(path)
    arrays = [a.copy() for a in model.arrays]
    raw = api.load_dataset()
    data = raw if mutate == "unscaled" else raw * np.float32(2) - np.float32(1)       # train_brain.jl:250-251
    pe = tables.embedding_table(T)
    _, _, acum = tables.beta_schedule(T)
    if mutate == "emb_swapped":
        pe = np.concatenate([pe[:, 64:], pe[:, :64]], axis=1)
    if mutate == "concat_swapped":
        # the only 128 => 64 3x3 conv is the one behind cat(up, h1): swap its two 64-channel input halves
        i9 = [i for i, n in enumerate(oracle.array_lengths()) if n == 9 * 128 * 64][-1]
        w = arrays[i9].reshape(64, 128, 3, 3)
        arrays[i9] = np.concatenate([w[:, 64:], w[:, :64]], axis=1).reshape(-1).copy()
    keep = oracle._conv_w
    if mutate == "transposed_kernel":
        oracle._conv_w = lambda flat, ci, co, k: keep(flat, ci, co, k).transpose(2, 3)
    net = oracle.Net(arrays)
    rng = np.random.default_rng(seed)
    acc = None
    try:
        for _ in range(nb):
            idx = rng.permutation(500)[:64]
            x0 = data[idx]
            ts = rng.integers(1, T + 1, 64)
            eps = rng.standard_normal(x0.shape).astype(np.float32)
            xt = oracle.q_sample(x0, ts, eps, acum)
            taps = {}
            with torch.no_grad():
                oracle.unet_forward(net, torch.tensor(xt), torch.tensor(pe[ts - 1]), train=True, update_stats=False, taps=taps)
            cur = []
            for l in range(1, 11):
                y = taps["y%d" % l]
                cur += [y.mean(dim=(0, 2, 3)).numpy(), y.var(dim=(0, 2, 3), unbiased=True).numpy()]
            acc = cur if acc is None else [a + c for a, c in zip(acc, cur)]
    finally:
        oracle._conv_w = keep
    stored = [model.arrays[i] for i, t in enumerate(oracle.trainable_mask()) if not t]
    assert len(stored) == 20
    return [float(np.linalg.norm(a / nb - s) / np.linalg.norm(s)) for a, s in zip(acc, stored)]


@pytest.mark.parametrize("name,T,max_tol,med_tol", [
    ("trained_model.bson", 500, 0.12, 0.035),        # measured (10 batches): max 0.071, median 0.013
    ("ddpm_epoch_95.bson", 5, 0.12, 0.035),          # the epoch checkpoints were trained with num_timesteps = 5: 0.049 / 0.018
    ("ddpm_epoch_5.bson", 5, 0.25, 0.12),            # early training, the moving average lags the weights: 0.139 / 0.080
])
def test_oracle_forward_reproduces_the_running_statistics_flux_wrote(oracle, name, T, max_tol, med_tol):
    r = _layer_stats(oracle, os.path.join(FIX, name), T, NB)
    assert max(r) < max_tol and float(np.median(r)) < med_tol, r


@pytest.mark.parametrize("mutate,T", [("unscaled", 500), ("emb_swapped", 500), ("concat_swapped", 500),
                                      ("transposed_kernel", 500), (None, 1000)])
def test_the_statistics_pin_sees_wrong_semantics(oracle, mutate, T):
    """Negative controls: each mutation must move at least one statistics vector far outside the band of the true
    semantics (measured: unscaled 0.25, embedding halves swapped 1.3, concat swapped 0.67, transposed kernels 0.47,
    T = 1000 instead of 500: 0.42; the true semantics: 0.07)."""
    r = _layer_stats(oracle, os.path.join(FIX, "trained_model.bson"), T, NB_MUT, mutate=mutate)
    assert max(r) > 0.18, (mutate, T, r)


def test_recorded_sweep_over_all_shipped_checkpoints():
    """tests/golden/checkpoint_stats_all.json (made by tests/golden/make_checkpoint_stats.py in the build container, where
    /root/reference is mounted): the same comparison for ALL 20 checkpoints the reference ships, plus the oracle's train-mode
    eps-MSE at each of them, which must fall along the epochs and end at the ~0.23 of the published training_loss.png."""
    import json
    from conftest import GOLDEN
    d = json.load(open(os.path.join(GOLDEN, "checkpoint_stats_all.json")))
    assert len(d) == 20
    ep = {int(k.split("_")[2].split(".")[0]): v for k, v in d.items() if k.startswith("ddpm_epoch_")}
    assert sorted(ep) == list(range(5, 100, 5))
    for e, v in ep.items():
        if e >= 15:                                  # earlier checkpoints: the moving average lags the fast-moving weights
            assert v["max_rel_l2"] < 0.10 and v["median_rel_l2"] < 0.03, (e, v)
    assert d["trained_model.bson"]["max_rel_l2"] < 0.10 and d["trained_model.bson"]["median_rel_l2"] < 0.03
    loss = [ep[e]["train_mode_eps_mse"] for e in sorted(ep)]
    assert loss[0] > 0.7 and all(loss[i + 2] < loss[i] for i in range(len(loss) - 2)), loss     # falls (one 5-epoch blip allowed)
    assert abs(loss[-1] - 0.23) < 0.02, loss[-1]
