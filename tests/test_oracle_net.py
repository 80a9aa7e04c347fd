"""Oracle network semantics: known answers, layer-definition cross-checks, gradients (CPU)."""
import os

import numpy as np
import pytest
import torch

from conftest import FIX, GOLDEN, config2_batch, rel_l2


def test_known_answer_network(oracle, model_arrays, tabs):
    """an independent fp32 restatement made during the survey, B=1, test-mode BN."""
    net = oracle.Net(model_arrays)
    i = np.arange(1, 33, dtype=np.float64)
    x = (np.sin(0.1 * i)[None, :] * np.cos(0.07 * i)[:, None] + 0.01 * i[None, :]).astype(np.float32)  # x[j][i]
    want = {2: (0.312139, 0.331127, 1.610716, 1.288811, 0.588145, 0.287682, 0.381552),
            250: (0.010025, 0.378946, 0.547103, 0.949831, 1.040147, -0.147795, -0.225356),
            500: (0.295712, 0.233559, 0.644030, 0.699858, 0.081792, 0.104264, 0.507330)}
    with torch.no_grad():
        for t, w in want.items():
            e = oracle.unet_forward(net, torch.tensor(x).view(1, 1, 32, 32), torch.tensor(tabs["pe"][t - 1]).view(1, -1)).numpy()[0, 0]
            J = lambda a, b: e[b - 1, a - 1]   # Julia eps[i,j]
            got = (e.mean(), e.std(), J(1, 1), J(32, 1), J(1, 32), J(5, 20), J(20, 5))
            assert np.allclose(got, w, atol=2e-5), (t, got, w)


def test_conv_definitions_agree(oracle, model_arrays):
    """torch cross-correlation on flipped weights == index-by-index true convolution (NNlib definition)."""
    net = oracle.Net(model_arrays)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 64, 6, 5)).astype(np.float32)
    y_t = torch.nn.functional.conv2d(torch.tensor(x), oracle._conv_w(net.flat[6], 64, 64, 3), net.flat[7], padding=1).detach().numpy()
    y_n = oracle.conv3x3_true_numpy(x, model_arrays[6], model_arrays[7], 64, 64)
    assert np.abs(y_t - y_n).max() < 2e-5
    x = rng.standard_normal((2, 128, 3, 4)).astype(np.float32)
    y_t = torch.nn.functional.conv_transpose2d(torch.tensor(x), oracle._convT_w(net.flat[36], 128, 64), net.flat[37], stride=2).detach().numpy()
    y_n = oracle.convT2x2_numpy(x, model_arrays[36], model_arrays[37], 128, 64)
    assert np.abs(y_t - y_n).max() < 2e-5
    # asymmetric delta input pins the orientation: x = delta at (j=1,i=2) -> y[j,i] = w[a,b] with i = 2+a-2, j = 1+b-2
    d = np.zeros((1, 64, 5, 5), np.float32)
    d[0, 3, 1, 2] = 1
    y = oracle.conv3x3_true_numpy(d, model_arrays[6], np.zeros(64, np.float32), 64, 64)
    w = np.asarray(model_arrays[6]).reshape(64, 64, 3, 3)  # [co][ci][b][a]
    for a in range(1, 4):
        for b in range(1, 4):
            # y[i,j] = w[a,b] x[i+2-a, j+2-b]  with x nonzero at (i0,j0)=(3,2) 1-based => i = i0-2+a, j = j0-2+b
            ii, jj = 3 - 2 + a, 2 - 2 + b
            assert np.isclose(y[0, 7, jj - 1, ii - 1], w[7, 3, b - 1, a - 1])


def test_checkpoint_self_consistency(oracle, model_arrays, dataset, tabs):
    """The shipped checkpoints reproduce their training losses only under these semantics
    (SURVEY.md Appendix D): trained_model @T=500 ~0.10-0.11, ddpm_epoch_95 @T=5 ~0.22."""
    from igdm_b200 import bson_io
    x0, ts, eps = config2_batch(dataset)
    net = oracle.Net(model_arrays)
    with torch.no_grad():
        xt = torch.tensor(oracle.q_sample(x0, ts, eps, tabs["acum"]))
        l = float(oracle.mse(oracle.unet_forward(net, xt, torch.tensor(tabs["pe"][ts - 1])), torch.tensor(eps)))
    assert 0.08 < l < 0.13, l
    a95, meta = bson_io.load_checkpoint(os.path.join(FIX, "ddpm_epoch_95.bson"))
    assert meta["epoch"] == 95 and abs(meta["eta"] - 2e-4) < 1e-9
    n95 = oracle.Net([a.flat for a in a95])
    _, _, ac5 = oracle.schedule(5)
    pe5 = oracle.embedding_table(5)
    tot = 0.0
    with torch.no_grad():
        for k in range(3):
            xs = dataset[k * 64:(k + 1) * 64]
            t5 = np.random.default_rng(10 + k).integers(1, 6, 64)
            e5 = np.random.default_rng(20 + k).standard_normal(xs.shape).astype(np.float32)
            tot += float(oracle.mse(oracle.unet_forward(n95, torch.tensor(oracle.q_sample(xs, t5, e5, ac5)), torch.tensor(pe5[t5 - 1])), torch.tensor(e5)))
    assert 0.18 < tot / 3 < 0.27, tot / 3


def test_golden_network(oracle, model_arrays, dataset, tabs):
    g = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))
    B = 8
    x0 = dataset[:B]
    ts = np.random.default_rng(1).integers(1, 501, B)
    eps = np.random.default_rng(2).standard_normal(x0.shape).astype(np.float32)
    xt = oracle.q_sample(x0, ts, eps, tabs["acum"])
    assert np.array_equal(xt.view(np.uint32), g["b8_xt_bits"])
    net = oracle.Net(model_arrays)
    with torch.no_grad():
        e_test = oracle.unet_forward(net, torch.tensor(xt), torch.tensor(tabs["pe"][ts - 1])).numpy()
        e_train = oracle.unet_forward(net, torch.tensor(xt), torch.tensor(tabs["pe"][ts - 1]), train=True).numpy()
    assert rel_l2(e_test, g["b8_eps_test"]) < 1e-5 and rel_l2(e_train, g["b8_eps_train"]) < 1e-5
    xT = np.random.default_rng(5).standard_normal((2, 1, 32, 32)).astype(np.float32)
    z = np.random.default_rng(6).standard_normal((5, 2, 1, 32, 32)).astype(np.float32)
    out = oracle.generate_image(oracle.Net(model_arrays), xT, z, tabs["acum"], tabs["pe"], t_start=6)
    assert np.abs(out - g["samp_t6"]).max() < 1e-4 and np.abs(out).max() <= 1.0


def test_golden_training(oracle, model_arrays, dataset, tabs):
    g = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))
    B = 8
    net = oracle.Net(model_arrays)
    opt = oracle.Adam(net.trainable(), eta=1e-4)
    losses = []
    for k in range(3):
        ts = np.random.default_rng(100 + k).integers(1, 501, B)
        eps = np.random.default_rng(200 + k).standard_normal((B, 1, 32, 32)).astype(np.float32)
        loss, grads = oracle.train_step(net, opt, dataset[k * B:(k + 1) * B], ts, eps, tabs["acum"], tabs["pe"])
        losses.append(loss)
        if k == 0:
            gn = np.array([np.linalg.norm(x) for x in grads])
            # conv biases that feed a train-mode BatchNorm have an exactly-zero gradient (BN removes the
            # mean); what Zygote/torch return there is rounding noise (~1e-7), so it is not compared
            sig = g["train3_grad_norms_step0"] > 1e-5
            assert sig.sum() == 34
            assert np.allclose(gn[sig], g["train3_grad_norms_step0"][sig], rtol=2e-3)
            assert (gn[~sig] < 1e-5).all()
    assert np.allclose(losses, g["train3_losses"], rtol=1e-4)


def test_backward_finite_differences(oracle, model_arrays, dataset, tabs):
    """fp64 finite differences of the oracle loss w.r.t. a few parameters of several arrays."""
    B = 4
    x0 = dataset[:B]
    ts = np.array([3, 77, 250, 499])
    eps = np.random.default_rng(0).standard_normal(x0.shape).astype(np.float32)
    net = oracle.Net(model_arrays, dtype=torch.float64)

    def loss_fn():
        l, _, _ = oracle.train_step_loss(net, x0, ts, eps, tabs["acum"], tabs["pe"], update_stats=False)
        return l

    l = loss_fn()
    l.backward()
    rng = np.random.default_rng(1)
    for k in (0, 1, 3, 6, 18, 36, 37, 50, 62, 63):   # conv W (incl. emb channels), bias, BN gamma, convT, final
        p = net.flat[k]
        for _ in range(2):
            j = int(rng.integers(0, p.numel()))
            h = 1e-6
            with torch.no_grad():
                old = float(p.view(-1)[j])
                p.view(-1)[j] = old + h
                lp = float(loss_fn())
                p.view(-1)[j] = old - h
                lm = float(loss_fn())
                p.view(-1)[j] = old
            fd = (lp - lm) / (2 * h)
            an = float(p.grad.view(-1)[j])
            assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)) + 2e-7, (k, j, fd, an)


def test_adam_first_step_is_eta_sign(oracle):
    p = torch.zeros(4, dtype=torch.float32)
    opt = oracle.Adam([p], eta=1e-4)
    opt.step([np.array([1.0, -2.0, 0.5, 0.0], np.float32)])
    # m/(1-b1) = g, sqrt(v/(1-b2)) = |g|  => p = -eta*sign(g) (up to eps)
    assert np.allclose(p.numpy(), [-1e-4, 1e-4, -1e-4, 0.0], atol=1e-9)
