"""CPU ORACLE for the DDPM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product (libddpm.so and
the host mirror under ``imagegenerationdiffusionmodels.jl_b200/``) never does.

PARITY UNPINNED: the reference (Julia + Flux 0.16.4 / NNlib 0.9.30 / Zygote 0.7.10 /
Optimisers 0.4.6, pinned in /root/reference/last_desperate_attempt/Manifest.toml)
cannot run in this image (no Julia) and its own test-suite holds no numeric golden
vector for this path (/root/reference/test/runtests.jl:1-51 asserts shapes and file
existence only).  What pins this restatement instead (tests/test_oracle_*.py):
  * NUMBERS THE REFERENCE'S OWN RUN PRODUCED: the BatchNorm running means / variances Flux
    wrote into the shipped checkpoints (20 vectors over all 10 layers, three checkpoints)
    are reproduced by this forward on the same dataset to 1-7 % -- a statistical pin of the
    data scaling, schedule, embedding layout, conv / pool / ConvTranspose / concat structure
    and BatchNorm semantics; mutated semantics miss by 25-130 %
    (tests/test_oracle_checkpoint_stats.py, which also lists what this cannot see);
  * the shipped checkpoints reproduce the published loss curve only under these
    semantics (ddpm_epoch_95 @T=5 -> eps-MSE ~0.22; trained_model @T=500 -> ~0.10);
  * closed-form anchors: schedule / embedding bit patterns,
    apply_noise == 0.079302*x + 14.892430*eps, embedding of t=0 == (0,1,0,1,..);
  * two independent conv implementations (torch cross-correlation on flipped weights
    vs. an index-by-index NumPy/C true convolution written from NNlib's definition);
  * fp64 finite-difference checks of the backward pass.

Every function cites the reference lines it restates.  Arrays use NumPy row-major
views of Julia's column-major data: Julia x[i,j,c,n] (W,H,C,N) == x[n][c][j][i].
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

try:  # torch is only the fp32 CPU conv/autograd engine of the oracle
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None

f32 = np.float32

# =============================================================================
# 1. Schedule  (src/train_brain.jl:20-24; duplicate src/generate_images.jl:14-18)
# =============================================================================


def _truncbits_f32(x: np.float32, nb: int) -> np.float32:
    """Base.truncbits: zero the ``nb`` low mantissa bits (twiceprecision.jl)."""
    u = np.array([x], dtype=np.float32).view(np.uint32)
    u &= np.uint32((0xFFFFFFFF << nb) & 0xFFFFFFFF)
    return u.view(np.float32)[0]


def _add12_f32(x: np.float32, y: np.float32):
    if abs(y) > abs(x):
        x, y = y, x
    h = f32(x + y)
    return h, f32(f32(x - h) + y)


def julia_range_f32(start: float, stop: float, length: int) -> np.ndarray:
    """``collect(range(Float32(start), Float32(stop), length=length))``.

    Restates Base ``range_start_stop_length`` -> ``_linspace`` (twiceprecision.jl,
    Julia 1.11): for these endpoints the rational shortcut is rejected
    (``rat(1f-4)`` = 0//1, and Float32(0/50) != 1f-4), so the twice-precision
    fallback builds a ``StepRangeLen{Float32,Float64,Float64}`` whose element i is
    ``Float32(ref + (i-offset)*step)`` evaluated in Float64."""
    start, stop = f32(start), f32(stop)
    n = int(length)
    if n == 1:
        return np.array([start], dtype=f32)
    delta = f32(stop - start)
    tmin = f32(-f32(start / delta))
    imin = int(np.round(f32(tmin * f32(n - 1) + f32(1))))
    if 1 < imin < n:
        t = f32(f32(imin - 1) / f32(n - 1))
        ref = f32(f32(f32(1) - t) * start + t * stop)
        step = f32(f32(ref - start) / f32(imin - 1)) if imin - 1 < n - imin else f32(f32(stop - ref) / f32(n - imin))
    elif imin <= 1:
        imin, ref, step = 1, start, f32(delta / f32(n - 1))
    else:
        imin, ref, step = n, stop, f32(delta / f32(n - 1))
    big = max(imin - 1, n - imin)
    nb = min(12, 0 if n < 2 else int(math.ceil(math.log2(big))))  # nbitslen(Float32, len, offset)
    step_hi = _truncbits_f32(step, nb)
    x1_hi, x1_lo = _add12_f32(f32(f32(1 - imin) * step_hi), ref)
    x2_hi, x2_lo = _add12_f32(f32(f32(n - imin) * step_hi), ref)
    a = f32(f32(start - x1_hi) - x1_lo)
    b = f32(f32(stop - x2_hi) - x2_lo)
    step_lo = f32(f32(b - a) / f32(n - 1))
    ref_lo = f32(a - f32(f32(1 - imin) * step_lo))
    ref64 = float(ref) + float(ref_lo)
    step64 = float(step_hi) + float(step_lo)
    idx = np.arange(1, n + 1, dtype=np.float64)
    return (ref64 + (idx - imin) * step64).astype(f32)


def schedule(T: int = 500, beta_min: float = 1e-4, beta_max: float = 0.02):
    """beta, alpha, alpha_cum exactly as the script computes them (all Float32):
    ``β = collect(range(β_min, β_max, length=T)); α = 1 .- β; α_cum = accumulate(*, α)``
    (src/train_brain.jl:20-24).  ``accumulate`` is a sequential left fold in Float32."""
    beta = julia_range_f32(f32(beta_min), f32(beta_max), T)
    alpha = (f32(1) - beta).astype(f32)
    acum = np.empty(T, dtype=f32)
    p = f32(1)
    for i in range(T):
        p = f32(p * alpha[i]) if i else alpha[0]
        acum[i] = p
    return beta, alpha, acum


# =============================================================================
# 2. Timestep embedding  (src/train_brain.jl:54-62; src/generate_images.jl:147-155)
# =============================================================================


def timestep_embedding(t: int, D: int = 128) -> np.ndarray:
    """``div = exp(-log(Float32(1e4)) * (2*(i-1)/(D-1)))``: the log is Float32, promoted to
    Float64 by the Float64 quotient; sin/cos in Float64; stored to a Float32 vector,
    interleaved [sin_1, cos_1, sin_2, cos_2, ...]."""
    pe = np.zeros(D, dtype=f32)
    neg_log = -float(np.log(f32(1e4)))  # Float32 log, widened
    for i in range(1, D // 2 + 1):
        div = math.exp(neg_log * (2 * (i - 1) / (D - 1)))
        pe[2 * i - 2] = f32(math.sin(t * div))
        pe[2 * i - 1] = f32(math.cos(t * div))
    return pe


def embedding_table(T: int = 500, D: int = 128) -> np.ndarray:
    """Rows t=1..T (row index t-1)."""
    return np.stack([timestep_embedding(t, D) for t in range(1, T + 1)]).astype(f32)


# =============================================================================
# 3. Forward noising
# =============================================================================


def q_sample(x0: np.ndarray, ts: Sequence[int], eps: np.ndarray, alpha_cum: np.ndarray) -> np.ndarray:
    """``x_t = a .* x0 .+ b .* ϵ`` with a=sqrt.(ᾱ[ts]), b=sqrt.(1 .- ᾱ[ts])
    (src/train_brain.jl:230-233).  Float32 throughout; ``ts`` 1-based.  Julia's fused
    broadcast evaluates a*x0 + b*eps per element (muladd is NOT implied), so two
    roundings of the products and one of the sum."""
    ts = np.asarray(ts, dtype=np.int64)
    ac = alpha_cum[ts - 1].astype(f32)
    a = np.sqrt(ac).astype(f32).reshape(-1, 1, 1, 1)
    b = np.sqrt((f32(1) - ac).astype(f32)).astype(f32).reshape(-1, 1, 1, 1)
    x0 = np.asarray(x0, dtype=f32)
    eps = np.asarray(eps, dtype=f32)
    return ((a * x0).astype(f32) + (b * eps).astype(f32)).astype(f32)


def apply_noise_f64(img: np.ndarray, eps: np.ndarray, num_noise_steps: int = 500,
                    beta_min: float = 1e-4, beta_max: float = 0.02) -> np.ndarray:
    """``apply_noise`` (src/ImageGenerationDiffusionModels.jl:60-73) on a host-supplied eps:
    ``for beta in beta_min:(beta_max-beta_min)/num_noise_steps:beta_max;
    img = sqrt(1-beta).*img + sqrt(beta).*epsilon``  -- Float64, SAME eps every step,
    num_noise_steps+1 betas.  The Float64 StepRange is restated as start + k*step (the
    twice-precision range reproduces the nominal decimal values to the last bit or one
    ulp; the recurrence is insensitive at 1e-13)."""
    img = np.asarray(img, dtype=np.float64).copy()
    eps = np.asarray(eps, dtype=np.float64)
    step = (beta_max - beta_min) / num_noise_steps
    for k in range(num_noise_steps + 1):
        beta = beta_min + k * step
        img = math.sqrt(1 - beta) * img + math.sqrt(beta) * eps
    return img


def apply_noise_coeffs(num_noise_steps: int = 500, beta_min: float = 1e-4, beta_max: float = 0.02):
    """Closed form of the recurrence above: out = A*img + B*eps."""
    A, B = 1.0, 0.0
    step = (beta_max - beta_min) / num_noise_steps
    for k in range(num_noise_steps + 1):
        beta = beta_min + k * step
        A, B = math.sqrt(1 - beta) * A, math.sqrt(1 - beta) * B + math.sqrt(beta)
    return A, B


# =============================================================================
# 4. Weights: Flux layouts -> cross-correlation layouts
# =============================================================================

# (kind, cin, cout) in constructor order (src/train_brain.jl:109-145)
LAYERS = [
    ("conv", 129, 64), ("bn", 64), ("conv", 64, 64), ("bn", 64),
    ("conv", 64, 128), ("bn", 128), ("conv", 128, 128), ("bn", 128),
    ("conv", 128, 128), ("bn", 128), ("conv", 128, 128), ("bn", 128),
    ("convT", 128, 64), ("conv", 64, 64), ("bn", 64), ("conv", 64, 64), ("bn", 64),
    ("conv", 128, 64), ("bn", 64), ("conv", 64, 64), ("bn", 64),
    ("conv1x1", 64, 1),
]
BN_EPS = f32(1e-5)       # Flux BatchNorm ϵ (read from the BSON)
BN_MOMENTUM = f32(0.1)


def array_lengths() -> List[int]:
    """Lengths of the 64 arrays in ABI/BSON order."""
    out = []
    for kind, ci, co in (l if len(l) == 3 else (l[0], l[1], l[1]) for l in LAYERS):
        if kind == "conv":
            out += [9 * ci * co, co]
        elif kind == "conv1x1":
            out += [ci * co, co]
        elif kind == "convT":
            out += [4 * ci * co, co]
        else:
            out += [ci] * 4
    return out


def trainable_mask() -> List[bool]:
    """True for arrays Adam updates (conv W,b; BN β,γ), False for BN μ,σ²."""
    out = []
    for l in LAYERS:
        out += [True, True, False, False] if l[0] == "bn" else [True, True]
    return out


class Net:
    """Parameter container mirroring ``SimpleUNet`` (src/train_brain.jl:89-96).

    conv weights are kept as torch tensors in *cross-correlation* OIHW layout derived from
    the Flux arrays: Julia w[a,b,ci,co] (column-major) == row-major [co][ci][b][a]; Flux
    ``Conv`` is a true convolution, i.e. cross-correlation with both spatial axes flipped
    (NNlib ``conv`` with flipped=false)."""

    def __init__(self, arrays: Sequence[np.ndarray], dtype=None):
        dtype = dtype or torch.float32
        lens = array_lengths()
        assert len(arrays) == 64, len(arrays)
        self.dtype = dtype
        self.flat = [torch.tensor(np.asarray(a, dtype=np.float32).reshape(-1), dtype=dtype) for a in arrays]
        for a, n in zip(self.flat, lens):
            assert a.numel() == n, (a.numel(), n)
        mask = trainable_mask()
        for a, m in zip(self.flat, mask):
            a.requires_grad_(m)

    def arrays(self) -> List[np.ndarray]:
        return [a.detach().to(torch.float32).numpy().copy() for a in self.flat]

    def trainable(self):
        return [a for a, m in zip(self.flat, trainable_mask()) if m]


def _conv_w(flat, ci, co, k):
    # Julia (k,k,ci,co) col-major -> [co][ci][k2][k1]; flip both spatial axes -> cross-correlation
    return flat.reshape(co, ci, k, k).flip(2, 3)


def _convT_w(flat, ci, co):
    # Julia (2,2,co,ci) col-major -> [ci][co][b][a] == torch conv_transpose2d layout, flipped (B2)
    return flat.reshape(ci, co, 2, 2).flip(2, 3)


def _bn(x, beta, gamma, mu, var, train: bool, new_stats: Optional[list]):
    """Flux BatchNorm(c, relu).  train: batch μ, biased σ² (two-pass),
    running-stat update with the unbiased variance; test: running stats."""
    C = x.shape[1]
    if train:
        m = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(dim=(0, 2, 3))
        v = ((x - mean.view(1, C, 1, 1)) ** 2).mean(dim=(0, 2, 3))
        if new_stats is not None:
            mom = float(BN_MOMENTUM)
            with torch.no_grad():
                new_mu = (1 - mom) * mu + mom * mean
                new_var = (1 - mom) * var + mom * (m / (m - 1)) * v
            new_stats.append((new_mu.detach(), new_var.detach()))
    else:
        mean, v = mu, var
    scale = gamma / torch.sqrt(v + float(BN_EPS))
    shift = beta - scale * mean
    return torch.relu(x * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1))


def unet_forward(net: Net, x, pe, train: bool = False, update_stats: bool = False,
                 act_round=None, taps: Optional[dict] = None):
    """``(m::SimpleUNet)((x, t_emb))`` (src/train_brain.jl:159-179).

    x  : [B,1,32,32] (== Julia 32x32x1xB);  pe : [B,128] embedding rows.
    ``act_round`` optionally rounds conv *inputs* (activations and weights) to a narrower
    format (emulation of a reduced-precision device path; None == exact fp32 oracle).
    ``taps`` (dict) collects named intermediates for per-layer parity tests."""
    r = (lambda t: t) if act_round is None else act_round
    p = net.flat
    B, _, H, W = x.shape
    new_stats: list = []
    ns = new_stats if (train and update_stats) else None
    i = 0

    def conv(h, ci, co, first=False):
        nonlocal i
        w = _conv_w(p[i], ci, co, 3)
        b = p[i + 1]
        i += 2
        if first and act_round is not None:
            # device path keeps the image channel and the embedding fold in fp32
            return F.conv2d(h, w, b, padding=1)
        return F.conv2d(r(h), r(w), b, padding=1)

    def bn(h):
        nonlocal i
        out = _bn(h, p[i], p[i + 1], p[i + 2], p[i + 3], train, ns)
        i += 4
        return out

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    # tmap = repeat(reshape(t_emb,1,1,:,B), H, W, 1, 1); cat(x, tmap; dims=3)  (:164-168)
    tmap = pe.view(B, -1, 1, 1).expand(B, pe.shape[1], H, W)
    h = torch.cat([x, tmap], dim=1)
    h = tap("a1", bn(tap("y1", conv(h, 129, 64, first=True))))
    h1 = tap("h1", bn(tap("y2", conv(h, 64, 64))))
    h = tap("p1", F.max_pool2d(h1, 2))                               # MaxPool((2,2)) (:117)
    h = tap("a3", bn(tap("y3", conv(h, 64, 128))))
    h = tap("a4", bn(tap("y4", conv(h, 128, 128))))
    h = tap("a5", bn(tap("y5", conv(h, 128, 128))))
    h = tap("a6", bn(tap("y6", conv(h, 128, 128))))
    wt, bt = _convT_w(p[i], 128, 64), p[i + 1]
    i += 2
    h = tap("u", F.conv_transpose2d(r(h), r(wt), bt, stride=2))     # ConvTranspose((2,2),128=>64,stride=2) (:130)
    h = tap("a7", bn(tap("y7", conv(h, 64, 64))))
    h = tap("a8", bn(tap("y8", conv(h, 64, 64))))
    h = torch.cat([h, h1], dim=1)                                    # cat(up_h3, h1_c; dims=3) (:175)
    h = tap("a9", bn(tap("y9", conv(h, 128, 64))))
    h = tap("a10", bn(tap("y10", conv(h, 64, 64))))
    wf, bf = p[i].reshape(1, 64, 1, 1), p[i + 1]
    i += 2
    out = F.conv2d(h, wf, bf)                                        # Conv((1,1),64=>1) (:142), kept fp32 on device
    assert i == 64
    if ns is not None:
        return out, new_stats
    return out


def apply_new_stats(net: Net, new_stats):
    """Write the running statistics a train-mode forward produced back into the net
    (Flux mutates μ, σ² in place during the forward)."""
    k = 0
    idx = 0
    for l in LAYERS:
        if l[0] == "bn":
            mu, var = new_stats[k]
            with torch.no_grad():
                net.flat[idx + 2].copy_(mu)
                net.flat[idx + 3].copy_(var)
            k += 1
            idx += 4
        else:
            idx += 2


def mse(pred, target):
    """Flux.Losses.mse = mean(abs2.(ŷ .- y)) (src/train_brain.jl:240)."""
    return ((pred - target) ** 2).mean()


def train_step_loss(net: Net, x0, ts, eps, alpha_cum, pe_table, update_stats=True, act_round=None):
    """``train_step`` (src/train_brain.jl:225-241) on host-supplied ts (1-based) and eps."""
    xt = torch.tensor(q_sample(x0, ts, eps, alpha_cum)).to(net.dtype)
    pe = torch.tensor(pe_table[np.asarray(ts) - 1]).to(net.dtype)
    res = unet_forward(net, xt, pe, train=True, update_stats=update_stats, act_round=act_round)
    if update_stats:
        pred, ns = res
    else:
        pred, ns = res, None
    return mse(pred, torch.tensor(np.asarray(eps, dtype=np.float32)).to(net.dtype)), pred, ns


class Adam:
    """Optimisers.Adam (0.4.6) as ``Flux.setup``/``update!`` apply it:
    m = β1 m + (1-β1) g;  v = β2 v + (1-β2) g²;  p -= η (m/(1-β1^t)) / (sqrt(v/(1-β2^t)) + ε);
    βt <- βt .* β afterwards.  All in the array eltype (Float32)."""

    def __init__(self, params, eta=1e-4, beta=(0.9, 0.999), eps=1e-8):
        self.params = list(params)
        self.eta, self.b1, self.b2, self.eps = f32(eta), f32(beta[0]), f32(beta[1]), f32(eps)
        self.m = [np.zeros(p.numel(), dtype=f32) for p in self.params]
        self.v = [np.zeros(p.numel(), dtype=f32) for p in self.params]
        self.bt1, self.bt2 = self.b1, self.b2

    def step(self, grads):
        one = f32(1)
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            g = np.asarray(g, dtype=f32).reshape(-1)
            m[:] = self.b1 * m + (one - self.b1) * g
            v[:] = self.b2 * v + (one - self.b2) * (g * g)
            upd = (m / (one - self.bt1)) / (np.sqrt(v / (one - self.bt2)) + self.eps) * self.eta
            with torch.no_grad():
                p -= torch.tensor(upd.astype(f32)).to(p.dtype).view_as(p)
        self.bt1 = f32(self.bt1 * self.b1)
        self.bt2 = f32(self.bt2 * self.b2)


def train_step(net: Net, opt: Adam, x0, ts, eps, alpha_cum, pe_table):
    """One iteration of the inner loop of ``main`` (src/train_brain.jl:265-274):
    withgradient -> update! ; returns (loss, grads as list of np arrays)."""
    for p in net.trainable():
        p.grad = None
    loss, _, ns = train_step_loss(net, x0, ts, eps, alpha_cum, pe_table, update_stats=True)
    loss.backward()
    grads = [p.grad.detach().to(torch.float32).numpy().reshape(-1).copy() for p in net.trainable()]
    opt.step(grads)
    apply_new_stats(net, ns)
    return float(loss.detach()), grads


# =============================================================================
# 5. Reverse process  (src/generate_images.jl:174-245)
# =============================================================================


def sampler_scalars(alpha_cum: np.ndarray, t: int):
    """Float32 scalars of one ``reverse_diffusion`` call, computed exactly in the order the
    source writes them (src/generate_images.jl:186-208).  t is 1-based, t_prev = t-1.
    Returns (sigma_t, sqrt_alpha_t, sqrt_alpha_prev, sqrt_post_var)."""
    a_t = f32(alpha_cum[t - 1])
    a_prev = f32(alpha_cum[t - 2]) if t > 1 else f32(1)
    beta_t = f32(f32(1) - a_t)
    beta_prev = f32(f32(1) - a_prev)
    sigma_t = f32(np.sqrt(beta_t))
    post_var = f32(f32(beta_prev * f32(f32(1) - a_t)) / f32(f32(1) - a_t))
    return sigma_t, f32(np.sqrt(a_t)), f32(np.sqrt(a_prev)), f32(np.sqrt(post_var))


def sampler_table(alpha_cum: np.ndarray) -> np.ndarray:
    """[T,4] Float32; row t-1 holds the scalars for step t (row 0 is the unreachable t=1 branch)."""
    T = len(alpha_cum)
    return np.array([sampler_scalars(alpha_cum, t) for t in range(1, T + 1)], dtype=f32)


def reverse_update(x_t: np.ndarray, eps_pred: np.ndarray, z: np.ndarray, scal) -> np.ndarray:
    """Elementwise tail of ``reverse_diffusion`` (src/generate_images.jl:196-208), Float32:
    pred_x0 = clamp((x_t - σ_t*ϵ̂)/sqrt(ᾱ_t), -1, 1); x_prev = sqrt(ᾱ_prev)*pred_x0 + sqrt(pv)*z."""
    s, sa, sp, spv = (f32(v) for v in scal)
    x0 = ((x_t - (s * eps_pred).astype(f32)).astype(f32) / sa).astype(f32)
    x0 = np.clip(x0, f32(-1), f32(1))
    return ((sp * x0).astype(f32) + (spv * z).astype(f32)).astype(f32)


def generate_image(net: Net, x_T: np.ndarray, z: np.ndarray, alpha_cum, pe_table,
                   t_start: Optional[int] = None, act_round=None) -> np.ndarray:
    """``generate_image`` (src/generate_images.jl:231-245) on host-supplied noise.
    x_T: [N,1,32,32]; z: [t_start-1, N,1,32,32] with z[k] used at step t = t_start-k.
    Loop ``for t in reverse(2:T)``, test-mode BN, final clamp to [-1,1]."""
    T = len(alpha_cum)
    t_start = T if t_start is None else t_start
    x = np.asarray(x_T, dtype=f32).copy()
    N = x.shape[0]
    with torch.no_grad():
        for k, t in enumerate(range(t_start, 1, -1)):
            pe = torch.tensor(pe_table[t - 1]).view(1, -1).expand(N, -1).to(net.dtype)
            e = unet_forward(net, torch.tensor(x).to(net.dtype), pe, train=False, act_round=act_round)
            e = e.to(torch.float32).numpy()
            x = reverse_update(x, e, z[k], sampler_scalars(alpha_cum, t))
    return np.clip(x, f32(-1), f32(1))


# =============================================================================
# 6. Index-by-index restatement of the layers (small cases; validates section 4's mapping)
# =============================================================================


def conv3x3_true_numpy(x: np.ndarray, w_julia_flat: np.ndarray, b: np.ndarray, ci: int, co: int) -> np.ndarray:
    """NNlib ``conv`` (true convolution, pad=1) written from its definition with Julia indices
    (SURVEY.md Appendix B1): y[i,j,co,n] = b[co] + Σ_{a,b,ci} w[a,b,ci,co]·x[i+2-a, j+2-b, ci, n].
    x: [N,ci,H,W] row-major view (x[n][c][j][i]);  w flat in Julia column-major order."""
    N, _, H, W = x.shape
    w = np.asarray(w_julia_flat, dtype=np.float64).reshape(co, ci, 3, 3)  # [co][ci][b][a]
    xp = np.zeros((N, ci, H + 2, W + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1] = x
    y = np.zeros((N, co, H, W), dtype=np.float64)
    for a in range(1, 4):          # first (fastest, "i"/W) kernel index
        for bb in range(1, 4):     # second ("j"/H) kernel index
            # x[i+2-a, j+2-b] with 1-based i,j -> padded 0-based index (i-1)+(2-a)+1 = i+2-a
            xs = xp[:, :, (3 - bb):(3 - bb) + H, (3 - a):(3 - a) + W]
            y += np.einsum("nchw,oc->nohw", xs, w[:, :, bb - 1, a - 1])
    return y + np.asarray(b, dtype=np.float64).reshape(1, co, 1, 1)


def convT2x2_numpy(x: np.ndarray, w_julia_flat: np.ndarray, b: np.ndarray, ci: int, co: int) -> np.ndarray:
    """Flux ConvTranspose((2,2), ci=>co, stride=2) == ∇conv_data of the flipped-kernel conv
    (SURVEY.md Appendix B2), 0-based: out[2i+1-a, 2j+1-b, co, n] += w[a,b,co,ci]·x[i,j,ci,n]."""
    N, _, H, W = x.shape
    w = np.asarray(w_julia_flat, dtype=np.float64).reshape(ci, co, 2, 2)  # [ci][co][b][a]
    out = np.zeros((N, co, 2 * H, 2 * W), dtype=np.float64)
    for a in range(2):
        for bb in range(2):
            out[:, :, (1 - bb)::2, (1 - a)::2] += np.einsum("nchw,co->nohw", x.astype(np.float64), w[:, :, bb, a])
    return out + np.asarray(b, dtype=np.float64).reshape(1, co, 1, 1)


# =============================================================================
# 7. Counter-based RNG of the device sampler (integer stream bit-exact)
# =============================================================================

_PHILOX_M0, _PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """Philox-4x32-10 (Salmon et al. 2011).  counter: [...,4] uint32, key: [...,2] uint32."""
    c = np.asarray(counter, dtype=np.uint64).copy()
    k = np.asarray(key, dtype=np.uint64).copy()
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_PHILOX_M0) * c[..., 0]
        p1 = np.uint64(_PHILOX_M1) * c[..., 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = (hi1 ^ c[..., 1] ^ k[..., 0]) & mask
        n2 = (hi0 ^ c[..., 3] ^ k[..., 1]) & mask
        c = np.stack([n0, lo1, n2, lo0], axis=-1)
        k = np.stack([(k[..., 0] + np.uint64(_PHILOX_W0)) & mask, (k[..., 1] + np.uint64(_PHILOX_W1)) & mask], axis=-1)
    return c.astype(np.uint32)


def device_normal(seed: int, image_index: np.ndarray, step: int, n_pix: int = 1024) -> np.ndarray:
    """The N(0,1) draws libddpm's in-kernel generator produces for (seed, global image
    index, step): one Philox call per 4 consecutive pixels, counter = (pixel_quad,
    image_index lo, image_index hi, step), key = (seed lo, seed hi); Box-Muller on
    u = (r + 0.5) * 2^-32:  z0 = sqrt(-2 ln u0) cos(2π u1), z1 = .. sin(..), same for (u2,u3).
    Integer stream is bit-exact with the device; the floats agree to ~1e-6."""
    image_index = np.asarray(image_index, dtype=np.uint64).reshape(-1)
    quads = np.arange(n_pix // 4, dtype=np.uint64)
    ctr = np.zeros((len(image_index), len(quads), 4), dtype=np.uint64)
    ctr[..., 0] = quads[None, :]
    ctr[..., 1] = (image_index & np.uint64(0xFFFFFFFF))[:, None]
    ctr[..., 2] = (image_index >> np.uint64(32))[:, None]
    ctr[..., 3] = np.uint64(step)
    key = np.zeros(ctr.shape[:-1] + (2,), dtype=np.uint64)
    key[..., 0] = np.uint64(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint64((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key).astype(np.float64)
    u = (r + 0.5) * (2.0 ** -32)
    rad0 = np.sqrt(-2.0 * np.log(u[..., 0]))
    rad1 = np.sqrt(-2.0 * np.log(u[..., 2]))
    ang0 = 2.0 * np.pi * u[..., 1]
    ang1 = 2.0 * np.pi * u[..., 3]
    z = np.stack([rad0 * np.cos(ang0), rad0 * np.sin(ang0), rad1 * np.cos(ang1), rad1 * np.sin(ang1)], axis=-1)
    return z.reshape(len(image_index), n_pix).astype(f32)
