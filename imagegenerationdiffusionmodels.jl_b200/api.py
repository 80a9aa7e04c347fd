"""Host mirror of the reference package's public API on top of libddpm.

README.md:16-30,47 of the reference documents ``generate_grid``, ``apply_noise``, ``train``,
``denoise_image``, ``generate_image`` and ``demo``; the arithmetic behind them lives in the two
driver scripts (src/train_brain.jl, src/generate_images.jl).  Here the same names keep the same
arguments, defaults, side effects (PNG files, BSON checkpoints) and error behaviour, while every
array operation of the DDPM path is one call into the C ABI.  The Julia version of this file is
``julia/src/ImageGenerationDiffusionModels.jl``; this one exists because the image has no Julia.
"""
from __future__ import annotations

import os
import struct
import zlib
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import bson_io, capi, tables

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = os.path.join(REPO_ROOT, "fixtures")
DEFAULT_DATA = os.path.join(FIXTURES, "SyntheticImages500.mat")
DEFAULT_MODEL = os.path.join(FIXTURES, "trained_model.bson")


# ----------------------------------------------------------------------------- small host utilities
def save_png(path: str, img01: np.ndarray):
    """``save(path, colorview(Gray, img))`` for an image already clamped to [0,1]."""
    a = np.clip(np.asarray(img01, dtype=np.float64), 0.0, 1.0)
    a8 = np.round(a * 255.0).astype(np.uint8)
    h, w = a8.shape
    raw = b"".join(b"\x00" + a8[r].tobytes() for r in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as fh:
        fh.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0))
                 + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def load_dataset(path: str = DEFAULT_DATA) -> np.ndarray:
    """``matread(path)["syntheticImages"]`` -> [N,1,32,32] Float32 (Julia (32,32,1,N) column-major
    == NumPy [n][c][j][i]); values as stored in the file (no rescale)."""
    from scipy.io import loadmat

    raw = loadmat(path)["syntheticImages"]
    raw = raw.reshape(raw.shape[0], raw.shape[1], 1, -1) if raw.ndim == 3 else raw
    return np.ascontiguousarray(np.transpose(raw, (3, 2, 1, 0)), dtype=np.float32)


@dataclass
class SimpleUNet:
    """Parameter container with the field layout of the reference struct
    (src/train_brain.jl:89-96): 64 Float32 arrays in BSON order."""

    arrays: List[np.ndarray]
    eta: float = 1e-4
    template: str = DEFAULT_MODEL

    @staticmethod
    def load(path: str = DEFAULT_MODEL) -> "SimpleUNet":
        """``@load path model`` (src/generate_images.jl:250)."""
        arrs, meta = bson_io.load_checkpoint(path)
        return SimpleUNet([a.flat for a in arrs], eta=meta.get("eta") or 1e-4, template=path)

    @staticmethod
    def init(seed: Optional[int] = None) -> "SimpleUNet":
        """``SimpleUNet(1)`` (src/train_brain.jl:109-145) with Flux's default initialisation:
        glorot_uniform conv weights, zero biases, BatchNorm gamma=1, beta=0, mu=0, var=1."""
        rng = np.random.default_rng(seed)
        arrays = []
        for dims in bson_io.expected_array_dims():
            if len(dims) == 4:
                k1, k2, c3, c4 = dims
                fan = k1 * k2 * (c3 + c4)
                lim = np.sqrt(6.0 / fan)
                arrays.append(rng.uniform(-lim, lim, size=int(np.prod(dims))).astype(np.float32))
            else:
                arrays.append(np.zeros(dims[0], np.float32))
        # BatchNorm blocks: (beta, gamma, mu, var) -> gamma = var = 1
        k = 0
        for kind, _ in bson_io.UNET_LAYERS:
            if kind == "bn":
                arrays[k + 1][:] = 1.0
                arrays[k + 3][:] = 1.0
                k += 4
            else:
                k += 2
        return SimpleUNet(arrays)

    def save(self, path: str, epoch: Optional[int] = None):
        """``@save path model opt [epoch]`` (src/train_brain.jl:295-300): same BSON document
        structure as the reference's files."""
        bson_io.save_checkpoint(path, self.template, self.arrays, epoch=epoch, eta=self.eta)


# ----------------------------------------------------------------------------- engine cache
_engines = {}


def engine(T: int = tables.T_DEFAULT, precision: int = capi.PREC_FP16, device: int = 0) -> capi.Handle:
    """One libddpm handle per (T, precision, device), tables supplied by the host."""
    key = (T, precision, device)
    if key not in _engines:
        h = capi.Handle(T=T, D=tables.D_EMBED, precision=precision, device=device)
        beta, _, acum = tables.beta_schedule(T)
        h.set_tables(beta, acum, tables.embedding_table(T))
        _engines[key] = h
    return _engines[key]


_model: Optional[SimpleUNet] = None


def default_model() -> SimpleUNet:
    global _model
    if _model is None:
        _model = SimpleUNet.load(DEFAULT_MODEL)
    return _model


# ----------------------------------------------------------------------------- public API
def generate_grid(data_path: str = DEFAULT_DATA, out_path: str = "grid.png") -> np.ndarray:
    """8x8 canvas of the first 64 dataset images, written to ``grid.png``
    (/root/reference/src/ImageGenerationDiffusionModels.jl:25-43).  Returns the 256x256 canvas
    (Julia canvas[row, col] with image i*8+j at block row i, block column j)."""
    imgs = load_dataset(data_path)[:64, 0]            # [n][j][i]
    canvas = np.zeros((8 * 32, 8 * 32), np.float32)   # canvas[r][c] == Julia canvas[r+1, c+1]
    for i in range(8):
        for j in range(8):
            # Julia block .= first64[:, :, idx] is indexed [first dim, second dim] = [i_w, j_h]
            canvas[i * 32:(i + 1) * 32, j * 32:(j + 1) * 32] = imgs[i * 8 + j].T
    save_png(out_path, np.clip(canvas, 0, 1))
    return canvas


def apply_noise(img, num_noise_steps: int = 500, beta_min: float = 0.0001, beta_max: float = 0.02,
                epsilon: Optional[np.ndarray] = None, out_path: Optional[str] = "noisy_img.png",
                rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """Forward noising (/root/reference/src/ImageGenerationDiffusionModels.jl:60-73): Float64,
    one ``epsilon = randn(size(img))`` reused for every beta of the 501-value schedule; writes
    ``noisy_img.png`` (clamped) and returns the unclamped Float64 array.  ``epsilon`` may be
    supplied for reproducibility (the reference draws it unseeded)."""
    img = np.asarray(img, dtype=np.float64)
    if epsilon is None:
        epsilon = (rng or np.random.default_rng()).standard_normal(img.shape)
    out = capi.apply_noise_f64(img, epsilon, tables.apply_noise_betas(num_noise_steps, beta_min, beta_max))
    if out_path and out.ndim == 2:
        save_png(out_path, np.clip(out, 0, 1))
    return out


def generate_image(model: Optional[SimpleUNet] = None, num_images: int = 1, image_size=(32, 32),
                   T: int = tables.T_DEFAULT, seed: int = 0, x_T: Optional[np.ndarray] = None,
                   z: Optional[np.ndarray] = None, precision: int = capi.PREC_FP16, device: int = 0,
                   first_index: int = 0) -> np.ndarray:
    """``generate_image(model; num_images, image_size)`` (src/generate_images.jl:231-245):
    x_T ~ N(0,1), ``for t in reverse(2:T)`` reverse_diffusion, final clamp.  Returns [N,1,H,W]
    (== Julia H x W x 1 x N).  Noise comes from the device generator unless x_T / z are given."""
    if tuple(image_size) != (32, 32):
        raise ValueError("the trained U-Net only supports 32x32 images")
    model = model or default_model()
    h = engine(T, precision, device)
    h.set_weights(model.arrays)
    return h.sample(num_images, x_T=x_T, z=z, seed=seed, first_index=first_index, t_start=T)


def denoise_image(noisy_img, model: Optional[SimpleUNet] = None, t_start: int = 100, T: int = tables.T_DEFAULT,
                  seed: int = 0, out_path: Optional[str] = "denoised_img.png", precision: int = capi.PREC_FP16,
                  device: int = 0) -> np.ndarray:
    """``denoise_image(noisy_img)``: the reverse loop started from a supplied 32x32 image at
    level ``t_start`` (README.md:26).  Keeps the reference's contract
    (/root/reference/src/ImageGenerationDiffusionModels.jl:90-98, test/runtests.jl:23-29):
    returns a 32x32 array clamped to [0,1] and writes ``denoised_img.png``."""
    x = np.asarray(noisy_img, dtype=np.float32)
    if x.shape != (32, 32):
        raise ValueError("denoise_image expects a 32x32 matrix")
    model = model or default_model()
    h = engine(T, precision, device)
    h.set_weights(model.arrays)
    # data space of the network is 2*img-1 (train_brain.jl:250-251); Julia matrix [i,j] -> [j][i]
    xs = (2.0 * x.T - 1.0).astype(np.float32).reshape(1, 1, 32, 32)
    out = h.sample(1, x_T=xs, seed=seed, t_start=t_start)[0, 0]
    den = np.clip((out.T + 1.0) / 2.0, 0.0, 1.0)
    if out_path:
        save_png(out_path, den)
    return den


@dataclass
class TrainResult:
    model: SimpleUNet
    losses: List[float] = field(default_factory=list)
    step_losses: List[float] = field(default_factory=list)
    stopped_early: bool = False


def train(data=DEFAULT_DATA, lr: float = 1e-4, epochs: int = 100, patience: int = 10, min_delta: float = 0.001,
          batch_size: int = 64, T: int = tables.T_DEFAULT, model: Optional[SimpleUNet] = None,
          rng: Optional[np.random.Generator] = None, schedule=None, save_dir: Optional[str] = ".",
          precision: int = capi.PREC_FP16, device: int = 0, log=print, resume_from: Optional[str] = None,
          save_optimizer_state: bool = False) -> TrainResult:
    """``train(data, lr, epochs, patience, min_delta)`` (README.md:23) == ``main`` of
    src/train_brain.jl:246-304: load, rescale ``imgs .*= 2; imgs .-= 1``, Adam(lr), epochs of
    shuffled mini-batches (last batch 52 of 500), early stopping, BSON checkpoint every 5 epochs
    and ``trained_model.bson`` at the end.

    Beyond the reference (which saves the Adam RULE only and cannot resume):
    ``save_optimizer_state`` also writes the moments next to every checkpoint (``<name>.adam.bson``) and
    ``resume_from=<checkpoint .bson>`` continues from that file's weights and, if present, its moment file.

    ``schedule(epoch, n_images)`` may return host-supplied draws
    ``(perm, [ts per batch], [eps per batch])`` for parity runs; otherwise ``rng`` draws them
    (the reference uses Julia's unseeded task-local RNG)."""
    imgs = load_dataset(data) if isinstance(data, str) else np.asarray(data, dtype=np.float32)
    imgs = (imgs.reshape(-1, 1, 32, 32) * np.float32(2) - np.float32(1)).astype(np.float32)
    n = imgs.shape[0]
    rng = rng or np.random.default_rng()
    if resume_from is not None:
        model = SimpleUNet.load(resume_from)
    model = model or SimpleUNet.init()
    h = engine(T, precision, device)
    h.set_weights(model.arrays)
    h.set_adam(float(np.float32(lr)), 0.9, 0.999, 1e-8)
    if resume_from is not None and os.path.exists(_adam_path(resume_from)):
        m, v, bt, steps, _ = bson_io.load_adam_state(_adam_path(resume_from))
        h.set_adam_state(m, v, bt, steps)
    model.eta = float(np.float32(lr))        # `opt = Adam(lr)` is what @save writes next to the model
    res = TrainResult(model)
    best, no_improve = float("inf"), 0
    for epoch in range(1, epochs + 1):
        if schedule is not None:
            perm, ts_list, eps_list = schedule(epoch, n)
        else:
            perm, ts_list, eps_list = rng.permutation(n), None, None
        total, nb = np.float32(0), 0
        for bi, i0 in enumerate(range(0, n, batch_size)):
            sel = perm[i0:i0 + batch_size]
            x0 = imgs[sel]
            B = len(sel)
            ts = ts_list[bi] if ts_list is not None else rng.integers(1, T + 1, B)
            eps = eps_list[bi] if eps_list is not None else rng.standard_normal(x0.shape).astype(np.float32)
            loss = h.train_step(x0, ts, eps)
            res.step_losses.append(loss)
            total = np.float32(total + np.float32(loss))
            nb += 1
        epoch_loss = float(total / np.float32(nb))
        res.losses.append(epoch_loss)
        if log:
            log(f"Epoch {epoch} | avg loss = {epoch_loss}")
        if epoch_loss < best - min_delta:
            best, no_improve = epoch_loss, 0
        else:
            no_improve += 1
        if no_improve > patience:
            if log:
                log(f"Early stopping: No significant improvement for {patience + 1} epochs")
            res.stopped_early = True
            break
        if save_dir is not None and epoch % 5 == 0:
            model.arrays = h.get_weights()
            _save(model, h, os.path.join(save_dir, f"ddpm_epoch_{epoch}.bson"), epoch, save_optimizer_state)
    model.arrays = h.get_weights()
    if save_dir is not None:
        _save(model, h, os.path.join(save_dir, "trained_model.bson"), None, save_optimizer_state)
    return res


def _adam_path(checkpoint: str) -> str:
    return (checkpoint[:-5] if checkpoint.endswith(".bson") else checkpoint) + ".adam.bson"


def _save(model: SimpleUNet, h, path: str, epoch, with_optimizer_state: bool):
    model.save(path, epoch=epoch)
    if with_optimizer_state:
        m, v, bt, steps = h.get_adam_state()
        bson_io.save_adam_state(_adam_path(path), m, v, bt, steps, model.eta)


def demo(out_dir: str = ".", seed: int = 0) -> dict:
    """``demo()`` (README.md:47-49): grid -> noise -> denoise -> generate, all but ``train``."""
    img = generate_grid(out_path=os.path.join(out_dir, "grid.png"))
    first = img[:32, :32]
    noisy = apply_noise(first, out_path=os.path.join(out_dir, "noisy_img.png"), rng=np.random.default_rng(seed))
    den = denoise_image(np.clip(first, 0, 1), out_path=os.path.join(out_dir, "denoised_img.png"), seed=seed)
    new = generate_image(num_images=1, seed=seed)
    save_png(os.path.join(out_dir, "generated_image_1.png"), (new[0, 0].T + 1.0) / 2.0)
    return {"grid": img, "noisy": noisy, "denoised": den, "generated": new}
