"""Host-side constants of the DDPM scripts: the beta schedule and the timestep embedding.

These are what the Julia host evaluates at load time (``const β = collect(range(β_min, β_max,
length=T))`` ... /root/reference/src/train_brain.jl:17-24, and ``timestep_embedding``
train_brain.jl:54-62) and hands to libddpm through ``ddpm_set_tables``; this is the Python
host's version of the same constants.  Host code, not the compute path.
"""
from __future__ import annotations

import math

import numpy as np

BETA_MIN = np.float32(1e-4)   # src/train_brain.jl:20
BETA_MAX = np.float32(0.02)   # src/train_brain.jl:21
T_DEFAULT = 500               # `const T = 5 #00` (train_brain.jl:18): 500 is the intended value
D_EMBED = 128                 # src/train_brain.jl:17


def beta_schedule(T: int = T_DEFAULT, beta_min=BETA_MIN, beta_max=BETA_MAX):
    """(beta, alpha, alpha_cum), all Float32.

    ``range(Float32, Float32, length=T)`` is a StepRangeLen{Float32,Float64,Float64}: element i
    is Float32(ref + i*step) evaluated in Float64 on the Float32-rounded end points;
    ``accumulate(*, α)`` is a sequential Float32 product."""
    b0, b1 = float(np.float32(beta_min)), float(np.float32(beta_max))
    i = np.arange(T, dtype=np.float64)
    beta = (b0 + i * ((b1 - b0) / (T - 1))).astype(np.float32) if T > 1 else np.array([b0], np.float32)
    alpha = (np.float32(1) - beta).astype(np.float32)
    acum = np.empty(T, np.float32)
    p = np.float32(1)
    for k in range(T):
        p = alpha[0] if k == 0 else np.float32(p * alpha[k])
        acum[k] = p
    return beta, alpha, acum


def timestep_embedding(t: int, D: int = D_EMBED) -> np.ndarray:
    """train_brain.jl:54-62 -- Float64 math on the Float32 ``log(1e4)``, stored as Float32,
    interleaved sin/cos, exponent denominator D-1."""
    pe = np.zeros(D, np.float32)
    neg_log = -float(np.log(np.float32(1e4)))
    for i in range(1, D // 2 + 1):
        div = math.exp(neg_log * (2 * (i - 1) / (D - 1)))
        pe[2 * i - 2] = np.float32(math.sin(t * div))
        pe[2 * i - 1] = np.float32(math.cos(t * div))
    return pe


def embedding_table(T: int = T_DEFAULT, D: int = D_EMBED) -> np.ndarray:
    """[T, D]; row t-1 is timestep_embedding(t)."""
    return np.stack([timestep_embedding(t, D) for t in range(1, T + 1)])


def apply_noise_betas(num_noise_steps: int = 500, beta_min: float = 1e-4, beta_max: float = 0.02) -> np.ndarray:
    """``beta_min:(beta_max-beta_min)/num_noise_steps:beta_max`` (Float64 StepRangeLen,
    /root/reference/src/ImageGenerationDiffusionModels.jl:62): num_noise_steps+1 values."""
    step = (beta_max - beta_min) / num_noise_steps
    return beta_min + np.arange(num_noise_steps + 1, dtype=np.float64) * step
