"""CPU statement of the numerical claim behind csrc/conv1_tc.cuh: a K = 9 dot product whose FP32 operands are both
split into two BF16 parts (x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, FP32 accumulation) stays ~2^-17 relative to the
magnitude of the terms, i.e. far below the FP16 rounding of the stored activation; one BF16 part alone would not."""
import numpy as np
import torch


def _bf16(v):
    return torch.tensor(np.asarray(v, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def test_two_part_bf16_split_is_16_bit_accurate():
    rng = np.random.default_rng(0)
    n = 100000
    x = (3.0 * rng.standard_normal((n, 9))).astype(np.float32)          # x_t values of the sampler
    w = (0.15 * rng.standard_normal((1, 9))).astype(np.float32)          # one output channel of the first conv
    exact = (x.astype(np.float64) * w.astype(np.float64)).sum(1)
    scale = (np.abs(x).astype(np.float64) * np.abs(w)).sum(1)            # magnitude of the terms being summed
    xh = _bf16(x); xl = _bf16(x - xh)
    wh = _bf16(w); wl = _bf16(w - wh)
    split = ((xh.astype(np.float64) * wh).sum(1) + (xl.astype(np.float64) * wh).sum(1)
             + (xh.astype(np.float64) * wl).sum(1)).astype(np.float32)
    hi_only = (xh.astype(np.float64) * wh).sum(1).astype(np.float32)
    err_split = np.abs(split - exact) / scale
    err_hi = np.abs(hi_only - exact) / scale
    assert err_split.max() < 2.0 ** -15, err_split.max()
    assert np.median(err_hi) > 50 * np.median(err_split)                # the low parts are what buys the accuracy
    # after the FP16 rounding of the stored activation the split result differs from the FP32 result in ~1 % of
    # the elements, by one rounding step or (small outputs after cancellation) by the 2^-15 * |terms| bound above
    a = np.maximum(split + np.float32(0.3), 0).astype(np.float16)
    b = np.maximum(exact + 0.3, 0).astype(np.float32).astype(np.float16)
    assert np.mean(a != b) < 0.02
    big = np.maximum(np.abs(a), np.abs(b)).astype(np.float16)
    ulp = np.maximum(np.spacing(big).astype(np.float32), (2.0 ** -15 * scale).astype(np.float32))
    assert (np.abs(a.astype(np.float32) - b.astype(np.float32)) <= 1.01 * ulp).all()
