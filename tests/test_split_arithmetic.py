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


def test_folded_constant_rides_on_the_contraction_exactly():
    """conv1f_tc_kernel (csrc/conv1_tc.cuh, second version): the K = 64 rows the kernel builds, restated in NumPy.
       A row = [(x_hi, x_lo) x 9 taps | x_hi x 9 | onehot(cls) x 3 | 0 x 10]
       B row = [(w_hi, w_hi) x 9      | w_lo x 9 | E_hi(9 cls), E_mid(9), E_lo(9) | 0 x 10]
    Claims: (1) every entry is BF16-representable; (2) the three BF16 parts of E[cls] reproduce the FP32 value exactly, so the
    contraction equals  split-conv + E[cls]  with no extra error; (3) a halo row (all-zero A row) contracts to exactly 0;
    (4) rounding then ReLU on the packed 16-bit value equals ReLU then rounding (the epilogue does the former)."""
    rng = np.random.default_rng(1)
    n = 20000
    x = (3.0 * rng.standard_normal((n, 9))).astype(np.float32)
    w = (0.15 * rng.standard_normal(9)).astype(np.float32)
    E = (2.0 * rng.standard_normal(9)).astype(np.float32)               # E[cls][co]*scale + shift for one output channel
    cls = rng.integers(0, 9, n)
    xh = _bf16(x); xl = _bf16(x - xh)
    wh = _bf16(w); wl = _bf16(w - wh)
    e0 = _bf16(E); r1 = (E - e0).astype(np.float32); e1 = _bf16(r1); e2 = _bf16((r1 - e1).astype(np.float32))
    assert np.array_equal((e0.astype(np.float64) + e1 + e2).astype(np.float32), E)          # (2): 24 significand bits
    A = np.zeros((n, 64), np.float32)
    A[:, 0:18:2] = xh; A[:, 1:18:2] = xl; A[:, 18:27] = xh
    onehot = np.eye(9, dtype=np.float32)[cls]
    A[:, 27:36] = onehot; A[:, 36:45] = onehot; A[:, 45:54] = onehot
    B = np.zeros(64, np.float32)
    B[0:18:2] = wh; B[1:18:2] = wh; B[18:27] = wl; B[27:36] = e0; B[36:45] = e1; B[45:54] = e2
    assert np.array_equal(_bf16(A), A) and np.array_equal(_bf16(B), B)                        # (1)
    got = (A.astype(np.float64) * B).sum(1)
    want = ((xh.astype(np.float64) * wh).sum(1) + (xl.astype(np.float64) * wh).sum(1) + (xh.astype(np.float64) * wl).sum(1)
            + E[cls].astype(np.float64))
    assert np.allclose(got, want, rtol=0, atol=1e-12)                                         # (2) in exact arithmetic
    assert (np.zeros(64, np.float32) * B).sum() == 0.0                                        # (3)
    v = got.astype(np.float32)
    assert np.array_equal(np.maximum(v.astype(np.float16), np.float16(0)), np.maximum(v, 0).astype(np.float16))   # (4)
