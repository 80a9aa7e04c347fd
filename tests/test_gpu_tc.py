"""tcgen05 kernels against the CUDA-core kernels of the same library and against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, config2_batch, rel_l2

pytestmark = pytest.mark.gpu
OUT = os.path.join(ROOT, "gpurun_out")
NAMES = ["y1", "a1", "y2", "a2", "p1", "y3", "a3", "y4", "a4", "y5", "a5", "y6", "a6", "u", "y7", "a7", "y8", "a8",
         "y9", "a9", "y10", "a10"]


def _dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, name), "w") as fh:
        json.dump(obj, fh, indent=1, default=float)


def _layers(h, xt, ts):
    h.predict_eps(xt, ts, train_mode=True)
    return {nm: h.debug_fetch(nm) for nm in NAMES}


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
def test_tc_matches_simt_per_layer(gpu_handles, oracle, model_arrays, dataset, tabs, mode):
    h = gpu_handles[mode]
    if h.counter("tc_available") != 1:
        pytest.fail("tcgen05 path unavailable on this device (cuTensorMapEncodeTiled entry point missing)")
    h.set_weights(model_arrays)
    B = 9   # odd batch: the last 128-row tile of every layer is ragged
    x0, ts, eps = config2_batch(dataset, B)
    xt = oracle.q_sample(x0, ts, eps, tabs["acum"])
    report = {}
    try:
        h.set_option("conv_impl", 1)
        ref = _layers(h, xt, ts)
        e_ref = h.predict_eps(xt, ts, train_mode=False)
        # base_offset 1 = descriptor base-offset field set to (start>>7)&7 for row-shifted starts.
        # Measured on B200 (round 1): mode 0 is exact, mode 1 reads garbage -- the 128B swizzle XOR is
        # taken from the absolute shared-memory address bits, so a row-advanced start needs no base offset.
        for bo in (0,):
            h.set_option("conv_impl", 2)
            got = _layers(h, xt, ts)
            e_got = h.predict_eps(xt, ts, train_mode=False)
            report[f"base_offset_{bo}"] = {nm: rel_l2(got[nm], ref[nm]) for nm in NAMES}
            report[f"base_offset_{bo}"]["eps_test_mode"] = rel_l2(e_got, e_ref)
    finally:
        h.set_option("conv_impl", 0)
        _dump(f"tc_vs_simt_{mode}.json", report)
    tol = 3e-3 if mode == "fp16" else 2e-2     # same inputs, different accumulation order + re-rounding
    bad = {k: v for k, v in report["base_offset_0"].items() if not (v <= tol)}
    assert not bad, bad


def test_tc_is_the_default_path(gpu_handles):
    h = gpu_handles["fp16"]
    assert h.counter("uses_tc") == 1
    assert gpu_handles["fp32"].counter("uses_tc") == 0
    assert gpu_handles["tf32"].counter("uses_tc") == 1


def test_fused_final_epilogue_matches_unfused(gpu_handles, model_arrays):
    """Sampler: last conv + final 1x1 conv + reverse update in one tcgen05 epilogue vs separate kernels.
    The fused path feeds the un-rounded FP32 activations of layer 10 into the 1x1 conv, so they agree to
    FP16 rounding of a10 only."""
    h = gpu_handles["fp16"]
    h.set_weights(model_arrays)
    xT = np.random.default_rng(0).standard_normal((5, 1, 32, 32)).astype(np.float32)
    z = np.random.default_rng(1).standard_normal((9, 5, 1, 32, 32)).astype(np.float32)
    try:
        h.set_option("fuse_final", 0)
        a = h.sample(5, x_T=xT, z=z, t_start=10)
        h.set_option("fuse_final", 1)
        b = h.sample(5, x_T=xT, z=z, t_start=10)
    finally:
        h.set_option("fuse_final", 1)
    assert np.abs(a - b).max() < 5e-3 and np.abs(a - b).mean() < 2e-4
    assert np.abs(b).max() <= 1.0


def test_cta_pair_convs_match_single_cta_convs(gpu_handles, oracle, model_arrays, dataset, tabs):
    """cta_group::2 kernels (two CTAs share one M=256 MMA stream, each staging half of the weight rows) against the
    single-CTA kernels: same K order per output element, so test-mode outputs and the sampler are bit-identical;
    train mode differs only through the atomically accumulated BatchNorm statistics (last-ulp noise)."""
    h = gpu_handles["fp16"]
    h.set_weights(model_arrays)
    rng = np.random.default_rng(3)
    xT = rng.standard_normal((5, 1, 32, 32)).astype(np.float32)
    z = rng.standard_normal((9, 5, 1, 32, 32)).astype(np.float32)
    rep = {}
    try:
        for B in (2, 9, 37):          # 2 and 9: the last pair-tile of every layer is ragged / half empty
            x0, ts, eps = config2_batch(dataset, B)
            xt = oracle.q_sample(x0, ts, eps, tabs["acum"])
            h.set_option("tc_pair", 0)
            e0 = h.predict_eps(xt, ts, train_mode=False)
            l0, g0 = h.loss_and_grad(x0, ts, eps)
            _, g0b = h.loss_and_grad(x0, ts, eps)     # same kernels again: run-to-run noise floor of train mode
            h.set_option("tc_pair", 31)
            e1 = h.predict_eps(xt, ts, train_mode=False)
            l1, g1 = h.loss_and_grad(x0, ts, eps)
            ks = (6, 12, 18, 24, 30, 36, 50, 56)
            rep[f"B{B}"] = {"eps_max_abs_diff": float(np.abs(e0 - e1).max()), "loss": [l0, l1],
                            "grad_rel": {k: rel_l2(g1[k], g0[k]) for k in ks},
                            "grad_rel_same_kernels_twice": {k: rel_l2(g0b[k], g0[k]) for k in ks}}
        h.set_option("tc_pair", 0)
        a = h.sample(5, x_T=xT, z=z, t_start=10)
        h.set_option("tc_pair", 31)
        b = h.sample(5, x_T=xT, z=z, t_start=10)
        rep["sampler_max_abs_diff"] = float(np.abs(a - b).max())
    finally:
        h.set_option("tc_pair", 31)
        _dump("pair_vs_single.json", rep)
    for B in (2, 9, 37):
        r = rep[f"B{B}"]
        assert r["eps_max_abs_diff"] == 0.0, rep
        assert abs(r["loss"][0] - r["loss"][1]) <= 1e-4 * abs(r["loss"][0]), rep
        # train mode is not run-to-run reproducible (atomically accumulated BatchNorm statistics flip FP16 roundings
        # and ReLU masks); the pair kernels must stay within a small multiple of that noise floor
        floor = max(r["grad_rel_same_kernels_twice"].values())
        assert max(r["grad_rel"].values()) < max(3.0 * floor, 1e-2), rep
    assert rep["sampler_max_abs_diff"] == 0.0, rep


@pytest.mark.parametrize("mode", [2, 1])
def test_first_conv_on_tensor_cores_matches_cuda_core_kernel(gpu_handles, model_arrays, mode):
    """conv1_tc.cuh (BF16 hi/lo split operands on tcgen05, FP32 accumulate; mode 2 = default: the timestep constants ride
    on the contraction as three BF16 parts, mode 1: added in the epilogue) against the FP32 CUDA-core first conv on
    the same device-resident input: the stored FP16 activation may differ by one rounding step in a few elements
    (the split keeps 16 significand bits per operand), never by more; the sampler output moves by < 1e-3."""
    h = gpu_handles["fp16"]
    h.set_weights(model_arrays)
    n = 37                                   # ragged last tile, images straddling tile boundaries
    rng = np.random.default_rng(5)
    xT = (2.0 * rng.standard_normal((n, 1, 32, 32))).astype(np.float32)
    z = rng.standard_normal((5, n, 1, 32, 32)).astype(np.float32)
    rep = {}
    try:
        out = {}
        for m in (0, mode):
            h.set_option("conv1_tc", m)
            out[m] = h.sample(n, x_T=xT, z=z, t_start=6)
        # the layer alone: ddpm_time_kernel("conv1") refills the set's x with the same Philox normals on every call and
        # leaves the first conv's output in the (otherwise aliased) a1 buffer
        h.set_option("conv1_tc", 0); h.time_kernel("conv1", n, 1); ref = h.debug_fetch("infer:a1")
        h.set_option("conv1_tc", mode); h.time_kernel("conv1", n, 1); got = h.debug_fetch("infer:a1")
        d = np.abs(got - ref)
        # one FP16 rounding step of the reference value, floored at 2e-5 absolute (values that straddle the ReLU zero)
        ulp = np.maximum(2.0 ** (np.floor(np.log2(np.maximum(np.abs(ref), 2.0 ** -14))) - 10), 2e-5)
        rep = {"frac_mismatch": float(np.mean(d > 0)), "max_err_ulp": float((d / ulp).max()),
               "max_abs_a1": float(np.abs(ref).max()), "frac_nonzero": float(np.mean(ref != 0)),
               "sampler_max_abs_diff": float(np.abs(out[0] - out[mode]).max())}
    finally:
        h.set_option("conv1_tc", 2)
        _dump(f"conv1_tc{mode}_vs_simt.json", rep)
    assert rep["max_abs_a1"] > 0.1 and rep["frac_nonzero"] > 0.2, rep     # the layer really ran on real data
    # measured: 0.27 % of the elements differ, all by exactly one rounding step; the bound allows the second step that
    # a small output after cancellation may take (tests/test_split_arithmetic.py)
    assert rep["frac_mismatch"] < 0.03 and rep["max_err_ulp"] <= 2.0, rep
    assert rep["sampler_max_abs_diff"] < 1e-3, rep


def test_tf32_mode_matches_fp32_mode_per_layer(gpu_handles, oracle, model_arrays, dataset, tabs):
    """DDPM_PREC_TF32: the 3x3 convolutions on tcgen05 kind::tf32 (FP32 tensors in HBM, 32-channel K chunks, Cout split over
    blockIdx.y) against the CUDA-core FP32 mode of the same library, every intermediate tensor of a train-mode forward,
    the sampler and the data gradients (B = 9: ragged last tile)."""
    hf, ht = gpu_handles["fp32"], gpu_handles["tf32"]
    for h in (hf, ht):
        h.set_weights(model_arrays)
    B = 9
    x0, ts, eps = config2_batch(dataset, B)
    xt = oracle.q_sample(x0, ts, eps, tabs["acum"])
    ref, got = _layers(hf, xt, ts), _layers(ht, xt, ts)
    rep = {nm: rel_l2(got[nm], ref[nm]) for nm in NAMES}
    lf, gf = hf.loss_and_grad(x0, ts, eps)
    lt, gt = ht.loss_and_grad(x0, ts, eps)
    rep["loss_rel"] = abs(lf - lt) / lf
    rep["grads"] = {k: rel_l2(gt[k], gf[k]) for k in (0, 6, 12, 18, 24, 30, 36, 38, 44, 50, 56, 62)}
    xT = np.random.default_rng(0).standard_normal((3, 1, 32, 32)).astype(np.float32)
    z = np.random.default_rng(1).standard_normal((9, 3, 1, 32, 32)).astype(np.float32)
    rep["sampler_max_abs_diff"] = float(np.abs(hf.sample(3, x_T=xT, z=z, t_start=10) - ht.sample(3, x_T=xT, z=z, t_start=10)).max())
    _dump("tf32_vs_fp32.json", rep)
    assert max(rep[nm] for nm in NAMES) < 3e-3, rep          # TF32 operand rounding: 2^-11 relative per product
    assert rep["loss_rel"] < 3e-4 and max(rep["grads"].values()) < 8e-2, rep      # measured 6e-5 / 3.2e-2 (B = 9)
    assert rep["sampler_max_abs_diff"] < 2e-2, rep


def test_inference_is_bitwise_repeatable(gpu_handles, model_arrays):
    """compute-sanitizer is closed on this pool (it left GPUs needing a reset), so data races in the mbarrier / TMEM /
    cluster pipelines are hunted the cheap way: the inference path has no atomics, so every kernel must be bit-for-bit
    repeatable -- 12 evaluations of a ragged batch and 3 runs of a 20-step sampler must agree exactly."""
    h = gpu_handles["fp16"]
    h.set_weights(model_arrays)
    rng = np.random.default_rng(8)
    x = rng.standard_normal((37, 1, 32, 32)).astype(np.float32)
    ts = rng.integers(1, 501, 37)
    first = h.predict_eps(x, ts)
    for _ in range(11):
        assert np.array_equal(h.predict_eps(x, ts), first)
    shared = h.predict_eps(x, np.full(37, 250))
    for _ in range(5):
        assert np.array_equal(h.predict_eps(x, np.full(37, 250)), shared)
    a = h.sample(300, seed=3, first_index=0, t_start=21)
    for _ in range(2):
        assert np.array_equal(h.sample(300, seed=3, first_index=0, t_start=21), a)
