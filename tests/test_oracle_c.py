"""The C restatement (oracle/ddpm_tables.c) and the NumPy restatement agree bit for bit (CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return C.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle_tables.so"))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.mark.parametrize("T", [5, 500, 1000])
def test_schedule_c_vs_numpy(clib, oracle, T):
    b, a, c = (np.empty(T, np.float32) for _ in range(3))
    clib.oracle_schedule(T, C.c_float(1e-4), C.c_float(0.02), _fp(b), _fp(a), _fp(c))
    bo, ao, co = oracle.schedule(T)
    assert np.array_equal(b.view(np.uint32), bo.view(np.uint32))
    assert np.array_equal(a.view(np.uint32), ao.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), co.view(np.uint32))


def test_embedding_and_scalars_c_vs_numpy(clib, oracle):
    _, _, acum = oracle.schedule(500)
    for t in (1, 2, 250, 499, 500):
        pe = np.empty(128, np.float32)
        clib.oracle_embedding(t, 128, _fp(pe))
        assert np.array_equal(pe.view(np.uint32), oracle.timestep_embedding(t).view(np.uint32))
        s = np.empty(4, np.float32)
        clib.oracle_sampler_scalars(_fp(acum), t, _fp(s))
        assert np.array_equal(s.view(np.uint32), np.array(oracle.sampler_scalars(acum, t), np.float32).view(np.uint32))


def test_q_sample_apply_noise_philox_c_vs_numpy(clib, oracle):
    _, _, acum = oracle.schedule(500)
    rng = np.random.default_rng(0)
    B, hw = 7, 1024
    x0 = rng.uniform(-3, 1.3, (B, 1, 32, 32)).astype(np.float32)
    eps = rng.standard_normal(x0.shape).astype(np.float32)
    ts = rng.integers(1, 501, B).astype(np.int32)
    out = np.empty_like(x0)
    clib.oracle_q_sample(_fp(x0), ts.ctypes.data_as(C.POINTER(C.c_int)), _fp(eps), _fp(acum), B, hw, _fp(out))
    assert np.array_equal(out.view(np.uint32), oracle.q_sample(x0, ts, eps, acum).view(np.uint32))
    img = rng.random((9, 11))
    e = rng.standard_normal((9, 11))
    o = np.empty_like(img)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    clib.oracle_apply_noise(dp(img), dp(e), C.c_long(img.size), 500, C.c_double(1e-4), C.c_double(0.02), dp(o))
    assert np.array_equal(o, oracle.apply_noise_f64(img, e))
    ctr = np.array([5, 77, 0, 123], np.uint32)
    key = np.array([42, 7], np.uint32)
    res = np.empty(4, np.uint32)
    up = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    clib.oracle_philox4x32_10(up(ctr), up(key), up(res))
    assert np.array_equal(res, oracle.philox4x32_10(ctr[None], key[None])[0])
