import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import igdm_b200  # noqa: E402,F401  registers the package


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


FIX = os.path.join(ROOT, "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def oracle():
    import ddpm_oracle
    return ddpm_oracle


@pytest.fixture(scope="session")
def model_arrays():
    from igdm_b200 import bson_io
    arrs, _ = bson_io.load_checkpoint(os.path.join(FIX, "trained_model.bson"))
    return [a.flat for a in arrs]


@pytest.fixture(scope="session")
def dataset():
    """[500,1,32,32] Float32 after the training rescale imgs*2-1 (train_brain.jl:250-251)."""
    from igdm_b200 import api
    return (api.load_dataset() * np.float32(2) - np.float32(1)).astype(np.float32)


@pytest.fixture(scope="session")
def tabs():
    from igdm_b200 import tables
    beta, alpha, acum = tables.beta_schedule(500)
    return {"beta": beta, "alpha": alpha, "acum": acum, "pe": tables.embedding_table(500)}


def config2_batch(dataset, B=64):
    """first B images, ts = default_rng(1), eps = default_rng(2)."""
    x0 = dataset[:B]
    ts = np.random.default_rng(1).integers(1, 501, B)
    eps = np.random.default_rng(2).standard_normal(x0.shape).astype(np.float32)
    return x0, ts, eps


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="session")
def gpu_handles(model_arrays, tabs):
    """One handle per precision mode, tables supplied by the host."""
    from igdm_b200 import capi
    if capi.device_count() < 1:
        pytest.skip("no CUDA device")
    hs = {}
    for name, prec in (("fp32", capi.PREC_FP32), ("fp16", capi.PREC_FP16), ("bf16", capi.PREC_BF16), ("tf32", capi.PREC_TF32)):
        h = capi.Handle(T=500, precision=prec)
        h.set_tables(tabs["beta"], tabs["acum"], tabs["pe"])
        h.set_weights(model_arrays)
        hs[name] = h
    yield hs
    for h in hs.values():
        h.close()
