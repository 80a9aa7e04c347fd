"""Host-side logic and the C-ABI surface (CPU only: no compute calls)."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import FIX, ROOT


def test_bson_roundtrip_is_byte_exact(tmp_path):
    from igdm_b200 import bson_io
    src = os.path.join(FIX, "trained_model.bson")
    arrs, meta = bson_io.load_checkpoint(src)
    assert len(arrs) == 64 and sum(a.flat.size for a in arrs) == 848961
    assert abs(meta["eta"] - 1e-4) < 1e-9 and meta["beta"] == (0.9, 0.999) and meta["eps"] == 1e-8
    assert [a.dims for a in arrs] == bson_io.expected_array_dims()
    out = tmp_path / "rt.bson"
    bson_io.save_checkpoint(str(out), src, arrs)
    assert hashlib.md5(out.read_bytes()).hexdigest() == hashlib.md5(open(src, "rb").read()).hexdigest()
    # patched payloads survive a reload
    new = [a.flat + np.float32(1) for a in arrs]
    bson_io.save_checkpoint(str(out), src, new, epoch=7)
    back, _ = bson_io.load_checkpoint(str(out))
    assert all(np.array_equal(b.flat, n) for b, n in zip(back, new))


def test_dataset_layout():
    from igdm_b200 import api
    d = api.load_dataset()
    assert d.shape == (500, 1, 32, 32) and d.dtype == np.float32
    assert abs(float(d.mean()) + 0.7190) < 1e-3 and abs(float(d.min()) + 1.0676) < 1e-3


def test_simpleunet_init_matches_flux_defaults():
    from igdm_b200 import api
    m = api.SimpleUNet.init(seed=0)
    assert len(m.arrays) == 64
    w = m.arrays[0]
    lim = np.sqrt(6.0 / (9 * (129 + 64)))
    assert np.abs(w).max() <= lim and np.abs(w).max() > 0.9 * lim
    assert np.all(m.arrays[1] == 0)                       # conv bias
    assert np.all(m.arrays[2] == 0) and np.all(m.arrays[3] == 1)   # BN beta, gamma
    assert np.all(m.arrays[4] == 0) and np.all(m.arrays[5] == 1)   # BN mu, var


def test_generate_grid(tmp_path):
    from igdm_b200 import api
    c = api.generate_grid(out_path=str(tmp_path / "grid.png"))
    assert c.shape == (256, 256) and (tmp_path / "grid.png").stat().st_size > 1000
    d = api.load_dataset()
    assert np.array_equal(c[32:64, 64:96], d[1 * 8 + 2, 0].T)


def test_shard_range_covers_everything():
    from igdm_b200 import dist
    for n in (0, 1, 7, 64, 65536, 65537):
        for w in (1, 2, 3, 4, 8):
            rs = [dist.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dist.shard_range(4, 4, 4)


def test_library_exports_every_declared_symbol():
    from igdm_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "libddpm.h")).read()
    declared = set(re.findall(r"\b(ddpm_[a-z0-9_]+)\s*\(", hdr)) - {"ddpm_handle"}
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    lib = capi.load_library()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ddpm_version() == 100
    lens = capi.array_lengths()
    assert len(lens) == 64 and sum(lens) == 848961


def test_no_gpu_means_loud_failure_not_fallback():
    from igdm_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DDPMError, match="no CUDA device"):
        capi.Handle()
    with pytest.raises(capi.DDPMError, match="no CUDA device"):
        capi.apply_noise_f64(np.zeros((2, 2)), np.zeros((2, 2)), np.array([0.1]))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "imagegenerationdiffusionmodels.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "ddpm_oracle" not in txt and "oracle/" not in txt, os.path.join(dirpath, f)


def test_every_option_key_is_documented_in_the_header():
    """ddpm_set_option keys accepted by the library (csrc/libddpm.cu) == keys listed in include/libddpm.h."""
    src = open(os.path.join(ROOT, "imagegenerationdiffusionmodels.jl_b200", "csrc", "libddpm.cu")).read()
    body = src[src.index("int ddpm_set_option("):src.index("int64_t ddpm_get_counter(")]
    accepted = set(re.findall(r'k == "([a-z0-9_]+)"', body))
    hdr = open(os.path.join(ROOT, "include", "libddpm.h")).read()
    doc = hdr[hdr.index("Option keys"):hdr.index("int ddpm_set_option(")]
    documented = set(re.findall(r"\b([a-z0-9]+(?:_[a-z0-9]+)+)\b", doc)) & (accepted | {"x"})
    assert accepted and accepted == documented, accepted ^ documented
