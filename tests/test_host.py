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


# ------------------------------------------------------------------------------ checkpoints, resume, epoch loop
def test_checkpoint_carries_the_live_rule(tmp_path):
    """`@save path model opt [epoch]` writes the optimiser rule in use and the current epoch
    (/root/reference/src/train_brain.jl:295-300), not whatever the template file held."""
    from igdm_b200 import api, bson_io
    m = api.SimpleUNet.load(os.path.join(FIX, "ddpm_epoch_95.bson"))       # template: eta 2e-4, epoch 95
    assert abs(m.eta - 2e-4) < 1e-9
    m.eta = 3e-4
    out = str(tmp_path / "final.bson")
    m.save(out)                                                            # trained_model.bson carries no epoch
    _, meta = bson_io.load_checkpoint(out)
    assert abs(meta["eta"] - 3e-4) < 1e-9 and meta["epoch"] is None
    m.save(out, epoch=10)
    _, meta = bson_io.load_checkpoint(out)
    assert meta["epoch"] == 10 and abs(meta["eta"] - 3e-4) < 1e-9


def test_adam_state_file_roundtrip(tmp_path):
    from igdm_b200 import bson_io, capi
    lens = [n for n in map(int, __import__("numpy").array(capi.array_lengths()))]
    rng = np.random.default_rng(0)
    m = [rng.standard_normal(n).astype(np.float32) for n in lens]
    v = [rng.random(n).astype(np.float32) for n in lens]
    p = str(tmp_path / "x.adam.bson")
    bson_io.save_adam_state(p, m, v, (0.9 ** 4, 0.999 ** 4), 3, 1e-4)
    M, V, bt, steps, eta = bson_io.load_adam_state(p)
    assert all(np.array_equal(a, b) for a, b in zip(m, M)) and all(np.array_equal(a, b) for a, b in zip(v, V))
    assert steps == 3 and abs(bt[0] - 0.9 ** 4) < 1e-7 and abs(eta - 1e-4) < 1e-12


class _FakeEngine:
    """Stands in for capi.Handle in the epoch-loop test: scripted losses, no GPU."""

    def __init__(self, losses, arrays):
        self.losses, self.arrays, self.k, self.adam = list(losses), arrays, 0, None

    def set_weights(self, a):
        self.arrays = [np.array(x, copy=True) for x in a]

    def get_weights(self):
        return self.arrays

    def set_adam(self, *a):
        self.adam = a

    def train_step(self, x0, ts, eps):
        assert x0.shape[0] == len(ts) == eps.shape[0] and x0.shape[1:] == (1, 32, 32)
        assert ts.min() >= 1 and ts.max() <= 500
        v = self.losses[min(self.k, len(self.losses) - 1)]
        self.k += 1
        return v

    def get_adam_state(self):
        z = [np.zeros_like(a) for a in self.arrays]
        return z, z, (0.9, 0.999), self.k


def test_epoch_loop_early_stopping_and_checkpoints(tmp_path, monkeypatch):
    """The host loop of /root/reference/src/train_brain.jl:263-300: mean loss per epoch over 7x64 + 52 images,
    stop when no epoch improved by more than min_delta for more than `patience` epochs, checkpoint every 5 epochs,
    final trained_model.bson."""
    from igdm_b200 import api
    model = api.SimpleUNet.load()
    per_epoch = [1.0, 0.8, 0.6, 0.5] + [0.4995] * 40          # improvement stalls after epoch 4
    fake = _FakeEngine([l for l in per_epoch for _ in range(8)], model.arrays)
    monkeypatch.setattr(api, "engine", lambda *a, **k: fake)
    logs = []
    res = api.train(epochs=30, patience=3, min_delta=0.001, save_dir=str(tmp_path), log=logs.append, model=model,
                    rng=np.random.default_rng(0), lr=2e-4, save_optimizer_state=True)
    # epochs 5,6,7,8 do not improve on 0.5 by > 1e-3: no_improve exceeds patience=3 at epoch 8
    assert res.stopped_early and len(res.losses) == 8 and len(res.step_losses) == 64
    assert abs(res.losses[0] - 1.0) < 1e-6 and abs(res.losses[3] - 0.5) < 1e-6
    assert any("Early stopping" in l for l in logs) and logs[0].startswith("Epoch 1 | avg loss = ")
    assert sorted(f for f in os.listdir(tmp_path) if f.endswith(".bson")) == [
        "ddpm_epoch_5.adam.bson", "ddpm_epoch_5.bson", "trained_model.adam.bson", "trained_model.bson"]
    from igdm_b200 import bson_io
    _, meta = bson_io.load_checkpoint(str(tmp_path / "ddpm_epoch_5.bson"))
    assert meta["epoch"] == 5 and abs(meta["eta"] - 2e-4) < 1e-9 and abs(fake.adam[0] - 2e-4) < 1e-9
    _, meta = bson_io.load_checkpoint(str(tmp_path / "trained_model.bson"))
    assert meta["epoch"] is None
    # batches: 7 x 64 + 52 per epoch
    assert fake.k == 64


# ------------------------------------------------------------------------------ the Julia binding vs the header
_C2JL = {
    "ddpm_handle*": {"Ptr{Cvoid}"}, "ddpm_handle**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "const float*": {"Ptr{Float32}"}, "float*": {"Ptr{Float32}", "Ref{Cfloat}"},
    "const float* const*": {"Ptr{Ptr{Float32}}"}, "float* const*": {"Ptr{Ptr{Float32}}"},
    "const int32_t*": {"Ptr{Int32}"}, "const int64_t*": {"Ptr{Int64}"}, "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "const double*": {"Ptr{Float64}"}, "double*": {"Ptr{Float64}"}, "uint8_t*": {"Ptr{UInt8}"},
    "void*": {"Ptr{UInt8}", "Ptr{Cvoid}"}, "const void*": {"Ptr{UInt8}", "Ptr{Cvoid}"}, "const char*": {"Cstring"},
    "int": {"Cint"}, "int64_t": {"Int64"}, "uint64_t": {"UInt64"}, "float": {"Cfloat"}, "void": set(),
}


def _header_signatures():
    hdr = open(os.path.join(ROOT, "include", "libddpm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    sigs = {}
    for ret, name, args in re.findall(r"\b(const char\*|int64_t|int)\s+(ddpm_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        types = []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            if a == "void":
                continue
            a = re.sub(r"\s+", " ", a)
            m = re.match(r"^(.*?[\*\s])([A-Za-z_][A-Za-z0-9_]*)$", a)      # strip the parameter name
            t = (m.group(1) if m and m.group(2) not in ("int", "float", "int64_t", "uint64_t") else a).strip()
            types.append(re.sub(r"\s*\*", "*", t).replace("* const", "* const"))
        sigs[name] = (ret, types)
    return sigs


def test_julia_ccalls_match_the_header():
    """Every `ccall` in julia/src/LibDDPM.jl must name a declared entry point with the declared arity and
    argument types (the Julia glue cannot be executed in this image, so it is checked statically)."""
    jl = open(os.path.join(ROOT, "imagegenerationdiffusionmodels.jl_b200", "julia", "src", "LibDDPM.jl")).read()
    sigs = _header_signatures()
    assert len(sigs) >= 25 and "ddpm_train_step" in sigs
    calls = re.findall(r"ccall\(\(:(ddpm_[a-z0-9_]+),\s*libddpm\),\s*(\w+),\s*\(([^)]*)\)", jl, flags=re.S)
    assert len(calls) >= 15
    ret_map = {"int": "Cint", "const char*": "Cstring", "int64_t": "Int64"}
    for name, ret, args in calls:
        assert name in sigs, f"{name} is not declared in include/libddpm.h"
        cret, ctypes_ = sigs[name]
        assert ret == ret_map[cret], (name, ret, cret)
        jl_types = [a.strip() for a in args.replace("\n", " ").split(",") if a.strip()]
        assert len(jl_types) == len(ctypes_), (name, jl_types, ctypes_)
        for jt, ct in zip(jl_types, ctypes_):
            assert ct in _C2JL, (name, ct)
            assert jt in _C2JL[ct], f"{name}: Julia passes {jt} where the header declares {ct}"
