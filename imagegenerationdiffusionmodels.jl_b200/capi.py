"""ctypes binding of libddpm.so -- the same signatures the Julia host binds with ``ccall``
(include/libddpm.h; julia/src/LibDDPM.jl).  No torch types, no fallbacks: if the shared library
is missing or there is no GPU, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libddpm.so")

PREC_FP32, PREC_FP16, PREC_BF16, PREC_TF32 = 0, 1, 2, 3
NUM_ARRAYS = 64

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_pp = C.POINTER(_f32p)

# name -> (restype, argtypes): exactly the declarations of include/libddpm.h
SIGNATURES = {
    "ddpm_last_error": (C.c_char_p, []),
    "ddpm_version": (C.c_int, []),
    "ddpm_device_count": (C.c_int, []),
    "ddpm_array_lengths": (C.c_int, [_i64p]),
    "ddpm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ddpm_destroy": (C.c_int, [C.c_void_p]),
    "ddpm_set_tables": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
    "ddpm_get_tables": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, _f32p]),
    "ddpm_set_weights": (C.c_int, [C.c_void_p, _pp, _i64p, C.c_int]),
    "ddpm_get_weights": (C.c_int, [C.c_void_p, _pp, _i64p, C.c_int]),
    "ddpm_set_adam": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]),
    "ddpm_get_adam_state": (C.c_int, [C.c_void_p, _pp, _pp, _i64p, C.c_int, _f32p, _i64p]),
    "ddpm_set_adam_state": (C.c_int, [C.c_void_p, _pp, _pp, _i64p, C.c_int, _f32p, C.c_int64]),
    "ddpm_q_sample": (C.c_int, [C.c_void_p, _f32p, _i32p, _f32p, C.c_int, _f32p]),
    "ddpm_predict_eps": (C.c_int, [C.c_void_p, _f32p, _i32p, C.c_int, C.c_int, _f32p]),
    "ddpm_train_step": (C.c_int, [C.c_void_p, _f32p, _i32p, _f32p, C.c_int, _f32p]),
    "ddpm_loss_and_grad": (C.c_int, [C.c_void_p, _f32p, _i32p, _f32p, C.c_int, _f32p, _pp, _i64p, C.c_int]),
    "ddpm_upload_dataset": (C.c_int, [C.c_void_p, _f32p, C.c_int64]),
    "ddpm_train_step_device": (C.c_int, [C.c_void_p, _i32p, C.c_int, C.c_uint64, C.c_int64, _f32p]),
    "ddpm_sample": (C.c_int, [C.c_void_p, _f32p, _f32p, C.c_uint64, C.c_int64, C.c_int64, C.c_int, _f32p]),
    "ddpm_sample_device": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int]),
    "ddpm_sample_fetch": (C.c_int, [C.c_void_p, C.c_int64, _f32p]),
    "ddpm_sample_fetch_u8": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint8)]),
    "ddpm_apply_noise_f64": (C.c_int, [_f64p, _f64p, C.c_int64, _f64p, C.c_int, _f64p]),
    "ddpm_comm_unique_id": (C.c_int, [C.c_void_p]),
    "ddpm_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ddpm_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "ddpm_get_counter": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "ddpm_timer_start": (C.c_int, [C.c_void_p]),
    "ddpm_timer_stop": (C.c_int, [C.c_void_p, _f32p]),
    "ddpm_time_kernel": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, _f32p, _f64p, _f64p]),
    "ddpm_debug_fetch": (C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64, _i64p]),
}


class DDPMError(RuntimeError):
    pass


_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen libddpm.so and type every exported entry point.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise DDPMError(f"{path} not found: build it with `python __graft_entry__.py` (there is no CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        msg = load_library().ddpm_last_error()
        raise DDPMError(msg.decode("utf-8", "replace") if msg else f"libddpm error {rc}")


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray, typ=_f32p):
    return a.ctypes.data_as(typ)


def device_count() -> int:
    return int(load_library().ddpm_device_count())


def array_lengths() -> List[int]:
    buf = (C.c_int64 * NUM_ARRAYS)()
    _check(load_library().ddpm_array_lengths(buf))
    return [int(v) for v in buf]


def apply_noise_f64(img: np.ndarray, eps: np.ndarray, betas: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.float64)
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    betas = np.ascontiguousarray(betas, dtype=np.float64)
    if img.shape != eps.shape:
        raise ValueError("img and eps must have the same shape")
    out = np.empty_like(img)
    _check(load_library().ddpm_apply_noise_f64(_ptr(img, _f64p), _ptr(eps, _f64p), img.size, _ptr(betas, _f64p),
                                               betas.size, _ptr(out, _f64p)))
    return out


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load_library().ddpm_comm_unique_id(buf))
    return buf.raw


class Handle:
    """One engine on one GPU (``ddpm_create`` .. ``ddpm_destroy``)."""

    def __init__(self, T: int = 500, D: int = 128, H: int = 32, W: int = 32, precision: int = PREC_FP16, device: int = 0):
        self.lib = load_library()
        self.T, self.D, self.H, self.W = T, D, H, W
        self.precision = precision
        self._h = C.c_void_p()
        _check(self.lib.ddpm_create(C.byref(self._h), T, D, H, W, precision, device))
        self._lens = array_lengths()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.ddpm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- tables / weights
    def set_tables(self, beta, alpha_cum, pe):
        beta, alpha_cum, pe = _f32(beta), _f32(alpha_cum), _f32(pe)
        if beta.shape != (self.T,) or alpha_cum.shape != (self.T,) or pe.shape != (self.T, self.D):
            raise ValueError("table shapes must be (T,), (T,), (T, D)")
        _check(self.lib.ddpm_set_tables(self._h, _ptr(beta), _ptr(alpha_cum), _ptr(pe)))

    def get_tables(self):
        beta = np.empty(self.T, np.float32)
        ac = np.empty(self.T, np.float32)
        pe = np.empty((self.T, self.D), np.float32)
        samp = np.empty((self.T, 4), np.float32)
        _check(self.lib.ddpm_get_tables(self._h, _ptr(beta), _ptr(ac), _ptr(pe), _ptr(samp)))
        return beta, ac, pe, samp

    def _array_args(self, arrays: Sequence[np.ndarray]):
        if len(arrays) != NUM_ARRAYS:
            raise ValueError(f"expected {NUM_ARRAYS} arrays, got {len(arrays)}")
        keep = [_f32(a).reshape(-1) for a in arrays]
        for a, n in zip(keep, self._lens):
            if a.size != n:
                raise ValueError(f"array length mismatch: {a.size} != {n}")
        ptrs = (_f32p * NUM_ARRAYS)(*[_ptr(a) for a in keep])
        lens = (C.c_int64 * NUM_ARRAYS)(*self._lens)
        return keep, ptrs, lens

    def set_weights(self, arrays: Sequence[np.ndarray]):
        keep, ptrs, lens = self._array_args(arrays)
        _check(self.lib.ddpm_set_weights(self._h, ptrs, lens, NUM_ARRAYS))

    def get_weights(self) -> List[np.ndarray]:
        outs = [np.empty(n, np.float32) for n in self._lens]
        keep, ptrs, lens = self._array_args(outs)
        _check(self.lib.ddpm_get_weights(self._h, ptrs, lens, NUM_ARRAYS))
        return keep

    def set_adam(self, eta=1e-4, beta1=0.9, beta2=0.999, eps=1e-8):
        _check(self.lib.ddpm_set_adam(self._h, eta, beta1, beta2, eps))

    def get_adam_state(self):
        """(m arrays, v arrays, (beta1^t, beta2^t), applied steps) -- the optimiser state a resume needs."""
        m = [np.zeros(n, np.float32) for n in self._lens]
        v = [np.zeros(n, np.float32) for n in self._lens]
        km, pm, lens = self._array_args(m)
        kv, pv, _ = self._array_args(v)
        bt = np.zeros(2, np.float32)
        steps = C.c_int64()
        _check(self.lib.ddpm_get_adam_state(self._h, pm, pv, lens, NUM_ARRAYS, _ptr(bt), C.byref(steps)))
        return km, kv, (float(bt[0]), float(bt[1])), int(steps.value)

    def set_adam_state(self, m, v, beta_t, steps: int):
        km, pm, lens = self._array_args(m)
        kv, pv, _ = self._array_args(v)
        bt = np.asarray(beta_t, dtype=np.float32).reshape(2)
        _check(self.lib.ddpm_set_adam_state(self._h, pm, pv, lens, NUM_ARRAYS, _ptr(bt), int(steps)))

    # ---- hot-path calls
    def _imgs(self, x, B=None) -> np.ndarray:
        x = _f32(x)
        n = x.size // (self.H * self.W)
        if x.size != n * self.H * self.W or (B is not None and n != B):
            raise ValueError("image batch has the wrong size")
        return x.reshape(n, self.H * self.W)

    def q_sample(self, x0, ts, eps) -> np.ndarray:
        ts = np.ascontiguousarray(ts, dtype=np.int32)
        B = ts.size
        x0, eps = self._imgs(x0, B), self._imgs(eps, B)
        out = np.empty_like(x0)
        _check(self.lib.ddpm_q_sample(self._h, _ptr(x0), _ptr(ts, _i32p), _ptr(eps), B, _ptr(out)))
        return out.reshape(B, 1, self.H, self.W)

    def predict_eps(self, x_t, ts, train_mode: bool = False) -> np.ndarray:
        ts = np.ascontiguousarray(ts, dtype=np.int32)
        B = ts.size
        x_t = self._imgs(x_t, B)
        out = np.empty_like(x_t)
        _check(self.lib.ddpm_predict_eps(self._h, _ptr(x_t), _ptr(ts, _i32p), B, 1 if train_mode else 0, _ptr(out)))
        return out.reshape(B, 1, self.H, self.W)

    def train_step(self, x0, ts, eps) -> float:
        ts = np.ascontiguousarray(ts, dtype=np.int32)
        B = ts.size
        x0, eps = self._imgs(x0, B), self._imgs(eps, B)
        loss = C.c_float()
        _check(self.lib.ddpm_train_step(self._h, _ptr(x0), _ptr(ts, _i32p), _ptr(eps), B, C.byref(loss)))
        return float(loss.value)

    def loss_and_grad(self, x0, ts, eps):
        ts = np.ascontiguousarray(ts, dtype=np.int32)
        B = ts.size
        x0, eps = self._imgs(x0, B), self._imgs(eps, B)
        outs = [np.zeros(n, np.float32) for n in self._lens]
        keep, ptrs, lens = self._array_args(outs)
        loss = C.c_float()
        _check(self.lib.ddpm_loss_and_grad(self._h, _ptr(x0), _ptr(ts, _i32p), _ptr(eps), B, C.byref(loss), ptrs, lens,
                                           NUM_ARRAYS))
        return float(loss.value), keep

    def upload_dataset(self, imgs):
        imgs = self._imgs(imgs)
        _check(self.lib.ddpm_upload_dataset(self._h, _ptr(imgs), imgs.shape[0]))

    def train_step_device(self, B: int, seed: int, step: int, idx=None, want_loss: bool = True) -> Optional[float]:
        loss = C.c_float()
        ip = None
        if idx is not None:
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            if idx.size != B:
                raise ValueError("idx must have B entries")
            ip = _ptr(idx, _i32p)
        _check(self.lib.ddpm_train_step_device(self._h, ip, B, seed, step, C.byref(loss) if want_loss else None))
        return float(loss.value) if want_loss else None

    def sample(self, N: int, x_T=None, z=None, seed: int = 0, first_index: int = 0, t_start: Optional[int] = None,
               out: Optional[np.ndarray] = None) -> np.ndarray:
        """``x_T`` / ``out`` are used in place when they already are C-contiguous Float32 (e.g. views of
        pinned host memory), so the copies inside the call are the only host<->device traffic."""
        t_start = self.T if t_start is None else t_start
        xp = zp = None
        if x_T is not None:
            x_T = self._imgs(x_T, N)
            xp = _ptr(x_T)
        if z is not None:
            z = _f32(z)
            if z.size != (t_start - 1) * N * self.H * self.W:
                raise ValueError("z must hold (t_start-1)*N*H*W values")
            zp = _ptr(z)
        if out is None:
            out = np.empty((N, self.H * self.W), np.float32)
        _check(self.lib.ddpm_sample(self._h, xp, zp, seed, N, first_index, t_start, _ptr(out)))
        return out.reshape(N, 1, self.H, self.W)

    def sample_device(self, N: int, seed: int = 0, first_index: int = 0, t_start: Optional[int] = None):
        _check(self.lib.ddpm_sample_device(self._h, seed, N, first_index, self.T if t_start is None else t_start))

    def sample_fetch(self, N: int) -> np.ndarray:
        out = np.empty((N, self.H * self.W), np.float32)
        _check(self.lib.ddpm_sample_fetch(self._h, N, _ptr(out)))
        return out.reshape(N, 1, self.H, self.W)

    def sample_fetch_u8(self, N: int) -> np.ndarray:
        out = np.empty((N, self.H * self.W), np.uint8)
        _check(self.lib.ddpm_sample_fetch_u8(self._h, N, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out.reshape(N, 1, self.H, self.W)

    # ---- multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, world: int, sync_bn: bool = True):
        buf = C.create_string_buffer(unique_id, 128)
        _check(self.lib.ddpm_comm_init(self._h, buf, rank, world, 1 if sync_bn else 0))

    # ---- instrumentation
    def set_option(self, key: str, value: int):
        _check(self.lib.ddpm_set_option(self._h, key.encode(), int(value)))

    def counter(self, key: str) -> int:
        return int(self.lib.ddpm_get_counter(self._h, key.encode()))

    def timer_start(self):
        _check(self.lib.ddpm_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(self.lib.ddpm_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def time_kernel(self, name: str, n_images: int, iters: int = 20):
        ms, by, fl = C.c_float(), C.c_double(), C.c_double()
        _check(self.lib.ddpm_time_kernel(self._h, name.encode(), n_images, iters, C.byref(ms), C.byref(by), C.byref(fl)))
        return float(ms.value), float(by.value), float(fl.value)

    def debug_fetch(self, name: str) -> np.ndarray:
        n = C.c_int64()
        self.lib.ddpm_debug_fetch(self._h, name.encode(), None, 0, C.byref(n))
        if n.value <= 0:
            _check(1)
        out = np.empty(n.value, np.float32)
        _check(self.lib.ddpm_debug_fetch(self._h, name.encode(), _ptr(out), n.value, C.byref(n)))
        return out
