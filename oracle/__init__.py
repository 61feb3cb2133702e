"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes access to ``oracle/liboracle.so`` (the C restatement of the reference's CPU
verifier, ``oracle/cpu_ref.c``) and, when it has been built, to
``oracle/_ref/libref_cpu.so`` (the reference's own verifier loops compiled from
``/root/reference/main.mm`` by ``oracle/build_ref.sh``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this package.  The product package ``flash_attention_metal_b200``
never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")

FP16, BF16 = 0, 1


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and _ref/libref_cpu.so (if /root/reference exists)."""
    src = os.path.join(_HERE, "cpu_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists(os.environ.get("REFERENCE_DIR", "/root/reference") + "/main.mm"):
        shim = os.path.join(_HERE, "ref_shim.cpp")
        if force or not os.path.exists(_REF_PATH) or os.path.getmtime(_REF_PATH) < os.path.getmtime(shim):
            subprocess.check_call([os.path.join(_HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")

_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_init_random.argtypes = [_f32p, C.c_long]
        L.oracle_init_random_seeded.argtypes = [_f32p, C.c_long, C.c_uint32]
        L.oracle_f32_to_half_array.argtypes = [_f32p, _u16p, C.c_long, C.c_int]
        L.oracle_half_to_f32_array.argtypes = [_u16p, _f32p, C.c_long, C.c_int]
        L.oracle_forward_faithful.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float]
        L.oracle_forward.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
        L.oracle_forward_f64.argtypes = [_f32p, _f32p, _f32p, _f64p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int]
        bw = [_f32p] * 7 + [C.c_int, C.c_int, C.c_float, C.c_int]
        L.oracle_backward.argtypes = bw
        L.oracle_backward_streaming.argtypes = bw
        L.oracle_backward_f64.argtypes = [_f32p] * 4 + [_f64p] * 3 + [C.c_int, C.c_int, C.c_float, C.c_int]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        for name in ("oracle_init_random", "oracle_init_random_seeded", "oracle_f32_to_half_array",
                     "oracle_half_to_f32_array", "oracle_forward_faithful", "oracle_forward",
                     "oracle_forward_f64", "oracle_backward", "oracle_backward_streaming",
                     "oracle_backward_f64", "oracle_set_num_threads"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def have_ref() -> bool:
    build()
    return os.path.exists(_REF_PATH)


def ref():
    """The reference's own CPU verifier loops (compiled from /root/reference)."""
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libref_cpu.so not built (no /root/reference here)")
        R = C.CDLL(_REF_PATH)
        R.ref_init_random.argtypes = [_f32p, C.c_int]
        R.ref_forward.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float]
        R.ref_forward_causal.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float]
        R.ref_backward.argtypes = [_u16p, _u16p, _f32p, _f32p, _f32p, C.c_int, C.c_int, C.c_float]
        R.ref_backward_buggy.argtypes = [_u16p, _u16p, _f32p, C.c_int, C.c_int, C.c_float]
        for name in ("ref_init_random", "ref_forward", "ref_forward_causal", "ref_backward", "ref_backward_buggy"):
            getattr(R, name).restype = None
        _ref = R
    return _ref


# ---------------------------------------------------------------------------
# numpy-level helpers
# ---------------------------------------------------------------------------
def init_random(size: int, seed: int = 42) -> np.ndarray:
    """main.mm:24-30: mt19937(seed) -> U(-1, 1) as float32."""
    out = np.empty(size, dtype=np.float32)
    lib().oracle_init_random_seeded(out, size, seed)
    return out


def to_half_bits(x: np.ndarray, dtype: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().oracle_f32_to_half_array(x.reshape(-1), out.reshape(-1), x.size, dtype)
    return out


def from_half_bits(b: np.ndarray, dtype: int) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint16)
    out = np.empty(b.shape, dtype=np.float32)
    lib().oracle_half_to_f32_array(b.reshape(-1), out.reshape(-1), b.size, dtype)
    return out


def round_to(x: np.ndarray, dtype: int) -> np.ndarray:
    """fp32 values rounded through fp16/bf16 storage (what the 16-bit kernels see)."""
    return from_half_bits(to_half_bits(x, dtype), dtype)


def forward(q, k, v, scale: float, causal: bool = False, faithful: bool = False):
    """Single head [N, D] fp32 forward.  Returns (O, L)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    n, d = q.shape
    o = np.empty((n, d), dtype=np.float32)
    if faithful:
        assert not causal
        lib().oracle_forward_faithful(q, k, v, o, n, d, scale)
        return o, None
    lse = np.empty(n, dtype=np.float32)
    lib().oracle_forward(q, k, v, o, lse.ctypes.data, n, d, scale, int(causal))
    return o, lse


def forward_batched(q, k, v, scale: float, causal: bool = False):
    """[..., N, D] fp32 forward over independent heads.  Returns (O, L[..., N])."""
    shp = q.shape
    n, d = shp[-2], shp[-1]
    q2 = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, n, d)
    k2 = np.ascontiguousarray(k, dtype=np.float32).reshape(-1, n, d)
    v2 = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, n, d)
    o = np.empty_like(q2)
    lse = np.empty(q2.shape[:2], dtype=np.float32)
    for h in range(q2.shape[0]):
        oh, lh = forward(q2[h], k2[h], v2[h], scale, causal)
        o[h], lse[h] = oh, lh
    return o.reshape(shp), lse.reshape(shp[:-1])


def forward_f64(q, k, v, scale: float, causal: bool = False):
    q = np.ascontiguousarray(q, dtype=np.float32)
    k = np.ascontiguousarray(k, dtype=np.float32)
    v = np.ascontiguousarray(v, dtype=np.float32)
    n, d = q.shape
    o = np.empty((n, d), dtype=np.float64)
    lse = np.empty(n, dtype=np.float64)
    lib().oracle_forward_f64(q, k, v, o, lse.ctypes.data, n, d, scale, int(causal))
    return o, lse


def backward(q, k, v, do, scale: float, causal: bool = False, streaming: bool = False):
    """Single head backward, fp32 (main.mm:1091-1179 formulas).  Returns (dQ, dK, dV)."""
    q, k, v, do = (np.ascontiguousarray(t, dtype=np.float32) for t in (q, k, v, do))
    n, d = q.shape
    dq, dk, dv = (np.empty((n, d), dtype=np.float32) for _ in range(3))
    fn = lib().oracle_backward_streaming if streaming else lib().oracle_backward
    fn(q, k, v, do, dq, dk, dv, n, d, scale, int(causal))
    return dq, dk, dv


def backward_f64(q, k, v, do, scale: float, causal: bool = False):
    q, k, v, do = (np.ascontiguousarray(t, dtype=np.float32) for t in (q, k, v, do))
    n, d = q.shape
    dq, dk, dv = (np.empty((n, d), dtype=np.float64) for _ in range(3))
    lib().oracle_backward_f64(q, k, v, do, dq, dk, dv, n, d, scale, int(causal))
    return dq, dk, dv


def num_threads() -> int:
    return int(lib().oracle_num_threads())
