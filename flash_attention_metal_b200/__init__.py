"""flash_attention_metal_b200 -- host-side mirror of the reference's operator surface.

The product is ``libflash_attn_b200.so`` (hand-written CUDA for sm_100a behind the
C ABI in ``include/flash_attn_b200.h``).  This module is only its ctypes binding:
one Python function per kernel of the reference, same name, same argument order
as the reference's buffer indices (``main.mm`` dispatch sites are cited in the
header).  Arguments are raw device pointers (``int``) or anything exposing
``data_ptr()`` (a torch CUDA tensor); nothing is computed in Python and there is
no CPU fallback -- if the shared library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# FA_B200_LIB: development override to A/B two builds of the library (never a fallback)
LIB_PATH = os.environ.get("FA_B200_LIB") or os.path.join(_HERE, "libflash_attn_b200.so")

FP16, BF16 = 0, 1
NAIVE, V1, V2 = 0, 1, 2
#: ring transports (FA_RING_TRANSPORT_* of the header)
TRANSPORT_AUTO, TRANSPORT_NCCL, TRANSPORT_NCCL_GATHER, TRANSPORT_PEER = 0, 1, 2, 3
#: backward implementations (fa_set_backward_algorithm)
BWD_TWO_KERNEL, BWD_FUSED = 0, 1

#: every symbol include/flash_attn_b200.h declares (tests check they are all exported)
EXPORTS = (
    "naive_attention", "flash_attention", "flash_attention_v2", "flash_attention_v2_batched",
    "flash_attention_simd", "flash_attention_v4_half", "flash_attention_backward",
    "fa_workspace_bytes_backward", "fa_host_attention_f32", "fa_host_attention_half",
    "fa_host_attention_fwd_bwd_half", "fa_host_attention_fwd_bwd_half_ex", "fa_host_release",
    "flash_attention_v4_half_rect", "fa_ring_unique_id_bytes", "fa_ring_get_unique_id", "fa_ring_create",
    "fa_ring_destroy", "fa_ring_workspace_bytes", "fa_ring_attention_forward", "fa_ring_plan", "fa_ring_local_rows",
    "fa_ring_workspace_bytes_backward", "fa_ring_attention_backward", "fa_ring_workspace_bytes_gather",
    "fa_ring_create_ex", "fa_ring_transport", "fa_ring_workspace_bytes_ex", "fa_ring_merge_plan",
    "flash_attention_backward_rect", "fa_rowsum_delta",
    "fa_mgpu_create", "fa_mgpu_destroy", "fa_mgpu_device_count", "fa_mgpu_stream", "fa_mgpu_synchronize",
    "fa_mgpu_sharded_forward", "fa_mgpu_sharded_backward", "fa_mgpu_ring_forward", "fa_mgpu_ring_backward",
    "fa_last_error", "fa_version", "fa_device_count", "fa_preload_kernels", "fa_set_backward_algorithm",
    "fa_get_backward_algorithm", "fa_launch_count", "fa_reset_launch_count",
)


class FlashAttnError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (no GPU needed)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", _HERE, "-j8"], stdout=out)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FlashAttnError(
                f"{LIB_PATH} is missing: build it with `make -C {_HERE}` or "
                "`python -c 'import __graft_entry__ as g; g.build()'` -- there is no fallback path")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
        for name in ("naive_attention", "flash_attention", "flash_attention_v2"):
            getattr(L, name).argtypes = [vp, vp, vp, vp, i32, i32, f32, i32, vp]
        L.flash_attention_v2_batched.argtypes = [vp, vp, vp, vp, i32, i32, f32, i64, i64, i32, i32, i32, vp]
        L.flash_attention_simd.argtypes = [vp, vp, vp, vp, i32, i32, f32, i32, vp]
        L.flash_attention_v4_half.argtypes = [vp, vp, vp, vp, i32, i32, f32, i64, i64, vp, i32, i32, i32, i32, vp]
        L.flash_attention_backward.argtypes = [vp] * 9 + [i32, i32, f32, i64, i64, i32, i32, i32, i32, vp, sz, vp]
        L.fa_workspace_bytes_backward.argtypes = [i32, i32, i32, i32]
        L.fa_workspace_bytes_backward.restype = sz
        L.fa_host_attention_f32.argtypes = [i32, vp, vp, vp, vp, i32, i32, f32, i32]
        L.fa_host_attention_half.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, i32, i32, i32, i32]
        L.fa_host_attention_fwd_bwd_half.argtypes = [vp] * 9 + [i32, i32, f32, i32, i32, i32, i32]
        L.fa_host_attention_fwd_bwd_half_ex.argtypes = [vp] * 9 + [i32, i32, f32, i32, i32, i32, i32, i32]
        L.fa_set_backward_algorithm.argtypes = [i32]
        L.fa_host_release.restype = None
        L.fa_last_error.restype = C.c_char_p
        L.fa_launch_count.restype = C.c_long
        L.fa_reset_launch_count.restype = None
        L.flash_attention_v4_half_rect.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, i64, i64, i64, i64, vp, i32, i32, i32, vp]
        L.fa_ring_get_unique_id.argtypes = [vp, i32]
        L.fa_ring_create.argtypes = [C.POINTER(vp), vp, i32, i32, i32]
        L.fa_ring_destroy.argtypes = [vp]
        L.fa_ring_workspace_bytes.argtypes = [i32, i32, i32, i32]
        L.fa_ring_workspace_bytes.restype = sz
        L.fa_ring_attention_forward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, i32, vp, sz, vp]
        L.fa_ring_workspace_bytes_gather.argtypes = [i32, i32, i32, i32, i32]
        L.fa_ring_workspace_bytes_gather.restype = sz
        L.fa_ring_workspace_bytes_backward.argtypes = [i32, i32, i32, i32]
        L.fa_ring_workspace_bytes_backward.restype = sz
        L.fa_ring_attention_backward.argtypes = [vp] * 10 + [i32, i32, i32, f32, i32, i32, vp, sz, vp]
        L.fa_ring_create_ex.argtypes = [C.POINTER(vp), vp, i32, i32, i32, i32]
        L.fa_ring_transport.argtypes = [vp]
        L.fa_ring_workspace_bytes_ex.argtypes = [i32, i32, i32, i32, i32, i32]
        L.fa_ring_workspace_bytes_ex.restype = sz
        L.flash_attention_backward_rect.argtypes = [vp] * 9 + [i32, i32, i32, f32, i64, i64, i64, i64, i32, i32, i32, i32, vp]
        L.fa_rowsum_delta.argtypes = [vp, vp, vp, i32, i32, i64, i64, i32, i32, i32, vp]
        pp = C.POINTER(vp)
        L.fa_mgpu_create.argtypes = [C.POINTER(vp), C.POINTER(i32), i32]
        L.fa_mgpu_destroy.argtypes = [vp]
        L.fa_mgpu_device_count.argtypes = [vp]
        L.fa_mgpu_stream.argtypes = [vp, i32]
        L.fa_mgpu_stream.restype = vp
        L.fa_mgpu_synchronize.argtypes = [vp]
        L.fa_mgpu_sharded_forward.argtypes = [vp] + [pp] * 5 + [i32, i32, f32, i32, C.POINTER(i32), i32]
        L.fa_mgpu_sharded_backward.argtypes = [vp] + [pp] * 9 + [i32, i32, f32, i32, C.POINTER(i32), i32]
        L.fa_mgpu_ring_forward.argtypes = [vp] + [pp] * 5 + [i32, i32, i32, f32, i32, i32]
        L.fa_mgpu_ring_backward.argtypes = [vp] + [pp] * 9 + [i32, i32, i32, f32, i32, i32]
        ip = C.POINTER(i32)
        L.fa_ring_plan.argtypes = [i32, i32, i32, i32, i32, ip, ip, ip, ip, ip, ip]
        L.fa_ring_local_rows.argtypes = [i32, i32, i32, i32, C.POINTER(i64), ip]
        L.fa_ring_merge_plan.argtypes = [i32, i32, i32, i32, i32, ip, ip, ip]
        _lib = L
    return _lib


def _ptr(x) -> int | None:
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):  # numpy array: host pointer (fa_host_* calls only)
        return x.ctypes.data
    raise TypeError(f"cannot take a pointer from {type(x)!r}")


def _stream(stream) -> int | None:
    if stream is None:
        return None
    return getattr(stream, "cuda_stream", stream)


def _check(rc: int) -> None:
    if rc != 0:
        raise FlashAttnError(f"flash_attn_b200 error {rc}: {lib().fa_last_error().decode()}")


# -- fp32 variants: (Q, K, V, O, N, D, scale) = reference buffer indices 0..6 --
def naive_attention(Q, K, V, O, N, D, scale, is_causal=False, stream=None):
    """kernels.metal:12-64 / main.mm:162-191."""
    _check(lib().naive_attention(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), N, D, scale, int(is_causal), _stream(stream)))


def flash_attention(Q, K, V, O, N, D, scale, is_causal=False, stream=None):
    """V1, kernels.metal:72-171 / main.mm:198-224."""
    _check(lib().flash_attention(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), N, D, scale, int(is_causal), _stream(stream)))


def flash_attention_v2(Q, K, V, O, N, D, scale, is_causal=False, stream=None):
    """V2, kernels.metal:462-596 / main.mm:259-275."""
    _check(lib().flash_attention_v2(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), N, D, scale, int(is_causal), _stream(stream)))


def flash_attention_v2_batched(Q, K, V, O, N, D, scale, batch_stride, head_stride, is_causal, B, H, stream=None):
    _check(lib().flash_attention_v2_batched(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), N, D, scale, batch_stride,
                                            head_stride, int(is_causal), B, H, _stream(stream)))


# -- 16-bit variants -----------------------------------------------------------
def flash_attention_simd(Q, K, V, O, N, D, scale, dtype=FP16, stream=None):
    """V3, kernels.metal:177-455 / main.mm:332-349."""
    _check(lib().flash_attention_simd(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), N, D, scale, dtype, _stream(stream)))


def flash_attention_v4_half(Q, K, V, O, N, D, scale, batch_stride, head_stride, L_out, is_causal, B=1, H=1,
                            dtype=FP16, stream=None):
    """V4, kernels.metal:600-883 / main.mm:414-440 (buffer indices 0..10, then B, H, dtype)."""
    _check(lib().flash_attention_v4_half(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), N, D, scale, batch_stride, head_stride,
                                         _ptr(L_out), int(is_causal), B, H, dtype, _stream(stream)))


def flash_attention_v4_half_rect(Q, K, V, O, Nq, Nk, D, scale, q_batch_stride, q_head_stride, kv_batch_stride,
                                 kv_head_stride, L_out, B=1, H=1, dtype=FP16, stream=None):
    """Cross-attention form (Nq != Nk, non-causal); the reference has only Nq == Nk."""
    _check(lib().flash_attention_v4_half_rect(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), Nq, Nk, D, scale, q_batch_stride,
                                              q_head_stride, kv_batch_stride, kv_head_stride, _ptr(L_out), B, H, dtype,
                                              _stream(stream)))


# -- ring / context-parallel attention (one process per GPU) -----------------------
def ring_unique_id() -> bytes:
    n = lib().fa_ring_unique_id_bytes()
    buf = C.create_string_buffer(n)
    _check(lib().fa_ring_get_unique_id(buf, n))
    return buf.raw


class Ring:
    """fa_ring_t wrapper.  `unique_id` comes from ring_unique_id() on rank 0 and must reach every rank.
    `transport` (TRANSPORT_*) must be the same on every rank; `self.transport` is the one in effect."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int, transport: int = TRANSPORT_AUTO):
        self.handle = C.c_void_p()
        self.rank, self.world = rank, world
        _check(lib().fa_ring_create_ex(C.byref(self.handle), unique_id, rank, world, device, transport))
        self.transport = int(lib().fa_ring_transport(self.handle))

    def workspace_bytes(self, n_local, D, H, dtype) -> int:
        """Forward workspace for this ring's transport."""
        return int(lib().fa_ring_workspace_bytes_ex(self.world, self.transport, n_local, D, H, dtype))

    def forward(self, Q, K, V, O, L_out, n_local, D, H, scale, is_causal, dtype, workspace, workspace_bytes, stream=None):
        _check(lib().fa_ring_attention_forward(self.handle, _ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(L_out), n_local, D,
                                               H, scale, int(is_causal), dtype, _ptr(workspace), workspace_bytes,
                                               _stream(stream)))

    def workspace_bytes_backward(self, n_local, D, H, dtype) -> int:
        return int(lib().fa_ring_workspace_bytes_backward(n_local, D, H, dtype))

    def backward(self, Q, K, V, O, dO, L, dQ, dK, dV, n_local, D, H, scale, is_causal, dtype, workspace, workspace_bytes,
                 stream=None):
        _check(lib().fa_ring_attention_backward(self.handle, _ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(L), _ptr(dQ),
                                                _ptr(dK), _ptr(dV), n_local, D, H, scale, int(is_causal), dtype,
                                                _ptr(workspace), workspace_bytes, _stream(stream)))

    def close(self):
        if self.handle:
            lib().fa_ring_destroy(self.handle)
            self.handle = C.c_void_p()


def _ptr_array(items):
    arr = (C.c_void_p * len(items))()
    for i, x in enumerate(items):
        arr[i] = _ptr(x)
    return arr


class Mgpu:
    """fa_mgpu_t wrapper: one process driving several GPUs.  Tensor arguments are lists with one device
    tensor (or pointer) per group device.  Calls enqueue on the group's streams; synchronize() waits."""

    def __init__(self, devices):
        self.handle = C.c_void_p()
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        _check(lib().fa_mgpu_create(C.byref(self.handle), arr, len(self.devices)))

    def stream(self, index) -> int:
        return lib().fa_mgpu_stream(self.handle, index)

    def synchronize(self):
        _check(lib().fa_mgpu_synchronize(self.handle))

    def _heads(self, heads):
        return (C.c_int * len(heads))(*heads)

    def sharded_forward(self, Q, K, V, O, L, N, D, scale, is_causal, heads, dtype):
        _check(lib().fa_mgpu_sharded_forward(self.handle, _ptr_array(Q), _ptr_array(K), _ptr_array(V), _ptr_array(O),
                                             None if L is None else _ptr_array(L), N, D, scale, int(is_causal),
                                             self._heads(heads), dtype))

    def sharded_backward(self, Q, K, V, O, dO, L, dQ, dK, dV, N, D, scale, is_causal, heads, dtype):
        _check(lib().fa_mgpu_sharded_backward(self.handle, *(_ptr_array(t) for t in (Q, K, V, O, dO, L, dQ, dK, dV)), N, D,
                                              scale, int(is_causal), self._heads(heads), dtype))

    def ring_forward(self, Q, K, V, O, L, n_local, D, H, scale, is_causal, dtype):
        _check(lib().fa_mgpu_ring_forward(self.handle, _ptr_array(Q), _ptr_array(K), _ptr_array(V), _ptr_array(O),
                                          None if L is None else _ptr_array(L), n_local, D, H, scale, int(is_causal), dtype))

    def ring_backward(self, Q, K, V, O, dO, L, dQ, dK, dV, n_local, D, H, scale, is_causal, dtype):
        _check(lib().fa_mgpu_ring_backward(self.handle, *(_ptr_array(t) for t in (Q, K, V, O, dO, L, dQ, dK, dV)), n_local, D,
                                           H, scale, int(is_causal), dtype))

    def close(self):
        if self.handle:
            lib().fa_mgpu_destroy(self.handle)
            self.handle = C.c_void_p()


def ring_plan(rank, world, step, n_local, is_causal):
    """(src_rank, q_off, q_rows, k_off, k_rows, block_causal) of the block `rank` computes at `step`."""
    out = [C.c_int() for _ in range(6)]
    _check(lib().fa_ring_plan(rank, world, step, n_local, int(is_causal), *[C.byref(o) for o in out]))
    return tuple(o.value for o in out)


def ring_merge_plan(rank, world, step, n_local, is_causal):
    """(lo, hi, half_rows): merge modes (0 only partial, 1 first, 2 middle, 3 last) of the step's forward launch."""
    out = [C.c_int() for _ in range(3)]
    _check(lib().fa_ring_merge_plan(rank, world, step, n_local, int(is_causal), *[C.byref(o) for o in out]))
    return tuple(o.value for o in out)


def ring_local_rows(rank, world, n_local, is_causal):
    """[(first_global_row, rows), ...] of the chunk(s) a rank holds, in local order."""
    first = (C.c_int64 * 2)()
    rows = (C.c_int * 2)()
    _check(lib().fa_ring_local_rows(rank, world, n_local, int(is_causal), first, rows))
    return [(int(first[i]), int(rows[i])) for i in range(2) if rows[i] > 0]


def flash_attention_backward_rect(Q, K, V, dO, L, delta, dQ, dK, dV, Nq, Nk, D, scale, q_batch_stride, q_head_stride,
                                  kv_batch_stride, kv_head_stride, acc_dq=False, B=1, H=1, dtype=FP16, stream=None):
    """Cross-attention gradients (Nq != Nk, non-causal); L and delta are those of the FULL softmax rows."""
    _check(lib().flash_attention_backward_rect(_ptr(Q), _ptr(K), _ptr(V), _ptr(dO), _ptr(L), _ptr(delta), _ptr(dQ), _ptr(dK),
                                               _ptr(dV), Nq, Nk, D, scale, q_batch_stride, q_head_stride, kv_batch_stride,
                                               kv_head_stride, int(acc_dq), B, H, dtype, _stream(stream)))


def rowsum_delta(O, dO, delta, N, D, batch_stride, head_stride, B=1, H=1, dtype=FP16, stream=None):
    """delta[b, h, i] = sum_d O * dO (kernels.metal:983-990)."""
    _check(lib().fa_rowsum_delta(_ptr(O), _ptr(dO), _ptr(delta), N, D, batch_stride, head_stride, B, H, dtype, _stream(stream)))


def set_backward_algorithm(algorithm: int) -> None:
    """BWD_TWO_KERNEL (default) or BWD_FUSED; process-wide."""
    _check(lib().fa_set_backward_algorithm(algorithm))


def get_backward_algorithm() -> int:
    return int(lib().fa_get_backward_algorithm())


def workspace_bytes_backward(N, D, B, H) -> int:
    return int(lib().fa_workspace_bytes_backward(N, D, B, H))


def flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, N, D, scale, batch_stride, head_stride, is_causal,
                             B=1, H=1, dtype=FP16, workspace=None, workspace_bytes=0, stream=None):
    """kernels.metal:905-1265 / main.mm:1027-1061 (buffer indices 0..14, then B, H, dtype, workspace)."""
    _check(lib().flash_attention_backward(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(dO), _ptr(L), _ptr(dQ), _ptr(dK),
                                          _ptr(dV), N, D, scale, batch_stride, head_stride, int(is_causal), B, H,
                                          dtype, _ptr(workspace), workspace_bytes, _stream(stream)))


# -- host-buffer calls (numpy arrays or host pointers) ---------------------------
def host_attention_f32(variant, Q, K, V, O, N, D, scale, is_causal=False):
    _check(lib().fa_host_attention_f32(variant, _ptr(Q), _ptr(K), _ptr(V), _ptr(O), N, D, scale, int(is_causal)))


def host_attention_half(Q, K, V, O, L_out, N, D, scale, is_causal, B, H, dtype):
    _check(lib().fa_host_attention_half(_ptr(Q), _ptr(K), _ptr(V), _ptr(O), _ptr(L_out), N, D, scale,
                                        int(is_causal), B, H, dtype))


def host_attention_fwd_bwd_half(Q, K, V, dO, O, L_out, dQ, dK, dV, N, D, scale, is_causal, B, H, dtype):
    _check(lib().fa_host_attention_fwd_bwd_half(_ptr(Q), _ptr(K), _ptr(V), _ptr(dO), _ptr(O), _ptr(L_out), _ptr(dQ),
                                                _ptr(dK), _ptr(dV), N, D, scale, int(is_causal), B, H, dtype))


def host_attention_fwd_bwd_half_ex(Q, K, V, dO, O, L_out, dQ, dK, dV, N, D, scale, is_causal, B, H, dtype, grad_dtype=-1):
    """grad_dtype -1: fp32 gradients (the reference's type); FP16 / BF16: gradients rounded on the device."""
    _check(lib().fa_host_attention_fwd_bwd_half_ex(_ptr(Q), _ptr(K), _ptr(V), _ptr(dO), _ptr(O), _ptr(L_out), _ptr(dQ),
                                                   _ptr(dK), _ptr(dV), N, D, scale, int(is_causal), B, H, dtype, grad_dtype))


def host_release() -> None:
    lib().fa_host_release()


def launch_count() -> int:
    return int(lib().fa_launch_count())


def reset_launch_count() -> None:
    lib().fa_reset_launch_count()


def version() -> int:
    return int(lib().fa_version())
