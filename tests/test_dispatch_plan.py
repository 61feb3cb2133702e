"""CPU test of the launch geometry (csrc/sched.cuh, host side): the decode the kernels run on the
block id is re-stated here and must visit every (batch, head, block) exactly once, in an order that
is heads-fastest inside a group and heaviest-block-first across the launch."""
import ctypes

import pytest

import flash_attention_metal_b200 as fa


def plan(uneven, bytes_per_head, n_blocks, H, B):
    L = fa.lib()
    L.fa_debug_dispatch.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_int)]
    L.fa_debug_dispatch.restype = None
    out = (ctypes.c_int * 4)()
    L.fa_debug_dispatch(int(uneven), bytes_per_head, n_blocks, H, B, out)
    return tuple(out)


def visit_order(group, grid, n_heads, H):
    """decode_block (sched.cuh) over block ids in hardware dispatch order (x fastest, then y, then z)."""
    gx, gy, gz = grid
    order = []
    for z in range(gz):
        for y in range(gy):
            for x in range(gx):
                if group <= 1:
                    order.append((z, y, x))  # (b, h, blk)
                else:
                    hh = z * group + x
                    if hh < n_heads:
                        order.append((hh // H, hh % H, y))
    return order


@pytest.mark.parametrize("uneven", [False, True])
@pytest.mark.parametrize("H,B,n_blocks,bytes_per_head", [
    (16, 1, 64, 8 << 20),      # flagship: 8 MB of K/V per head -> groups of 6, 6, 4 heads
    (12, 8, 16, 1 << 20),      # GPT-2 shape: 96 heads in two groups of 48
    (1, 1, 64, 4 << 20), (5, 2, 4, 256 << 10), (3, 1, 1, 1 << 30), (7, 3, 9, 30 << 20),
])
def test_every_block_is_visited_once_in_a_balanced_order(uneven, H, B, n_blocks, bytes_per_head):
    group, gx, gy, gz = plan(uneven, bytes_per_head, n_blocks, H, B)
    n_heads = H * B
    assert 1 <= group <= n_heads
    assert max(gx, gy, gz) <= 65535 or gx == n_blocks  # grid.y / grid.z limits
    order = visit_order(group, (gx, gy, gz), n_heads, H)
    assert sorted(order) == sorted((b, h, k) for b in range(B) for h in range(H) for k in range(n_blocks))
    if not uneven:
        assert group == 1  # equal work: per-head order, best L2 locality
        return
    # uneven work: a group's streamed tensors fit the L2 budget (48 MB) unless one head alone exceeds it
    assert group == 1 or group * bytes_per_head <= 48 << 20
    # groups are balanced: sizes differ by at most one group's rounding
    n_groups = -(-n_heads // group)
    assert n_heads - (n_groups - 1) * group >= 1
    # inside a group: all heads of block k are dispatched before any head of block k + 1
    for g in range(n_groups):
        seq = [(blk, b * H + h) for (b, h, blk) in order if (b * H + h) // group == g]
        assert seq == sorted(seq)
