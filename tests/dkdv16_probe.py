"""Development probe (not a pytest file): the dK/dV kernel with 16 element-wise warps (bwd_dkdv16_kernel) against the
default 8-warp kernel: bitwise equality of all three gradients (same arithmetic, same summation order) and the time of
the whole backward with each."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

setw = fa.lib().fa_debug_set_dkdv_warps
setw.argtypes = [ctypes.c_int]


def run(B, H, n, d, causal, reps=10):
    g = torch.Generator(device="cuda").manual_seed(5)
    Q, K, V, dO = (torch.rand((B, H, n, d), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(4))
    O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
    scale = d ** -0.5
    st = torch.cuda.current_stream()
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, fa.BF16, st)
    wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    out = {}
    for warps in (8, 16):
        setw(warps)
        grads = [torch.full((B, H, n, d), float("nan"), device="cuda") for _ in range(3)]
        call = lambda: fa.flash_attention_backward(Q, K, V, O, dO, L, *grads, n, d, scale, H * n * d, n * d, causal, B, H, fa.BF16, ws, wsb, st)
        for _ in range(3): call()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            call(); ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        out[warps] = ([x.clone() for x in grads], ts[len(ts) // 2])
    setw(8)
    same = [torch.equal(a, b) for a, b in zip(out[8][0], out[16][0])]
    rel = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out[16][0], out[8][0])]
    flop = 2.5 * 4.0 * B * H * n * n * d * (0.5 if causal else 1.0)
    print(f"B={B} H={H} N={n} d={d} causal={int(causal)}: backward 8 warps {out[8][1]:.3f} ms ({flop / out[8][1] / 1e9:.0f} TF)  16 warps {out[16][1]:.3f} ms "
          f"({flop / out[16][1] / 1e9:.0f} TF)  speed-up {out[8][1] / out[16][1]:.3f}x  bitwise equal dQ/dK/dV {same}  rel diff {rel[1]:.1e} {rel[2]:.1e}", flush=True)


for s in [(1, 2, 333, 64, True), (1, 2, 200, 128, False), (8, 12, 4096, 64, True), (1, 16, 16384, 64, True), (1, 16, 16384, 64, False),
          (1, 16, 16384, 128, True), (16, 8, 1024, 64, False)]:
    try:
        run(*s)
    except Exception as e:
        print("FAILED", s, str(e)[:300], flush=True)
        break
