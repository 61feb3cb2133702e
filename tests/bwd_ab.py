"""Development probe (not a pytest file): the fused five-GEMM backward (bwd_fused.cu, mode 0) against the
two-kernel form (bwd_tc.cu, mode 1) on the same inputs: max relative difference of every gradient, bitwise
repeatability of the fused form, and the time of both."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

setter = lambda two_kernel: fa.set_backward_algorithm(fa.BWD_TWO_KERNEL if two_kernel else fa.BWD_FUSED)


def run(B, H, n, d, causal, reps=10):
    g = torch.Generator(device="cuda").manual_seed(5)
    Q, K, V, dO = (torch.rand((B, H, n, d), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16) for _ in range(4))
    O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
    scale = d ** -0.5
    st = torch.cuda.current_stream()
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, fa.BF16, st)
    wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    out = {}
    for mode in (1, 0):
        setter(mode)
        grads = [torch.full((B, H, n, d), float("nan"), device="cuda") for _ in range(3)]
        call = lambda: fa.flash_attention_backward(Q, K, V, O, dO, L, *grads, n, d, scale, H * n * d, n * d, causal, B, H, fa.BF16, ws, wsb, st)
        call(); torch.cuda.synchronize()
        first = [x.clone() for x in grads]
        for _ in range(2): call()
        torch.cuda.synchronize()
        same = all(torch.equal(a, b) for a, b in zip(first, grads))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            call(); ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
        out[mode] = (first, ts[len(ts) // 2], same)
    setter(1)
    flop = 2.5 * 4.0 * B * H * n * n * d * (0.5 if causal else 1.0)
    rel = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(out[0][0], out[1][0])]
    print(f"B={B} H={H} N={n} d={d} causal={int(causal)}: two-kernel {out[1][1]:.3f} ms ({flop / out[1][1] / 1e9:.0f} TF)  fused {out[0][1]:.3f} ms "
          f"({flop / out[0][1] / 1e9:.0f} TF)  speed-up {out[1][1] / out[0][1]:.2f}x  rel diff dQ/dK/dV {rel[0]:.2e} {rel[1]:.2e} {rel[2]:.2e}  "
          f"finite {all(torch.isfinite(x).all().item() for x in out[0][0])}  repeatable fused {out[0][2]} two-kernel {out[1][2]}", flush=True)


if __name__ == "__main__":
    shapes = [(1, 16, 16384, 128, True), (1, 16, 16384, 128, False), (8, 12, 4096, 64, True), (1, 4, 16384, 128, False), (1, 16, 16384, 64, True)]
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        shapes = [(1, 2, 512, 128, True), (1, 2, 512, 64, False)] + shapes + [(1, 1, 32768, 128, True), (16, 8, 1024, 64, False)]
    for s in shapes:
        try:
            run(*s)
        except Exception as e:  # a kernel fault poisons the context: stop
            print("FAILED", s, str(e)[:300], flush=True)
            break
