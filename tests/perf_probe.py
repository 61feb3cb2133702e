"""Development timing probe (not a pytest file): forward TFLOP/s at a few shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

def bench(B, H, n, d, causal, dtype=fa.BF16, reps=10):
    tdt = torch.bfloat16 if dtype == fa.BF16 else torch.float16
    Q, K, V = (torch.randn((B, H, n, d), device="cuda").to(tdt) for _ in range(3))
    O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
    scale = d ** -0.5
    st = torch.cuda.current_stream()
    run = lambda: fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, dtype, st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        run(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    flop = 4.0 * B * H * n * n * d * (0.5 if causal else 1.0)
    print(f"B={B} H={H} N={n} d={d} causal={int(causal)}: median {ts[len(ts)//2]:.3f} ms best {ts[0]:.3f} ms -> {flop/ts[len(ts)//2]/1e9:.1f} TFLOP/s (best {flop/ts[0]/1e9:.1f})", flush=True)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "d64":
    bench(8, 12, 4096, 64, True)
    bench(8, 12, 4096, 64, False)
    bench(1, 16, 16384, 64, True)
    bench(1, 16, 16384, 64, False)
    bench(1, 16, 16384, 128, True)
    bench(1, 16, 16384, 128, False)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "small":
    for n in (512, 1024, 2048, 4096, 8192, 16384):
        bench(1, 1, n, 64, False, fa.FP16)
    for n in (1024, 4096, 16384):
        bench(1, 1, n, 64, True)
        bench(1, 1, n, 128, True)
    bench(1, 4, 4096, 128, True)
    bench(2, 8, 2048, 64, True)

if __name__ == "__main__" and len(sys.argv) == 1:
    bench(1, 16, 16384, 128, True)
    bench(1, 16, 16384, 128, False)
    bench(8, 12, 4096, 64, True)
    bench(1, 16, 4096, 128, True)
    bench(1, 16, 8192, 128, False)
    bench(16, 8, 1024, 64, False)
    bench(1, 1, 16384, 64, False, fa.FP16)
    bench(1, 1, 1024, 64, False, fa.FP16)
    bench(1, 1, 128, 64, False, fa.FP16)


def bench_bwd(B, H, n, d, causal, dtype=fa.BF16, reps=10):
    tdt = torch.bfloat16 if dtype == fa.BF16 else torch.float16
    Q, K, V, dO = (torch.randn((B, H, n, d), device="cuda").to(tdt) for _ in range(4))
    O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
    dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
    scale = d ** -0.5
    st = torch.cuda.current_stream()
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, dtype, st)
    wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    run = lambda: fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, H * n * d, n * d, causal, B, H, dtype, ws, wsb, st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        run(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    flop = 2.5 * 4.0 * B * H * n * n * d * (0.5 if causal else 1.0)
    print(f"BWD B={B} H={H} N={n} d={d} causal={int(causal)}: median {ts[len(ts)//2]:.3f} ms -> {flop/ts[len(ts)//2]/1e9:.1f} TFLOP/s (5-GEMM accounting; hardware does 7)", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bwd64":
    bench_bwd(8, 12, 4096, 64, True)
    bench_bwd(1, 16, 16384, 64, True)
    bench_bwd(1, 16, 16384, 64, False)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bwd":
    bench_bwd(1, 16, 16384, 128, True)
    bench_bwd(1, 16, 16384, 128, False)
    bench_bwd(8, 12, 4096, 64, True)
    bench_bwd(16, 8, 1024, 64, False)
