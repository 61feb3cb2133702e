"""Development probe (not a pytest file): fp32 V2 (flash_attention_v2 / _batched) TFLOP/s against the 74 TFLOP/s FFMA peak."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

def bench(B, H, n, d, causal, reps=10):
    Q, K, V = (torch.randn((B, H, n, d), device="cuda") for _ in range(3))
    O = torch.empty_like(Q)
    st = torch.cuda.current_stream()
    run = lambda: fa.flash_attention_v2_batched(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, causal, B, H, st)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        run(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    flop = 4.0 * B * H * n * n * d * (0.5 if causal else 1.0)
    tf = flop / ts[len(ts) // 2] / 1e9
    print(f"V2 fp32 B={B} H={H} N={n} d={d} causal={int(causal)}: median {ts[len(ts)//2]:.3f} ms -> {tf:.1f} TFLOP/s = {100 * tf / 74.4:.0f} % of the 74.4 TFLOP/s FFMA peak", flush=True)

for shape in [(16, 8, 4096, 64, False), (16, 8, 4096, 64, True), (1, 1, 16384, 64, False), (16, 8, 1024, 64, False), (4, 8, 4096, 128, False)]:
    bench(*shape)
