"""Profiling target, round 2 (not a pytest file): ONE launch of every kernel the round's evidence names, in this order:
  flagship (B=1 H=16 N=16384 d=128 causal bf16): forward, backward two-kernel (delta, dK/dV, dQ), backward fused;
  config 4 (B=8 H=12 N=4096 d=64 causal bf16): forward, backward two-kernel;
  small N, single head d=64 fp16 non-causal: N = 128, 512, 1024 (the harness' memory/latency-bound regime);
  fp32 V2: B=16 H=8 N=4096 d=64.
usage: python tests/ncu_target_r2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa


def fwd_bwd(B, H, n, d, causal, fused=False, bwd=True, dtype=fa.BF16):
    tdt = torch.bfloat16 if dtype == fa.BF16 else torch.float16
    Q, K, V, dO = (torch.randn((B, H, n, d), device="cuda").to(tdt) for _ in range(4))
    O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
    scale = d ** -0.5
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, causal, B, H, dtype)
    if bwd:
        dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
        wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, H * n * d, n * d, causal, B, H, dtype, ws, wsb)
        if fused:
            fa.set_backward_algorithm(fa.BWD_FUSED)
            fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, H * n * d, n * d, causal, B, H, dtype, ws, wsb)
            fa.set_backward_algorithm(fa.BWD_TWO_KERNEL)
    torch.cuda.synchronize()


fwd_bwd(1, 16, 16384, 128, True, fused=True)
fwd_bwd(8, 12, 4096, 64, True)
for n in (128, 512, 1024):
    fwd_bwd(1, 1, n, 64, False, bwd=False, dtype=fa.FP16)
Qf = torch.randn((16, 8, 4096, 64), device="cuda")
Of = torch.empty_like(Qf)
fa.flash_attention_v2_batched(Qf, Qf, Qf, Of, 4096, 64, 0.125, 8 * 4096 * 64, 4096 * 64, False, 16, 8)
torch.cuda.synchronize()
print("ok")
