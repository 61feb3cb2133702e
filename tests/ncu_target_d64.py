"""Profiling target (not a pytest file): one forward and one two-kernel backward at config 4 (B=8 H=12 N=4096 d=64 causal bf16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa
B, H, n, d = 8, 12, 4096, 64
Q, K, V, dO = (torch.randn((B, H, n, d), device="cuda").to(torch.bfloat16) for _ in range(4))
O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
for _ in range(2):
    fa.flash_attention_v4_half(Q, K, V, O, n, d, d ** -0.5, H * n * d, n * d, L, True, B, H, fa.BF16)
    fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, d ** -0.5, H * n * d, n * d, True, B, H, fa.BF16, ws, wsb)
torch.cuda.synchronize()
print("ok")
