"""Profiling target (not a pytest file): a few launches of the flagship forward (and backward
when built) so ncu has something short to replay.  usage: python tests/ncu_target.py [fwd|bwd]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import flash_attention_metal_b200 as fa

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, H, n, d = 1, 16, 16384, 128
Q, K, V, dO = (torch.randn((B, H, n, d), device="cuda").to(torch.bfloat16) for _ in range(4))
O = torch.empty_like(Q); L = torch.empty((B, H, n), device="cuda")
scale = d ** -0.5
for _ in range(4):
    fa.flash_attention_v4_half(Q, K, V, O, n, d, scale, H * n * d, n * d, L, True, B, H, fa.BF16)
if what == "bwd":
    dQ, dK, dV = (torch.empty((B, H, n, d), device="cuda") for _ in range(3))
    wsb = fa.workspace_bytes_backward(n, d, B, H); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, n, d, scale, H * n * d, n * d, True, B, H, fa.BF16, ws, wsb)
torch.cuda.synchronize()
print("ok")
