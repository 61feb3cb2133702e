"""Regenerate tests/golden/reference_cpu.npz from the reference's own CPU verifier.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py

The arrays are the outputs of the loops lifted unmodified from
/root/reference/main.mm (oracle/build_ref.sh -> oracle/_ref/libref_cpu.so):
  init8/init_sum      main.mm:24-30    initRandom, first 8 values and checksum at 128*64
  fwd128              main.mm:128-159  non-causal forward, N=128, D=64, Q=K=V=initRandom
  fwd1024_rows        same loops at N=1024 (BASELINE config "reference harness shape"):
                      rows 0, 1, 511, 1023 and the fp64 sum of all outputs
  causal128           main.mm:550-578  causal forward, N=128
  bwd128_dq/dk/dv     main.mm:1092-1179 backward, N=128, inputs as main.mm:946-967
                      (0.01*initRandom rounded to fp16, K=V=Q, dO=Q), fp16 decoded
                      correctly (see oracle/ref_shim.cpp)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402


def main():
    R = oracle.ref()
    D, scale = 64, np.float32(1.0 / np.sqrt(64.0))
    out = {}
    x = np.empty(128 * D, np.float32)
    R.ref_init_random(x, x.size)
    out["init8"] = x[:8].copy()
    out["init_sum"] = np.float64(x.astype(np.float64).sum())
    q = x.reshape(128, D)
    o = np.empty_like(q)
    R.ref_forward(q, q, q, o, 128, D, scale)
    out["fwd128"] = o.copy()
    oc = np.empty_like(q)
    R.ref_forward_causal(q, q, q, oc, 128, D, scale)
    out["causal128"] = oc.copy()
    x2 = np.empty(1024 * D, np.float32)
    R.ref_init_random(x2, x2.size)
    q2 = x2.reshape(1024, D)
    o2 = np.empty_like(q2)
    R.ref_forward(q2, q2, q2, o2, 1024, D, scale)
    out["fwd1024_rows"] = o2[[0, 1, 511, 1023]].copy()
    out["fwd1024_sum"] = np.float64(o2.astype(np.float64).sum())
    qb = oracle.to_half_bits(q * np.float32(0.01), oracle.FP16)
    dq, dk, dv = (np.empty((128, D), np.float32) for _ in range(3))
    R.ref_backward(qb.reshape(-1), qb.reshape(-1).copy(), dq, dk, dv, 128, D, scale)
    out["bwd128_qbits"] = qb
    out["bwd128_dq"], out["bwd128_dk"], out["bwd128_dv"] = dq, dk, dv
    path = os.path.join(os.path.dirname(__file__), "reference_cpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
