#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the attention hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the headline metric and the
north_star target are quoted on; it fits one GPU): causal attention, bf16, B=1,
H=16, N=16384, d=128, forward + backward.  One "step" = one forward + one backward
over that batch through the C ABI of libflash_attn_b200.so.  With N GPUs every rank
runs the same batch on its own GPU (batch x heads sharding, no data-path collective:
heads never interact, kernels.metal:622) -> weak scaling, value = whole-job TFLOP/s.

FLOP accounting: forward 4*B*H*N^2*d, halved when causal;
backward 2.5x forward; softmax flops not counted.

Printed keys beyond the base contract:
  roofline      the dominant kernel against the measured bf16 tensor peak
  cpu_baseline  the oracle (CPU port of the reference's verifier) on a bounded sample
  e2e           same metric through the host-buffer C-ABI call (H2D + kernels + D2H)
  fwd_tflops / bwd_tflops / fwd_ms / bwd_ms   the two passes separately
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, H, N, D = 1, 16, 16384, 128
CAUSAL = True
WORKLOAD = "causal attention fwd+bwd, bf16, B=1 H=16 N=16384 d=128 (BASELINE configs[2])"


def fwd_flops(b=B, h=H, n=N, d=D, causal=CAUSAL) -> float:
    return 4.0 * b * h * n * n * d * (0.5 if causal else 1.0)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "MEASURED_PEAKS.json"}
    # B200_PROFILING.md fallback
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (a polling
    thread, ~2 ms period); falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index: int):
        self.index, self.samples, self._stop, self._thr = index, [], threading.Event(), None
        self.nv = None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.smax = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            return
        self._thr = threading.Thread(target=self._poll, daemon=True)
        self._thr.start()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            try:
                return int(ids[self.index])
            except (ValueError, IndexError):
                pass
        return self.index

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.time(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self, t0: float, t1: float):
        if self.nv is None:
            return self._smi_fallback()
        self._stop.set()
        self._thr.join(timeout=1.0)
        nv = self.nv
        rows = [x for x in self.samples if t0 <= x[0] <= t1] or self.samples[-3:]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in rows))
        sm = [r[1] for r in rows]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.smax,
                "power_w_max": max((r[3] for r in rows), default=None), "samples": len(rows), "reasons": reasons}

    def _smi_fallback(self):
        try:
            out = subprocess.check_output(
                ["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                text=True).strip().split(",")
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": ["nvml unavailable: single nvidia-smi sample after the run"]}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no clock source available"]}


# ---------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on host cores
# ---------------------------------------------------------------------------------
def cpu_step_sample(n: int, heads: int, threads: int, prefer_ref: bool = True):
    """One CPU "step" (forward + backward) over `heads` independent heads of length n, d=128, on
    `threads` host threads.  Returns (seconds, flops, kind, description).

    kind "reference": the reference's own CPU verifier loops compiled from main.mm into
    oracle/_ref -- causal forward (main.mm:550-578) and its backward (main.mm:1092-1179, which
    exists only non-causal with K = V = Q); one head per host thread (the loops are serial).
    kind "port": the oracle restatement (oracle/cpu_ref.c), causal forward and causal backward,
    OpenMP over rows."""
    import numpy as np

    import oracle

    scale = float(D ** -0.5)
    rng = np.random.default_rng(0)
    q, k, v, do = (rng.uniform(-1, 1, (heads, n, D)).astype(np.float32) for _ in range(4))
    use_ref = prefer_ref and oracle.have_ref()
    if use_ref:
        R = oracle.ref()
        from concurrent.futures import ThreadPoolExecutor

        o = np.empty_like(q)
        qb = oracle.to_half_bits(q, oracle.FP16)
        dob = oracle.to_half_bits(do, oracle.FP16)
        g = [np.empty_like(q) for _ in range(3)]

        def one(hh):  # ctypes drops the GIL
            R.ref_forward_causal(q[hh], k[hh], v[hh], o[hh], n, D, scale)
            R.ref_backward(qb[hh].reshape(-1), dob[hh].reshape(-1), g[0][hh], g[1][hh], g[2][hh], n, D, scale)

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(heads)))
        dt = time.perf_counter() - t0
        flops = fwd_flops(1, heads, n, D, True) + 2.5 * fwd_flops(1, heads, n, D, False)
        desc = (f"{heads} independent heads x N={n} x d={D} per step, fp32: reference CPU loops compiled from main.mm into "
                f"oracle/_ref -- causal forward (main.mm:550-578) + its non-causal backward (main.mm:1092-1179); "
                f"one head per host thread")
        return dt, flops, "reference", desc
    oracle.lib().oracle_set_num_threads(threads)
    t0 = time.perf_counter()
    for hh in range(heads):
        oracle.forward(q[hh], k[hh], v[hh], scale, True)
        oracle.backward(q[hh], k[hh], v[hh], do[hh], scale, True)
    dt = time.perf_counter() - t0
    desc = (f"{heads} heads x N={n} x d={D}, fp32 oracle port (oracle/cpu_ref.c): causal forward + causal backward, "
            f"OpenMP over rows on {threads} threads")
    return dt, 3.5 * fwd_flops(1, heads, n, D, True), "port", desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_s = 512  # the reference's backward decodes fp16 inside its inner loops: ~2 s per head at N=512
    heads = cores
    for _ in range(min(args.warmup, 2)):
        cpu_step_sample(256, heads, cores)
    total, flops_total, kind, sample = 0.0, 0.0, "port", ""
    for _ in range(args.steps):
        dt, flops, kind, sample = cpu_step_sample(n_s, heads, cores)
        total += dt
        flops_total += flops
    value = flops_total / total / 1e12
    line = {
        "impl": "reference", "metric": "attention_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import flash_attention_metal_b200 as fa

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fa.lib()
    scale = float(D ** -0.5)
    st = torch.cuda.current_stream()

    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    mk = lambda: torch.rand((B, H, N, D), device="cuda", generator=g).mul_(2).sub_(1).to(torch.bfloat16)
    Q, K, V, dO = mk(), mk(), mk(), mk()
    O = torch.empty_like(Q)
    L = torch.empty((B, H, N), device="cuda")
    dQ, dK, dV = (torch.empty((B, H, N, D), device="cuda") for _ in range(3))
    wsb = fa.workspace_bytes_backward(N, D, B, H)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
    bs, hs = H * N * D, N * D

    def fwd():
        fa.flash_attention_v4_half(Q, K, V, O, N, D, scale, bs, hs, L, CAUSAL, B, H, fa.BF16, st)

    have_bwd = True

    def bwd():
        fa.flash_attention_backward(Q, K, V, O, dO, L, dQ, dK, dV, N, D, scale, bs, hs, CAUSAL, B, H, fa.BF16, ws, wsb, st)

    fwd()
    try:
        bwd()
    except fa.FlashAttnError as e:
        if "not implemented" not in str(e):
            raise
        have_bwd = False
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        fwd()
        if have_bwd:
            bwd()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    # K steps; an event after each pass so forward and backward are also known separately
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    fa.reset_launch_count()
    barrier()
    t0 = time.time()
    evs[0].record(st)
    for i in range(args.steps):
        fwd()
        evs[2 * i + 1].record(st)
        if have_bwd:
            bwd()
        evs[2 * i + 2].record(st)
    barrier()
    t1 = time.time()
    launches = fa.launch_count()
    clocks = sampler.stop(t0, t1)
    total_ms = evs[0].elapsed_time(evs[-1])
    fwd_ms = statistics.mean(evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(args.steps))
    bwd_ms = statistics.mean(evs[2 * i + 1].elapsed_time(evs[2 * i + 2]) for i in range(args.steps)) if have_bwd else None
    if world > 1:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    step_flops = fwd_flops() * (3.5 if have_bwd else 1.0)
    value = step_flops * world / (ms_per_step * 1e-3) / 1e12

    # ---- e2e: the same step (forward + backward) with HOST buffers in and out, through the
    #      host-buffer C-ABI call: pinned Q,K,V,dO -> device, kernels, O,L,dQ,dK,dV -> host ----
    e2e = None
    if not args.no_e2e:
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        hq, hk, hv, hdo = (pin((B, H, N, D), torch.bfloat16) for _ in range(4))
        for hsrc, dsrc in ((hq, Q), (hk, K), (hv, V), (hdo, dO)):
            hsrc.copy_(dsrc)
        ho = pin((B, H, N, D), torch.bfloat16)
        hl = pin((B, H, N), torch.float32)
        nbytes = B * H * N * D * 2
        if have_bwd:
            hg = [pin((B, H, N, D), torch.float32) for _ in range(3)]
            call = lambda: fa.host_attention_fwd_bwd_half(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), hdo.data_ptr(),
                                                          ho.data_ptr(), hl.data_ptr(), hg[0].data_ptr(), hg[1].data_ptr(),
                                                          hg[2].data_ptr(), N, D, scale, CAUSAL, B, H, fa.BF16)
            h2d, d2h = 4 * nbytes, nbytes + B * H * N * 4 + 3 * 2 * nbytes
        else:
            call = lambda: fa.host_attention_half(hq.data_ptr(), hk.data_ptr(), hv.data_ptr(), ho.data_ptr(), hl.data_ptr(),
                                                  N, D, scale, CAUSAL, B, H, fa.BF16)
            h2d, d2h = 3 * nbytes, nbytes + B * H * N * 4
        for _ in range(2):
            call()
        barrier()
        w0 = time.perf_counter()
        esteps = max(3, min(args.steps, 10))
        for _ in range(esteps):
            call()  # synchronous: returns after the last device->host copy has landed
        barrier()
        e_ms = (time.perf_counter() - w0) * 1e3 / esteps
        if world > 1:
            t = torch.tensor([e_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": step_flops * world / (e_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": esteps,
               "pcie_gbs": (h2d + d2h) / (e_ms * 1e-3) / 1e9,
               "what": ("forward+backward" if have_bwd else "forward") + " through the host-buffer C-ABI call "
                       "(pinned host inputs -> device, kernels, every output -> host; copies pipelined over head groups)"}
        # The host call runs the heads in groups; a small group takes the key-split forward (partial
        # results merged in a different order), so the comparison with the all-heads device run is
        # to one bf16 rounding step, not bitwise.  A misplaced head or copy would be O(1) off.
        fo = (ho.cuda().float() - O.float()).abs().max().item()
        assert fo <= 2.0 ** -8 * max(1.0, O.float().abs().max().item()), f"e2e forward result differs from the device-resident result by {fo}"
        if have_bwd:
            gq = (hg[0].cuda() - dQ).abs().max().item()
            assert gq <= 1e-2 * dQ.abs().max().item(), f"e2e dQ differs from the device-resident result by {gq}"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    fwd_tf = fwd_flops() / (fwd_ms * 1e-3) / 1e12
    bwd_tf = 2.5 * fwd_flops() / (bwd_ms * 1e-3) / 1e12 if have_bwd else None
    roof = {"bound": "tensor", "kernel": "fwd_tc_kernel<128,bf16> (flash_attention_v4_half)", "achieved": fwd_tf,
            "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": fwd_tf / peaks["bf16_tflops"],
            "peak_source": peaks["source"] + " (cuBLAS bf16 burst)", "frac_of_nominal_2250": fwd_tf / 2250.0,
            "frac_of_sustained": fwd_tf / peaks["bf16_tflops_sustained"] if peaks["bf16_tflops_sustained"] else None,
            "traffic": None, "flops_per_launch": fwd_flops(), "launch_ms": fwd_ms}
    if have_bwd:
        # the backward recomputes S and dP in both of its kernels: 7 GEMM-units of work for the 5 counted
        roof["backward"] = {"kernels": "bwd_dkdv_kernel + bwd_dq_kernel (+ bwd_delta_kernel)", "achieved": bwd_tf,
                            "frac": bwd_tf / peaks["bf16_tflops"], "issued_mma_tflops": bwd_tf * 7.0 / 5.0,
                            "issued_frac": bwd_tf * 7.0 / 5.0 / peaks["bf16_tflops"], "launch_ms": bwd_ms}
    prof = os.path.join(ROOT, "profiles", "fwd_traffic.json")
    if os.path.exists(prof):
        roof["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")

    cpu = None
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        # bounded sample of the same step (forward + backward), ~10-30 s of CPU work
        dt, flops, kind, desc = cpu_step_sample(2048, max(2, cores // 4), cores, prefer_ref=False)
        cpu = {"value": flops / dt / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": kind, "seconds": dt, "sample": desc}

    line = {
        "metric": "attention_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pass": "fwd+bwd" if have_bwd else "fwd only (backward not built)",
                   "B": B, "H_per_gpu": H, "N": N, "d": D, "causal": CAUSAL, "sharding": f"batch x heads, {world} rank(s), no collective",
                   "l2": "inputs (Q,K,V,dO = 256 MiB) exceed the 126 MB L2; no explicit flush"},
        "fwd_ms": fwd_ms, "bwd_ms": bwd_ms, "fwd_tflops": fwd_tf, "bwd_tflops": bwd_tf,
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
