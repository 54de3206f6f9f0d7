#!/usr/bin/env python3
"""bench.py -- RIPP inner-pairing-product hot path on B200 (contract: see the task statement).

Workload at N GPUs (weak scaling): BASELINE.json configs[1], `PairingInnerProduct` + AFGHO16
commitment of 2^16 (G1, G2) BLS12-381 pairs PER GPU (afgho16/mod.rs:30-32 ->
inner_products/src/lib.rs:77-116).  A step = one commitment: multi-Miller loop over the rank's
slice, one Fq12 partial per rank, all-gather of the partials (NCCL) when N > 1, one shared final
exponentiation.  Metric: Miller-loop pairs per second, whole job.

  value     device-resident inputs, CUDA-event timed, L2 flushed between steps
  e2e       the same commitment through the host-pointer C-ABI call (ripp_pairing_ip, arkworks
            Jacobian layout in pinned host memory): H2D + normalise + kernels + D2H of the GT result
  roofline  integer-pipe (IMAD.WIDE) roofline of the Miller kernel: algorithmic MAC32 per pair x
            pairs / kernel time, against the IMAD.WIDE peak measured live by ripp_bench_imad
  cpu_baseline / --impl reference: the CPU restatement of the reference path (oracle/), see DESIGN.md
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic op counts of the reference algorithm (own Fq12 accumulator per pair), measured by
# tests/test_hostsim.py::test_op_counts; one Fq Montgomery product = 2*12^2 + 12 = 300 MAC32.
FQ_MUL_PER_MILLER_PAIR = 6700
FQ_MUL_PER_FINAL_EXP = 8276
MAC32_PER_FQ_MUL = 300
LOG_N = 16


def _clock_sampler(stop, samples):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    idx = os.environ.get("LOCAL_RANK", "0")
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", "-i", idx, "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=5).stdout.strip()
            if out:
                samples.append([x.strip() for x in out.split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def _clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    reasons = set()
    for s in samples:
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None,
            "reasons": sorted(reasons)}


def cpu_reference_pairs_per_s(sample_pairs):
    """CPU restatement of cfg_multi_pairing timed on a bounded sample of the same workload."""
    from oracle import cpu_baseline

    return cpu_baseline.pairing_pairs_per_s(sample_pairs)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline

    vals = []
    for _ in range(args.warmup):
        cpu_baseline.pairing_pairs_per_s(cpu_baseline.SAMPLE_PAIRS)
    t_total = 0.0
    info = None
    for _ in range(args.steps):
        info = cpu_baseline.pairing_pairs_per_s(cpu_baseline.SAMPLE_PAIRS)
        vals.append(info["value"])
        t_total += info["seconds"]
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "miller_pairs_per_s", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32-limb modular (BLS12-381 Fq)",
        "data": "synthetic",
        "config": {"workload": "PairingInnerProduct + AFGHO16 commit, 2^%d pairs per GPU (bounded CPU sample per step)" % LOG_N},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--logn", type=int, default=LOG_N)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from ripp_b200 import _lib, codec, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)
    # one non-default stream shared by torch (events, NCCL ordering) and the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    n = 1 << args.logn

    # ---- inputs: this rank's slice of the global vectors, generated on the GPU -----------------
    a_dev = synth.g1_points_dev(ctx, "cfg2-m", n, seed=rank)
    b_dev = synth.g2_points_dev(ctx, "cfg2-k", n, seed=rank)
    partial = torch.zeros(144, dtype=torch.int32, device="cuda")
    gathered = torch.zeros(world * 144, dtype=torch.int32, device="cuda")
    result = torch.zeros(144, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        if world == 1:
            ctx.pairing_ip_dev(a_dev, b_dev, n, result.data_ptr())
        else:
            ctx.miller_partial_dev(a_dev, b_dev, n, partial.data_ptr())
            dist.all_gather_into_tensor(gathered, partial)
            ctx.gt_combine_dev(gathered.data_ptr(), world, result.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    stop, samples = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, samples), daemon=True)
    th.start()

    # ---- device-resident timing: K steps, each bracketed by events, L2 flushed in between ---------
    launches0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.zero_()
        s.record()
        step()
        e.record()
    barrier()
    launches = ctx.launches - launches0
    ms_total = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n * args.steps / (ms_total * 1e-3)

    # ---- dominant kernel alone (Miller loop + warp/tree product, no final exp) ------------------
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s, e in kev:
        flush.zero_()
        s.record()
        ctx.miller_partial_dev(a_dev, b_dev, n, partial.data_ptr())
        e.record()
    torch.cuda.synchronize()
    k_ms = sum(s.elapsed_time(e) for s, e in kev) / args.steps

    # ---- end to end through the host-pointer C ABI ---------------------------------------------
    a_aff = a_dev.download((n, 24))
    b_aff = b_dev.download((n, 48))
    one_q = codec.fq_enc(1)
    g1_jac = torch.empty((n, 36), dtype=torch.int32).pin_memory()
    g2_jac = torch.empty((n, 72), dtype=torch.int32).pin_memory()
    g1_np, g2_np = g1_jac.numpy().view(np.uint32), g2_jac.numpy().view(np.uint32)
    g1_np[:, :24] = a_aff
    g1_np[:, 24:] = one_q
    g2_np[:, :48] = b_aff
    g2_np[:, 48:60] = one_q
    g2_np[:, 60:] = 0
    for _ in range(2):
        host_out = ctx.pairing_ip(g1_np, g2_np)
    assert (host_out.view(np.int32) == result.cpu().numpy()).all() or world > 1
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_out = ctx.pairing_ip(g1_np, g2_np)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())

    stop.set()
    th.join(timeout=2)

    # ---- roofline ------------------------------------------------------------------------------
    imad_peak, _ = ctx.bench_imad(0, 4096)
    imad32_peak, _ = ctx.bench_imad(1, 4096)
    chain_peak, _ = ctx.bench_imad(2, 4096)
    macs_per_launch = n * FQ_MUL_PER_MILLER_PAIR * MAC32_PER_FQ_MUL
    achieved = macs_per_launch / (k_ms * 1e-3)

    if rank == 0:
        line = {
            "metric": "miller_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32-limb modular (BLS12-381 Fq)", "data": "synthetic",
            "config": {"workload": "PairingInnerProduct + AFGHO16 commit, 2^%d (G1,G2) pairs per GPU, BLS12-381 (BASELINE configs[1])" % args.logn,
                       "pairs_per_gpu": n, "l2": "flushed between steps (256 MiB memset)",
                       "parallelism": "input slices per GPU, all-gather of one Fq12 partial per rank"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": n * (144 + 288), "d2h_bytes_per_step": 576},
            "gpu_launches": launches,
            "clocks": _clocks_summary(samples),
            "roofline": {"bound": "int32 multiply pipe (IMAD.WIDE)", "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
                         "unit": "TMAC32/s", "frac": achieved / imad_peak, "traffic": None,
                         "kernel": "k_miller (+ Fq12 product tree)", "kernel_ms": k_ms,
                         "peak_source": "ripp_bench_imad measured in this run (independent IMAD.WIDE.U32 chains)",
                         "peak_imad32_tmacs": imad32_peak / 1e12, "peak_carry_chain_tmacs": chain_peak / 1e12,
                         "mac32_per_pair": FQ_MUL_PER_MILLER_PAIR * MAC32_PER_FQ_MUL},
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                info = cpu_reference_pairs_per_s(None)
                line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
