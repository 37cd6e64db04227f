#!/usr/bin/env python
"""bench.py - images/sec of the IR-SDE reverse-diffusion sampler (BASELINE.json metric).

A "step" is ONE complete chain x_T -> x_0 (T=100 network forwards + T fused updates) over one batch of
synthetic LQ images: BASELINE config 2 (IR-SDE deraining, 8x3x256x256 per GPU, T=100, ConditionalUNet
nf=64 depth=4, bf16 tcgen05 path, CUDA-graph step replay).  Weak scaling: every rank owns its own batch
of 8 images, no data-path collective (one weight broadcast before, one gather of x0 after).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # the reference algorithm's PyTorch-CPU path (oracle port)
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec at 256x256 T=100 IR-SDE"
WORKLOADS = {
    # name: (B per GPU, H, W, T, nf, depth, max_sigma, eps)
    "c2": (8, 256, 256, 100, 64, 4, 10, 0.005),
    "small": (2, 64, 64, 10, 16, 2, 10, 0.005),  # plumbing check only, never a bench line
}


def synth(B, H, W, seed=1234):
    import torch
    g = torch.Generator().manual_seed(seed)
    lq = torch.rand(B, 3, H, W, generator=g)
    return lq, g


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.p = index, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        loaded = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": statistics.median(loaded) if loaded else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def build_info():
    d = {}
    try:
        for line in open(os.path.join(ROOT, "image-restoration-sde_b200", "BUILD_INFO")):
            k, _, v = line.strip().partition("=")
            d[k] = v
    except Exception:
        pass
    return d


def ncu_traffic_per_launch():
    """(mean DRAM bytes (read + write) per tcgen05 conv launch, source description) from the newest committed
    `ncu --set full` capture of this workload (profiles/rNN_conv_tc_ncu_full_one_step.csv + .meta.json), or (None, why).
    Not measured live: ncu cannot run inside a timed bench.  The description carries the commit / source hash the capture
    was taken from and whether that equals the build being timed."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_tc_ncu_full_one_step.csv")))
    files = [f for f in files if "_v1_" not in f]
    if not files:
        return None, "no profiles/r*_conv_tc_ncu_full_one_step.csv"
    path = files[-1]
    try:
        rows = list(csv.DictReader(open(path)))[1:]   # first row holds the units (Mbyte)
        tot = sum(float(r["dram__bytes_read.sum"]) + float(r["dram__bytes_write.sum"]) for r in rows)
        meta = {}
        try:
            meta = json.load(open(path[:-4] + ".meta.json"))
        except Exception:
            pass
        bi = build_info()
        same = bool(meta.get("csrc_sha256")) and meta.get("csrc_sha256") == bi.get("csrc_sha256")
        src = ("%s: mean dram read+write bytes per launch over %d launches of one step, ncu --set full; captured from commit %s "
               "(csrc %s); this build: commit %s (csrc %s) -> %s" % (
                   os.path.relpath(path, ROOT), len(rows), meta.get("commit", "?"), meta.get("csrc_sha256", "?"),
                   bi.get("commit", "?"), bi.get("csrc_sha256", "?"), "same sources" if same else "DIFFERENT sources"))
        return tot * 1e6 / len(rows), src
    except Exception as e:
        return None, "unreadable %s: %s" % (path, e)


def usable_cores():
    """Host cores this process may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        pass
    # PyTorch's CPU conv stops scaling (and on this pool's 128-thread hosts gets slower) beyond ~32 threads
    return max(1, min(n, 32))


def cpu_reference_sample(wl, steps_per_sample=3, threads=None):
    """The reference algorithm on host cores: oracle port (PyTorch CPU fp32, all threads), bounded sample =
    `steps_per_sample` network steps of ONE image of the workload; every step of the chain costs the same,
    so img/s = 1 / (T * s_per_step).  Returns (images_per_s, seconds, description)."""
    import torch
    from oracle import irsde_oracle as O
    B, H, W, T, nf, depth, ms, eps = WORKLOADS[wl]
    threads = threads or usable_cores()
    torch.set_num_threads(threads)
    P = O.make_weights(3, 3, nf, depth, seed=0)
    lq, g = synth(1, H, W)
    sc = O.Schedule(ms, T, "cosine", eps)
    xT = lq + torch.randn(lq.shape, generator=g) * sc.max_sigma
    zs = torch.randn((steps_per_sample,) + tuple(lq.shape), generator=g)
    net = lambda x, t: O.unet_forward(P, x, lq, t, nf, depth)
    with torch.no_grad():
        t0 = time.perf_counter()
        O.reverse_chain(sc, net, xT, lq, zs, "sde", T=steps_per_sample)
        dt = time.perf_counter() - t0
    s_per_step = dt / steps_per_sample
    return 1.0 / (T * s_per_step), dt, ("%d network steps of 1x3x%dx%d (nf=%d depth=%d) on %d threads; "
                                        "img/s extrapolated as 1/(T*s_per_step), T=%d" % (steps_per_sample, H, W, nf, depth, threads, T))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    wl = args.workload
    B, H, W, T, nf, depth, ms, eps = WORKLOADS[wl]
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        v, dt, sample = cpu_reference_sample(wl, steps_per_sample=2 if i < args.warmup else 6)
        if i >= args.warmup:
            vals.append(v); secs.append(dt)
    value = len(vals) / sum(1.0 / v for v in vals)  # harmonic mean = total images / total time
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sum(secs) / len(secs),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(wl), "note": "reference algorithm, PyTorch CPU fp32 (oracle port)"},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": usable_cores(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_name(wl):
    B, H, W, T, nf, depth, ms, eps = WORKLOADS[wl]
    return ("IR-SDE deraining reverse_sde, %dx3x%dx%d per GPU, T=%d, ConditionalUNet nf=%d depth=%d, max_sigma=%d "
            "cosine eps=%g" % (B, H, W, T, nf, depth, ms, eps))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import irsde_b200
    from irsde_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = args.workload
    B, H, W, T, nf, depth, ms, eps = WORKLOADS[wl]
    prec = args.precision

    # ---- model: random-init weights of the named architecture, broadcast once from rank 0
    torch.manual_seed(0)
    net = irsde_b200.ConditionalUNet(3, 3, nf, depth=depth, precision=prec).to(dev)
    if world > 1:
        irsde_b200.broadcast_weights(net, src=0)
    sde = irsde_b200.IRSDE(ms, T, schedule="cosine", eps=eps, device=dev)
    sde.set_model(net)
    sde.use_graph = not args.no_graph
    sde.image_base = rank * B   # per-image Philox: rank r owns the global images [r*B, (r+1)*B)

    lq_cpu, g = synth(B, H, W, seed=1234 + rank)
    xT_cpu = lq_cpu + torch.randn(lq_cpu.shape, generator=g) * sde.max_sigma
    lq, xT = lq_cpu.to(dev), xT_cpu.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) device-resident throughput: inputs in HBM, in-kernel Philox noise, graph replay
    sde.rng = "philox"
    sde.set_mu(lq)
    for _ in range(args.warmup):
        sde.reverse_sde(xT)
    barrier()
    ctx = net._ctx
    l0 = int(ctx.L.irsde_launch_count(ctx.h))
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        x0 = sde.reverse_sde(xT)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = int(ctx.L.irsde_launch_count(ctx.h)) - l0
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * args.steps / (ms_total / 1000.0)
    assert torch.isfinite(x0).all()

    # ---- (2) end to end through the public API with HOST buffers (pinned): H2D of LQ and x_T, the
    #          sampler exactly as test.py drives it (torch RNG, one randn_like per step), D2H of x0
    sde.rng = "torch"
    lq_pin, xT_pin = lq_cpu.pin_memory(), xT_cpu.pin_memory()
    out_pin = torch.empty_like(lq_cpu).pin_memory()

    def e2e_step():
        a = lq_pin.to(dev, non_blocking=True)
        b = xT_pin.to(dev, non_blocking=True)
        sde.set_mu(a)
        y = sde.reverse_sde(b)
        out_pin.copy_(y, non_blocking=True)
        if world > 1:  # the job's single gather of results (25 MB at config 3); part of the step
            bufs = [torch.empty_like(y) for _ in range(world)] if rank == 0 else None
            dist.gather(y, bufs, dst=0)

    for _ in range(max(1, args.warmup // 3)):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1000.0)
    nbytes = lq_cpu.numel() * 4

    # ---- (3) roofline of the dominant kernel (tcgen05 tap-GEMM conv): instrumented pass, CUDA events
    #          on the launching stream around every op of `prof_steps` sampler steps (same process,
    #          right after the timed region; graph replay bypassed so each launch can be bracketed)
    roof, breakdown = None, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained (B200_PROFILING.md)"
        ncat = 6
        ms_a, fl_a, n_a = (ctypes.c_double * ncat)(), (ctypes.c_double * ncat)(), (ctypes.c_int64 * ncat)()
        sde.rng = "philox"
        sde.set_mu(lq)
        prof_steps = min(T, 5)
        _lib.check(ctx.L.irsde_profile_begin(ctx.h), ctx.h)
        sde.reverse_sde(xT, T=prof_steps)
        _lib.check(ctx.L.irsde_profile_end(ctx.h, ms_a, fl_a, n_a, ncat), ctx.h)
        names = ["tcgen05_conv", "simt_conv", "layernorm", "attention", "misc", "update"]
        breakdown = {names[i]: {"ms_per_step": ms_a[i] / prof_steps, "ops_per_step": n_a[i] / prof_steps,
                                "tflops": (fl_a[i] / (ms_a[i] * 1e-3) / 1e12) if ms_a[i] > 0 and fl_a[i] > 0 else None}
                     for i in range(ncat) if n_a[i] > 0}
        k = 0 if n_a[0] > 0 else 1
        traffic, traffic_src = ncu_traffic_per_launch()
        if ms_a[k] > 0:
            ach = fl_a[k] / (ms_a[k] * 1e-3) / 1e12
            pk = peak_tf if k == 0 else 75.0
            roof = {"kernel": "conv_tc_kernel (tcgen05 tap-GEMM conv)" if k == 0 else "conv_simt_kernel (fp32)",
                    "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                    "peak_source": peak_src if k == 0 else "nominal fp32 FMA peak",
                    "flops_per_launch": fl_a[k] / n_a[k], "avg_launch_ms": ms_a[k] / n_a[k], "launches": int(n_a[k]),
                    "share_of_step": ms_a[k] / sum(ms_a), "traffic": traffic if k == 0 else None,
                    "traffic_source": traffic_src,
                    "how": "CUDA events around each launch, %d sampler steps, non-graph pass after the timed region" % prof_steps}

    # ---- (4) CPU baseline (rank 0, N=1 only): bounded sample of the same workload on host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, sample = cpu_reference_sample(wl, steps_per_sample=8)
        cpu = {"value": v, "unit": "images/s", "cores": usable_cores(), "kind": "port", "sample": sample, "seconds": dt}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": prec, "data": "synthetic",
                "config": {"workload": workload_name(wl), "global_batch": world * B, "parallelism": "batch-sharded dp%d" % world,
                           "graph": sde.use_graph, "l2": "per-step activation working set (>4 GB) exceeds the 126 MB L2; no flush needed",
                           "noise": "in-kernel Philox for `value`; torch.randn_like per step (reference RNG order) for e2e"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes},
                "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "breakdown": breakdown,
                "device_bytes": int(ctx.L.irsde_device_bytes(ctx.h))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
