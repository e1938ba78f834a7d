#!/usr/bin/env python
"""bench.py - the DxMI sampler-rollout benchmark (BASELINE.json metric: T-step sampling images/sec).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B] [--T T]
  (N > 1: launched by the driver as  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...)

A "step" is one rollout of the hot path over one batch of synthetic inputs: T U-Net denoising steps, each followed
by the Gaussian transition with its learned sigma, then the energy / value net on the final samples.
Workload (config.workload): BASELINE.json configs[1] - CIFAR-10, T=4, batch 256 per GPU, bf16 tensor-core math
(--workload in64 runs configs[2]'s per-GPU shard instead: ImageNet-64 EDM, T=10, batch 64 per GPU).
The DDGAN backbone that config names is not in the reference tree, so the in-tree stand-in
`VARSampler(n_timesteps=4)` + `unet_small.Model` is used and labelled as such.

Numbers printed (one JSON line from rank 0):
  value        images/s, inputs (noise) resident in HBM, CUDA-event time of K steps, max over ranks
  e2e          images/s through the public API (`sampler.sample` + `value`) from PINNED HOST noise, with the H2D copy
               of every step's noise and the D2H read of samples + energies inside the timed region
  roofline     the tcgen05 implicit-GEMM kernel: algorithmic FLOPs / CUDA-event time of every GEMM launch, measured
               live in a separate pass of the same workload, against MEASURED_PEAKS.json
  cpu_baseline the CPU oracle port of the same rollout on the host cores (bounded sample)
--impl reference times the reference algorithm's CPU implementation (oracle port, all host threads) per step.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "dxmi_rollout_images_per_sec"
UNIT = "images/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


GF = {"cifar": (12.444, 1.613), "in64": (219.314, 6.454), "lsun": (2238.707, 0.0)}  # algorithmic GFLOP / image: U-Net forward, value net
SHAPE = {"cifar": (3, 32, 32), "in64": (3, 64, 64), "lsun": (3, 256, 256)}
DEFAULTS = {"cifar": (4, 256), "in64": (10, 64), "lsun": (4, 64), "c4": (10, 128)}  # (T, images per GPU per step)


def workload_name(wl, T, B):
    if wl == "cifar":
        return (f"CIFAR-10 DDPM U-Net (in-tree stand-in for the absent DDGAN backbone) DxMI T={T} sampler rollout + energy "
                f"eval, batch {B}/GPU, bf16 tcgen05 (BASELINE.json configs[1])")
    if wl == "lsun":
        return (f"LSUN Bedroom 256 EDM U-Net (models/cm, unconditional) DxMI T={T} ancestral sampler rollout (rho=4, stochastic "
                f"last step), batch {B}/GPU, bf16 tcgen05 (BASELINE.json configs[4]; no energy eval: its value net is not in "
                f"the reference tree, SURVEY F8)")
    return (f"ImageNet64 EDM U-Net (models/cm, class-conditional) DxMI T={T} ancestral sampler rollout + energy eval, batch "
            f"{B}/GPU, bf16 tcgen05 (BASELINE.json configs[2], per-GPU shard of the 512-image batch)")


def cpu_oracle_rollout_fn(wl, T):
    """The reference algorithm on the CPU (oracle port): returns f(B, seed) -> seconds for one rollout + energy."""
    import torch

    from oracle import nets, samplers, synth

    torch.set_num_threads(os.cpu_count() or 1)
    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "ddpm_shapes.json")))
    vsd = synth.synth_state_dict({k: tuple(v) for k, v in shapes["value"].items()}, seed=1)
    if wl == "cifar":
        sd = synth.synth_state_dict({k: tuple(v) for k, v in shapes["net"].items()})
        sched = samplers.var_schedule(T)
        log_betas = sched["log_betas_init"]

        def run(B, seed=0):
            noise = synth.synth_noise(T, B, SHAPE[wl], seed=seed)
            t0 = time.perf_counter()
            with torch.no_grad():
                d = samplers.var_rollout(lambda x, t: nets.ddpm_unet_forward(sd, x, t), sched, log_betas, noise)
                nets.value_forward(vsd, d["sample"])
            return time.perf_counter() - t0

        return run
    from common import EDM_IN64_CFG, adm_oracle_kwargs
    from diffusion_by_maxentirl_b200.models.cm.script_util import create_model_and_diffusion

    unet, _ = create_model_and_diffusion(**EDM_IN64_CFG)  # parameter shapes only (no CUDA call)
    sd = synth.synth_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()})
    sd = {k: (v[..., None] if v.dim() == 3 else v) for k, v in sd.items()}
    akw = adm_oracle_kwargs(EDM_IN64_CFG)
    sched = samplers.edm_schedule(T)
    log_betas = sched["log_betas_init"]

    def run(B, seed=0):
        noise = synth.synth_noise(T, B, SHAPE[wl], seed=seed)
        noise[0] = noise[0] * 80.0
        y = synth.synth_labels(B, seed=seed)
        t0 = time.perf_counter()
        with torch.no_grad():
            d = samplers.edm_rollout(lambda x, t, yy: nets.adm_unet_forward(sd, x, t, yy, fp16_torso=False, **akw), sched,
                                     log_betas, noise, y)
            nets.value_forward(vsd, d["sample"])
        return time.perf_counter() - t0

    return run


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    wl = args.workload
    T = args.T or DEFAULTS[wl][0]
    Bs = 8 if wl == "cifar" else 1
    run = cpu_oracle_rollout_fn(wl, T)
    for _ in range(max(1, min(args.warmup, 2)) if wl == "cifar" else 0):
        run(Bs)
    steps = max(1, min(args.steps, 10 if wl == "cifar" else 2))
    t = [run(Bs, seed=i) for i in range(steps)]
    tot = sum(t)
    v = Bs * steps / tot
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(wl, T, Bs) + " - CPU oracle port of the reference algorithm, bounded sample of "
                               f"{Bs} images per step", "T": T, "batch_per_step": Bs},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} rollouts x {Bs} images, torch {torch.__version__} CPU fp32, {cores} threads"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def eager_cuda_baseline(wl, T, B, dev):
    """The honest bar: the reference's arithmetic in eager PyTorch on the same GPU - the oracle's functional
    networks moved to CUDA (cuDNN / cuBLAS / ATen), best of fp32-with-TF32 and autocast-bf16 for the DDPM U-Net, the
    reference's own fp16 torso for the ADM U-Net. Baseline only: nothing here is on the product path."""
    import torch

    from oracle import nets, samplers, synth

    shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "ddpm_shapes.json")))
    vsd = {k: v.to(dev) for k, v in synth.synth_state_dict({k: tuple(v) for k, v in shapes["value"].items()}, seed=1).items()}
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    noise = [torch.randn(B, *SHAPE[wl], device=dev) for _ in range(T + 1)]
    if wl == "cifar":
        sd = {k: v.to(dev) for k, v in synth.synth_state_dict({k: tuple(v) for k, v in shapes["net"].items()}).items()}
        sched = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in samplers.var_schedule(T).items()}
        lb = sched["log_betas_init"]

        def work(autocast):
            def net(x, t):
                if autocast:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return nets.ddpm_unet_forward(sd, x, t).float()
                return nets.ddpm_unet_forward(sd, x, t)

            with torch.no_grad():
                d = samplers.var_rollout(net, sched, lb, noise)
                return nets.value_forward(vsd, d["sample"])

        modes = (("fp32/TF32", False), ("autocast-bf16", True))
    else:
        from common import EDM_IN64_CFG, adm_oracle_kwargs, build_edm

        # synthetic weights with the reference's convert_to_fp16() applied (fp16 torso convs incl. biases; fp32 norms / embeddings)
        unet, _, _ = build_edm(EDM_IN64_CFG, T, device=dev)
        sd = {k: (v[..., None] if v.dim() == 3 else v).detach().clone() for k, v in unet.state_dict().items()}
        del unet
        akw = adm_oracle_kwargs(EDM_IN64_CFG)
        sched = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in samplers.edm_schedule(T).items()}
        lb = sched["log_betas_init"]
        y = torch.randint(0, 1000, (B,), device=dev)
        noise[0] = noise[0] * 80.0

        def work(autocast):
            with torch.no_grad():
                d = samplers.edm_rollout(lambda x, t, yy: nets.adm_unet_forward(sd, x, t, yy, fp16_torso=True, **akw), sched, lb, noise, y)
                return nets.value_forward(vsd, d["sample"])

        modes = (("fp16 torso (reference convert_to_fp16)", False),)
    best = None
    with torch.device(dev):  # the oracle's factory calls (torch.ones, arange, ...) land on the GPU
        for name, ac in modes:
            for _ in range(2):
                work(ac)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 3
            e0.record()
            for _ in range(n):
                work(ac)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            if best is None or ms < best[1]:
                best = (name, ms)
    del sd, vsd
    torch.cuda.empty_cache()
    return {"value": B / best[1] * 1e3, "unit": UNIT, "ms_per_step": best[1], "mode": best[0],
            "what": f"oracle functional networks in eager PyTorch {torch.__version__} on the same GPU (cuDNN/cuBLAS), same rollout, "
                    f"batch {B}, device-timed, 3 steps after 2 warm-ups"}


def traffic_table():
    """dram__bytes per launch of the hot kernels from the committed ncu --set full capture (profiles/r02_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def measure(wl, T, B, K, W, args, ctx, full):
    """One workload on this rank: device-timed value, e2e through the public API, roofline legs. `full`: the main line
    (clock sampling, CPU baseline); secondary workloads use fewer steps."""
    import torch
    import torch.distributed as dist

    from common import EDM_IN64_CFG, EDM_LSUN_CFG, VALUE_CFG, build_ddpm, build_edm, load_synth_into
    from diffusion_by_maxentirl_b200 import _lib as L
    from diffusion_by_maxentirl_b200.dist import PackedRollout

    rank, world, dev, lib = ctx["rank"], ctx["world"], ctx["dev"], ctx["lib"]
    shape = SHAPE[wl]
    g = torch.Generator().manual_seed(1234 + rank)
    if wl == "cifar":
        net, sampler, value, sd, vsd = build_ddpm(T, device=dev)
        labels = None

        @torch.no_grad()
        def rollout(noise):  # noise [T+1, B, C, H, W] on the device
            d = sampler.sample(noise.shape[1], device=dev, noise=noise)
            return d, value(d["sample"], T)
    elif wl == "lsun":
        net, sampler, sd = build_edm(EDM_LSUN_CFG, T, device=dev, stochastic_last=True, rho=4.0)
        value, labels = None, None

        @torch.no_grad()
        def rollout(noise):
            return sampler.sample(noise.shape[1], device=dev, x0=noise[0] * 80.0, noise=noise[1:]), None
    else:
        from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2
        from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue

        net, sampler, sd = build_edm(EDM_IN64_CFG, T, device=dev)
        value = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
        load_synth_into(value, seed=1)
        value.to(dev).eval()
        labels = torch.randint(0, 1000, (B,), generator=g).to(dev)

        @torch.no_grad()
        def rollout(noise):  # noise[0] is x_0 / sigma_max
            d = sampler.sample(noise.shape[1], device=dev, i_class=labels[:noise.shape[1]], x0=noise[0] * 80.0, noise=noise[1:])
            return d, value(d["sample"], T)

    eager_rollout = rollout
    # N > 1: the last transition kernel writes u8 samples and the value head the energies into ONE packed buffer
    packed = PackedRollout(B, shape, dev, world=world, with_energy=value is not None) if world > 1 else None
    graph_gather = bool(int(os.environ.get("DXMI_GRAPH_GATHER", "0")))
    if not args.no_graph:
        # capture the public-API call once (diffusion_by_maxentirl_b200.graph.GraphedRollout) and replay it per step
        from diffusion_by_maxentirl_b200.graph import GraphedRollout

        rollout = GraphedRollout(sampler, B, dev, value=value, labels=labels, packed=packed,
                                 capture_gather=graph_gather)  # same signature: rollout(noise) -> (d_sample, energies)

    n_host_bufs = 2
    host_noise = [torch.randn(T + 1, B, *shape, generator=g).pin_memory() for _ in range(n_host_bufs)]
    dev_noise = host_noise[0].to(dev)
    flush = ctx["flush"]

    def gather(d, e):
        # the path's only collective: ONE all-gather of the packed (u8 samples | fp32 energies) buffer (generate_large.py:43-50)
        if args.no_graph:
            from diffusion_by_maxentirl_b200.dist import gather_packed

            return gather_packed(d["sample"], e)
        if graph_gather:
            return packed.unpack()  # the collective ran inside the graph
        return packed.all_gather()

    def rollout_resident():
        d, e = rollout(dev_noise)
        if world > 1:
            gather(d, e)
        return d, e

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- warm-up
    l0 = lib.dxmi_launch_count()
    eager_rollout(dev_noise)
    launches_per_rollout = lib.dxmi_launch_count() - l0  # kernels of ours in one rollout (a graph replay launches the same set)
    for _ in range(W):
        rollout_resident()
    sync_all()

    # ---------------------------------------------------------------- parity of the sharded run (N > 1): bit for bit
    shard_check = None
    if world > 1:
        d, e = rollout(dev_noise)
        s_all, e_all = gather(d, e)
        torch.cuda.synchronize()
        from diffusion_by_maxentirl_b200.dist import quantize_u8

        own = quantize_u8(d["sample"])
        ok = torch.equal(s_all[rank * B:(rank + 1) * B], own)
        if e is not None:
            ok = ok and torch.equal(e_all[rank * B:(rank + 1) * B], e.reshape(-1).float())
        # rank 0 re-runs rank 1's shard on ITS noise: N ranks' shards must equal a 1-GPU run on the same noise
        peer = 1
        gp = torch.Generator().manual_seed(1234 + peer)
        peer_labels = torch.randint(0, 1000, (B,), generator=gp) if wl == "in64" else None
        peer_noise = torch.randn(T + 1, B, *shape, generator=gp)
        ok1 = True
        if rank == 0:
            nb = min(B, 8)
            if wl == "in64":
                dd = sampler.sample(nb, device=dev, i_class=peer_labels[:nb].to(dev), x0=peer_noise[0, :nb].to(dev) * 80.0,
                                    noise=peer_noise[1:, :nb].contiguous().to(dev))
            elif wl == "lsun":
                dd = sampler.sample(nb, device=dev, x0=peer_noise[0, :nb].to(dev) * 80.0, noise=peer_noise[1:, :nb].contiguous().to(dev))
            else:
                dd = sampler.sample(nb, device=dev, noise=peer_noise[:, :nb].contiguous().to(dev))
            ok1 = torch.equal(quantize_u8(dd["sample"]), s_all[peer * B:peer * B + nb])
        flag = torch.tensor([int(ok), int(ok1)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        shard_check = {"gather_equals_local_outputs": bool(flag[0].item()),
                       "rank1_shard_equals_single_gpu_rerun_on_rank0": bool(flag[1].item())}
        assert flag.min().item() == 1, f"sharded rollout differs from the single-GPU run: {shard_check}"

    # ---------------------------------------------------------------- value: device-timed, inputs resident in HBM
    clocks = ClockSampler(ctx["local_rank"]) if (full and rank == 0) else None
    if clocks:
        clocks.start()
    evs = []
    sync_all()
    for _ in range(K):
        flush.fill_(1)  # L2 flush between timed iterations (untimed)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rollout_resident()
        b.record()
        evs.append((a, b))
    sync_all()
    launches = launches_per_rollout * K
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value_ips = world * B * K / (ms_total / 1e3)

    # ---------------------------------------------------------------- e2e: public API from pinned host noise
    # two-deep pipeline: the H2D copy of step i+1 (copy stream) overlaps rollout i; results are read back every step
    d2h_samples = torch.empty(B, *shape).pin_memory()
    d2h_energy = torch.empty(B, 1).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    dev_bufs = [torch.empty_like(dev_noise) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            dev_bufs[i % 2].copy_(host_noise[i % n_host_bufs], non_blocking=True)
            ready[i % 2].record(copy_stream)

    sync_all()
    for ev in consumed:
        ev.record(main_stream)
    t0 = time.perf_counter()
    prefetch(0)
    for i in range(K):
        if i + 1 < K:
            prefetch(i + 1)
        main_stream.wait_event(ready[i % 2])
        d, e = rollout(dev_bufs[i % 2])
        consumed[i % 2].record(main_stream)
        if world > 1:
            gather(d, e)
        d2h_samples.copy_(d["sample"], non_blocking=True)
        if e is not None:
            d2h_energy.copy_(e, non_blocking=True)
        torch.cuda.synchronize()  # the step's result is on the host before the next step is issued
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ips = world * B * K / float(t.item())
    clk = clocks.stop() if clocks else None

    # ---------------------------------------------------------------- rooflines (rank 0, live, eager launches on one stream)
    roof = None
    if rank == 0:
        pk, pk_kind = peaks()
        traffic = traffic_table()
        lib.dxmi_set_option(b"time_gemms", 1)
        kk = max(2, min(K, 4))
        eager_rollout(dev_noise)
        torch.cuda.synchronize()
        ms, fl, nl = C.c_double(), C.c_double(), C.c_longlong()
        L.check(lib.dxmi_gemm_timing(C.byref(ms), C.byref(fl), C.byref(nl)))  # drop the warm-up records
        for cat in (1, 2):
            L.check(lib.dxmi_aux_timing(cat, C.byref(ms), C.byref(fl), C.byref(nl)))
        for _ in range(kk):
            eager_rollout(dev_noise)  # per-launch CUDA events need eager launches
        torch.cuda.synchronize()
        L.check(lib.dxmi_gemm_timing(C.byref(ms), C.byref(fl), C.byref(nl)))
        achieved = fl.value / (ms.value * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"]
        step_ms = ms_total / K
        roof = {"bound": "tensor",
                "kernel": "every tcgen05 contraction launch of the step: conv_gemm2p_kernel / conv_gemm2_kernel (persistent implicit GEMM, "
                          "cta_group::2 pair and one-CTA variants) + the fused attention kernels (attnblk256_kernel / attn_fwd_kernel)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic.get("conv_gemm2p_kernel"),
                "peak_source": f"{pk_kind} MEASURED_PEAKS.json bf16_tflops_sustained",
                "launches_per_step": nl.value // kk, "ms_per_step": ms.value / kk,
                "share_of_step": (ms.value / kk) / step_ms, "algorithmic_gflop_per_step": fl.value / kk / 1e9,
                "whole_step_frac": None}
        model_gflop = (T * GF[wl][0] + (GF[wl][1] if value is not None else 0.0)) * B
        roof["whole_step_frac"] = (model_gflop / 1e3) / (step_ms * 1e-3) / peak
        # `frac` counts the FLOPs the tensor pipe EXECUTES (2 M N K of every launch).  The reference model's arithmetic (GF per
        # image) is larger: the nearest-2x upsample + 3x3 conv layers run as four 2x2 phase convolutions at 4/9 of the reference FLOPs
        # (gemm up2 mode).  `whole_step_frac` divides the REFERENCE arithmetic by the step time.
        roof["reference_model_gflop_per_step"] = model_gflop
        roof["note"] = ("frac = executed tcgen05 FLOPs / kernel time / peak; whole_step_frac = reference-model FLOPs / step time / peak "
                        "(upsample+conv layers execute 4/9 of their reference FLOPs as phase convolutions)")
        hbm = []
        for cat, name, key in ((1, "GroupNorm family: gn_finalize_k + gn_apply_ab_k / gn_apply_fused_k (normalise + SiLU, one bf16 read + "
                                   "one bf16 write per element)", "gn_apply_ab_k"),
                               (2, "transition step var_step_k / edm_step_k (x' = mu + sigma z, logp; fp32 tensors)", "step_k")):
            bms, by, bn = C.c_double(), C.c_double(), C.c_longlong()
            L.check(lib.dxmi_aux_timing(cat, C.byref(bms), C.byref(by), C.byref(bn)))
            if bn.value:
                ach = by.value / (bms.value * 1e-3) / 1e9
                hbm.append({"bound": "hbm", "kernel": name, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": ach / pk["hbm_gbs"], "traffic": traffic.get(key), "launches_per_step": bn.value // kk,
                            "ms_per_step": bms.value / kk, "share_of_step": (bms.value / kk) / step_ms,
                            "algorithmic_mb_per_step": by.value / kk / 1e6})
        roof["hbm_kernels"] = hbm
        lib.dxmi_set_option(b"time_gemms", 0)

    res = {"wl": wl, "T": T, "B": B, "K": K, "W": W, "value": value_ips, "ms_per_step": ms_total / K, "e2e": e2e_ips,
           "h2d": host_noise[0].numel() * 4, "d2h": d2h_samples.numel() * 4 + (d2h_energy.numel() * 4 if value is not None else 0),
           "launches": int(launches), "clocks": clk, "roofline": roof, "shard_check": shard_check, "has_value": value is not None}
    # free this workload's plans / graphs before the next one is built
    del rollout, eager_rollout, sampler, net, value
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    return res


def measure_c4(B, K, W, args, ctx):
    """BASELINE.json configs[3]: one DxMI training iteration of the CIFAR-10 DDPM T=10 configuration, batch B per GPU, under
    DistributedDataParallel when N > 1 (trainer.py:230-408 / train_cifar10.py:141-205): rollout (eval, no grad) -> energy update of
    the value net on cat(real, x_T) -> T TD updates of the value net -> sampler update (train mode, dropout 0.1, sample_step with
    grad on B buffer rows; value term + running cost - entropy).  Clip-by-global-norm 0.1 + Adam run as the fused multi-tensor
    kernels (train_ops.FusedAdam), the running cost as its fused forward / backward kernel.  Synthetic "real" images."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from torch.nn.parallel import DistributedDataParallel as DDP

    from common import DDPM_CFG, VALUE_CFG, load_synth_into
    from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model
    from diffusion_by_maxentirl_b200.models.DxMI.var_sampler import VARSampler
    from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2
    from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue
    from diffusion_by_maxentirl_b200.train_ops import FusedAdam, running_cost

    rank, world, dev, lib = ctx["rank"], ctx["world"], ctx["dev"], ctx["lib"]
    T = 10
    net = Model(**DDPM_CFG)  # dropout 0.1 as in configs/cifar10/T10.yaml
    sampler = VARSampler(net, n_timesteps=T, sample_shape=[3, 32, 32], trainable_beta="fix_last")
    load_synth_into(net)
    sampler.to(dev)
    v = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
    load_synth_into(v, seed=1)
    v.to(dev)
    inner_net = net
    if world > 1:
        sampler.net = DDP(sampler.net, device_ids=[dev.index], output_device=dev.index)  # train_cifar10.py:303-309
        v = DDP(v, device_ids=[dev.index], output_device=dev.index)
    params_not_beta = [p for n_, p in inner_net.named_parameters() if "log_betas" not in n_]
    opt_s = FusedAdam([{"params": [inner_net.log_betas], "lr": 1e-4}, {"params": params_not_beta, "lr": 1e-6}])  # train_cifar10.py:287-290
    opt_v = FusedAdam(v.parameters(), lr=1e-5)
    g = torch.Generator().manual_seed(77 + rank)
    images = (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).to(dev)
    betas_q = torch.linspace(1e-4, 2e-2, T).to(dev)
    tau1, tau2 = 0.01, 0.1
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

    from diffusion_by_maxentirl_b200.graph import GraphedRollout

    sampler.eval()
    # the no-grad rollout is a static launch list: replayed as ONE CUDA graph that follows the optimizer steps (live=True re-packs the
    # bf16 operands in place before each replay; the learned sigmas are recomputed from log_betas inside the graph)
    rollout = None if args.no_graph else GraphedRollout(sampler, B, dev, live=True)

    def iteration(timed):
        if timed:
            ev[0].record()
        sampler.eval()
        if rollout is not None:
            d, _ = rollout()
        else:
            with torch.no_grad():
                d = sampler.sample(B, device=dev)
        if timed:
            ev[1].record()
        xs = torch.stack(d["l_sample"])  # [T+1, B, ...]
        out = v(torch.cat([images, xs[-1]]), T)  # energy update (trainer.py:244-264)
        pos, neg = out[:B], out[B:]
        d_loss = pos.mean() - neg.mean() + 0.05 * ((pos ** 2).mean() + (neg ** 2).mean())
        opt_v.zero_grad(set_to_none=True)
        d_loss.backward()
        opt_v.step()
        if timed:
            ev[2].record()
        for i in range(T):  # T TD updates (trainer.py:276-326)
            tt = T - 1 - i
            state, nxt = xs[tt], xs[tt + 1]
            tvec = torch.full((B,), tt, device=dev, dtype=torch.long)
            with torch.no_grad():
                rc = running_cost(state, nxt, betas_q[T - tvec - 1])
                target = v(nxt, tt + 1).flatten() + tau2 * rc - tau1 * torch.log(d["sigma"][tt].flatten())
            v_loss = F.mse_loss(v(state, tt).flatten(), target)
            opt_v.zero_grad(set_to_none=True)
            v_loss.backward()
            opt_v.step(max_norm=0.1)
        if timed:
            ev[3].record()
        sampler.train()  # sampler update (trainer.py:348-389)
        idx_t = torch.randint(0, T, (B,), device=dev)
        state = xs[idx_t, torch.arange(B, device=dev)]
        ds = sampler.sample_step(state, idx_t)
        nt = (idx_t < T - 1).float()
        rc = running_cost(state, ds["sample"], betas_q[T - idx_t - 1])
        for p_ in v.parameters():
            p_.requires_grad_(False)
        s_loss = (v(ds["sample"], idx_t + 1).flatten() + (tau2 * rc - tau1 * ds["entropy"].flatten()) * nt).mean()
        for p_ in v.parameters():
            p_.requires_grad_(True)
        opt_s.zero_grad(set_to_none=True)
        s_loss.backward()
        opt_s.step(max_norm=0.1)
        if timed:
            ev[4].record()
        return d_loss, v_loss, s_loss

    l0 = lib.dxmi_launch_count()
    iteration(False)
    launches_per_iter = lib.dxmi_launch_count() - l0
    for _ in range(max(W - 1, 1)):
        iteration(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    acc = [0.0] * 4
    for _ in range(K):
        losses = iteration(True)
        torch.cuda.synchronize()
        for j in range(4):
            acc[j] += ev[j].elapsed_time(ev[j + 1])
    tot = torch.tensor([sum(acc)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms = float(tot.item()) / K
    res = {"metric": "dxmi_training_images_per_sec", "value": world * B / ms * 1e3, "unit": UNIT, "steps": K, "warmup": W,
           "ms_per_step": ms,
           "phases_ms": {"rollout": acc[0] / K, "energy_update": acc[1] / K, "td_updates_x10": acc[2] / K, "sampler_update": acc[3] / K},
           "model_tflops": 30.8 * B / 128 * world / ms * 1e3,  # 30.8 TFLOP per iteration per GPU at B = 128
           "config": {"workload": f"CIFAR-10 DDPM T=10 DxMI training iteration (rollout + energy update + 10 TD updates + sampler update, "
                                  f"dropout 0.1), batch {B}/GPU, bf16 tcgen05 forward + backward (BASELINE.json configs[3])",
                      "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"ddp{world}",
                      "collective": ("DDP gradient all-reduce (NCCL): 1 x 143 MB U-Net + 11 x 20.5 MB value net per iteration" if world > 1 else "none"),
                      "launch": ("rollout: CUDA graph replay (graph.GraphedRollout(live=True)); " if rollout is not None else "rollout: eager; ") +
                                "value / U-Net training steps: eager (autograd)", "optimizer": "train_ops.FusedAdam (clip 0.1 folded in)"},
           "gpu_launches": int(launches_per_iter * K), "losses": [float(x) for x in losses]}
    del rollout, sampler, v, net, opt_s, opt_v
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cifar", choices=["cifar", "in64", "lsun", "c4"],
                    help="cifar = BASELINE configs[1] (driver default); in64 = configs[2] per-GPU shard (ImageNet-64 EDM T=10); "
                         "lsun = configs[4] per-GPU shard (LSUN-256 EDM T=4, sampler only)")
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step")
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the north-star target workloads in `secondary`")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--block-n-256", type=int, default=None)
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=INT", help="dxmi_set_option before the plans are built (A/B switches)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from diffusion_by_maxentirl_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at the first collective; stdout must carry the JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    if args.block_n_256 is not None:
        lib.dxmi_set_option(b"block_n_256", args.block_n_256)
    for kv in args.opt:
        name, val = kv.split("=")
        L.check(lib.dxmi_set_option(name.encode(), int(val)), f"option {name}")

    wl = args.workload
    T = args.T or DEFAULTS[wl][0]
    B = args.batch or DEFAULTS[wl][1]
    K, W = args.steps, args.warmup
    ctx = {"rank": rank, "local_rank": local_rank, "world": world, "dev": dev, "lib": lib,
           "flush": torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)}  # > 126 MB L2

    if wl == "c4":  # the training configuration on its own (also part of `secondary` of the default run)
        r = measure_c4(args.batch or 128, max(2, min(K, 5)), 3, args, ctx)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
        if rank == 0:
            r.update({"n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic"})
            print(json.dumps(r), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    main_res = measure(wl, T, B, K, W, args, ctx, full=True)

    # ---------------------------------------------------------------- secondary: the north-star TARGET workloads
    secondary = []
    if not args.no_secondary and wl == "cifar" and args.T is None and args.batch is None:
        for swl, sT, sB, sK in (("cifar", 10, 256, max(3, min(K, 8))), ("in64", 10, 64, max(3, min(K, 5)))):
            r = measure(swl, sT, sB, sK, 3, args, ctx, full=False)
            secondary.append(r)

    c4 = None
    if not args.no_secondary and wl == "cifar" and args.T is None and args.batch is None:
        c4 = measure_c4(128, 3, 3, args, ctx)

    # ---------------------------------------------------------------- eager-CUDA bar + CPU baseline (rank 0, N = 1 only)
    eager = None
    cpu = None
    if rank == 0 and world == 1 and wl != "lsun":
        if not args.no_eager_baseline:
            eager = eager_cuda_baseline(wl, T, B, dev)
            for r in secondary:
                r["eager"] = eager_cuda_baseline(r["wl"], r["T"], r["B"], dev)
        if not args.no_cpu_baseline:
            run = cpu_oracle_rollout_fn(wl, T)
            Bs = 8 if wl == "cifar" else 1
            if wl == "cifar":
                run(Bs)
            ts, t_begin = [], time.perf_counter()
            while len(ts) < (3 if wl == "cifar" else 1) or (time.perf_counter() - t_begin < 10 and len(ts) < 20):
                ts.append(run(Bs, seed=len(ts)))
            cores = torch.get_num_threads()
            cpu = {"value": Bs * len(ts) / sum(ts), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{len(ts)} rollouts x {Bs} images (T={T} + energy), oracle port, torch CPU fp32, {cores} threads; "
                             "images/s is per-image normalised (the CPU arm runs a bounded batch, not the GPU batch)"}

    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)

    def cfg(r):
        return {"workload": workload_name(r["wl"], r["T"], r["B"]), "T": r["T"], "batch_per_gpu": r["B"], "global_batch": r["B"] * world,
                "parallelism": f"dp{world}", "l2": "256 MiB write between timed steps (L2 flush)",
                "launch": "eager" if args.no_graph else "CUDA graph replay of the public-API call (graph.GraphedRollout)",
                "collective": "one all_gather of the packed (u8 samples | fp32 energies) buffer per step" if world > 1 else "none"}

    def e2e(r):
        return {"value": r["e2e"], "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                "pipeline": "H2D of step i+1 on a copy stream overlaps rollout i; samples + energies read back every step"}

    if rank == 0:
        r = main_res
        gf_u, gf_v = GF[wl]
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg(r), "e2e": e2e(r),
            "gpu_launches": r["launches"] + sum(x["launches"] for x in secondary) + (c4["gpu_launches"] if c4 else 0),
            "training_config": c4,
            "clocks": r["clocks"], "roofline": r["roofline"], "cpu_baseline": cpu, "eager_cuda_baseline": eager,
            "model_tflops": r["value"] * (T * gf_u + (gf_v if r["has_value"] else 0.0)) / 1e3,
            "shard_check": r["shard_check"],
            "secondary": [{"metric": METRIC, "value": x["value"], "unit": UNIT, "steps": x["K"], "warmup": x["W"],
                           "ms_per_step": x["ms_per_step"], "config": cfg(x), "e2e": e2e(x), "roofline": x["roofline"],
                           "eager_cuda_baseline": x.get("eager"), "shard_check": x["shard_check"],
                           "model_tflops": x["value"] * (x["T"] * GF[x["wl"]][0] + GF[x["wl"]][1]) / 1e3} for x in secondary],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
