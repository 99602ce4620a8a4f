#!/usr/bin/env python
"""bench.py -- throughput of the Escort direct-sparse-convolution hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload alexnet|googlenet|resnet50]

One "step" = one forward pass of every sparse conv layer of the workload over one batch of synthetic input
(BASELINE.json configs[1] by default: AlexNet conv2..conv5 pruned to 85-88 %, batch 256 per GPU).  With N > 1 the
batch is sharded: every rank runs its own 256 images, no data-path collective (inference), "scaling": "weak".

Our arm prints ONE JSON line with
  value     images/s with the inputs resident in HBM (CUDA events, max over ranks),
  parity    per layer: relative L2 of the TIMED plan's outputs (the autotuned kernel, 8 images) against the reference's
            own CPU kernels (oracle/_ref); the run exits non-zero above 1e-4,
  train     (default workload only) a short ResNet-50 forward + masked-backward step with the gradient exchange issued
            through the library (escort_allreduce_grads on a real ncclComm_t): both of the reference's modes -- per
            layer on a side stream, or one flat all-reduce after the backward -- are measured and the faster is timed,
  clocks    SM clock / throttle reasons polled through NVML every 5 ms during the timed region,
  e2e       images/s through the C-ABI with HOST buffers: pinned-host -> device copies of every layer input that
            comes from outside the path and device -> host copies of its outputs inside the timed region,
  roofline  the dominant kernel against max(nnz-FLOPs / FP32-FMA peak, compulsory bytes / HBM bandwidth),
  cpu_baseline  the reference's own CPU direct sparse conv (oracle/_ref, AVX2 + OpenMP) timed on this host.
`--impl reference` times that CPU path alone (rank 0 only) and prints the same line with "impl": "reference".

Only the cpu_baseline / --impl reference legs touch oracle/; the product path is the CUDA library through ctypes.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "sparse_conv_images_per_sec"
UNIT = "images/s"


def _peaks():
    """HBM peak from the driver-written MEASURED_PEAKS.json, else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _workload(name, train=False):
    from caffe_escoin_b200 import workloads as wl
    specs = {"alexnet": wl.ALEXNET, "googlenet": wl.GOOGLENET, "resnet50": wl.RESNET50, "lenet": wl.LENET}[name]
    label = {"alexnet": "alexnet_conv2-conv5_pruned85-88_fwd_b256",
             "googlenet": "googlenet_v1_3x3_5x5_pruned75_fwd_b128",
             "resnet50": "resnet50_branch2b_3x3_pruned70_fwd_b256",
             "lenet": "lenet5_conv1-2_pruned80_fwd_b64"}[name]
    if train:
        label = label.replace("_fwd_", "_fwd+masked-bwd_")
    return specs, label


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML -- the library behind
    nvidia-smi -- polled every 5 ms from a thread, so that a region of a few milliseconds still gets samples (an
    `nvidia-smi -lms` child needs ~100 ms to start and misses an 18 ms region); the nvidia-smi child remains the fallback
    where pynvml is missing.  stop() always returns at least one sample (taken right after the region if none fell in)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.proc = None
        self.thread = None
        self.samples = []     # (sm_mhz, max_mhz, reason bitmask)
        self._stop = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._sample()
            self.samples.clear()
            import threading
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.proc = None

    def _sample(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        self.samples.append((float(sm), float(mx), int(mask)))

    def _run(self):
        while not self._stop:
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.005)

    def _summary(self, sm, mx, reasons, n, note=None):
        busy = sorted(sm)[len(sm) // 2:]   # samples under load = the upper half (the sampler also sees the idle edges)
        out = {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": n}
        if note:
            out["note"] = note
        return out

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            note = None
            if not self.samples:
                try:
                    self._sample()
                    note = "no sample fell inside the region: one taken right after it"
                except Exception:
                    return None
            reasons = {n for n, bit in self.REASONS for s_ in self.samples if s_[2] & bit}
            return self._summary([s_[0] for s_ in self.samples], [s_[1] for s_ in self.samples], reasons,
                                 0 if note else len(self.samples), note)
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            return None
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return self._summary(sm, mx, reasons, len(sm))


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU kernels (oracle/_ref), all host threads
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_run(specs, sample, steps, warmup):
    """Times `steps` passes of the workload's layers over `sample` images with the reference's CPU direct sparse
    conv (sconv_unit_stride<W,K> where the reference has the specialisation, else caffe_cpu_sconv_default;
    pad copy included; omp-parallel over images).  Returns (images_per_s, ms_per_step, threads, kind)."""
    from caffe_escoin_b200 import workloads as wl
    from oracle import pyoracle as po
    po.build()
    kind = "reference" if po.have_ref() else "port"
    # all host cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which is not the reference's set-up)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    layers = []
    for li, spec in enumerate(specs):
        n = min(sample, spec.N)
        d = wl.make_layer_data(spec, li, N=n)
        g = po.Geom(n, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = po.weight_align(d["w"], g)
        layers.append((spec, g, csr, d))
    n = min(sample, specs[0].N)

    def step():
        for spec, g, csr, d in layers:
            if kind == "reference":
                po.ref_conv_forward(d["x"], csr, g, d["bias"], relu=True, threads=threads)
            else:
                po.conv_forward(d["x"], csr, g, d["bias"], relu=True, threads=threads)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps * 1e3, int(threads), kind


def _config(label, N, world):
    """The workload as both arms name it (identical dict: the driver compares the two lines)."""
    return {"workload": label, "batch_per_gpu": N, "global_batch": N * world}


def cpu_reference_bounded(specs, steps, warmup, budget_s=120.0):
    """One protocol for the reference arm and the cpu_baseline leg: the full batch per step unless the host is too slow
    for the budget (calibrated on a 16-image slice)."""
    sample = specs[0].N
    ips, _, threads, kind = cpu_reference_run(specs, min(16, sample), 1, 1)
    if sample / ips * (steps + warmup) > budget_s:
        sample = max(8, int(budget_s * ips / (steps + warmup)))
    ips, ms, threads, kind = cpu_reference_run(specs, sample, steps, warmup)
    desc = "%d of %d images per step x %d steps (%d warm-up) through all %d layers, %s" % (
        sample, specs[0].N, steps, warmup, len(specs),
        "reference AVX2 sconv_unit_stride / caffe_cpu_sconv_default + OpenMP over images" if kind == "reference"
        else "oracle C port + OpenMP over images")
    return ips, ms, threads, kind, desc


def run_reference(args, specs, label):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, ms, threads, kind, desc = cpu_reference_bounded(specs, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": _config(label, specs[0].N, args.gpus),
            "detail": {"device": "host CPU", "threads": threads},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def _pin_to_gpu_numa(local):
    """Run (and allocate pinned memory) on the cores NVML reports as local to this rank's GPU; returns the previous
    affinity so that the CPU baseline can use the whole host again."""
    try:
        prev = os.sched_getaffinity(0)
    except AttributeError:
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= prev
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    return prev


def _make_layers(specs, args, capi, wl, torch, train, tune_cache):
    """WeightAlign through the C ABI -> plan -> tuning (measured once per layer shape, copied to the other layers of
    the shape) -> device tensors.  Untimed set-up."""
    layers, tuned = [], {}
    dirty = False
    for li, spec in enumerate(specs):
        d = wl.make_layer_data(spec, li)
        geom = capi.make_geom(spec.Cin, spec.Cout, spec.H, spec.H, spec.k, spec.stride, spec.pad, 1, spec.group)
        csr = capi.weight_align(torch.from_numpy(d["w"]).cuda(), geom)
        plan = capi.Plan(geom, csr)
        shape = (spec.Cin, spec.Cout, spec.H, spec.k, spec.stride, spec.pad, spec.group, round(spec.sparsity, 3))
        if args.variant is not None:
            plan.set_variant(args.variant)
        elif shape in tuned:
            plan.copy_tuning(tuned[shape])
        elif spec.name in tune_cache and not train:
            plan.set_config(*tune_cache[spec.name])
            tuned[shape] = plan   # (the other layers of this shape copy it instead of measuring again)
        else:
            plan.autotune(spec.N)
            if train:
                plan.autotune_backward(spec.N)
            tuned[shape] = plan
            tune_cache[spec.name] = list(plan.get_config())
            dirty = True
        x = torch.from_numpy(d["x"]).cuda()
        b = torch.from_numpy(d["bias"]).cuda() if d["bias"] is not None else None
        y = torch.empty((spec.N, spec.Cout, plan.Ho, plan.Wo), device="cuda")
        flops, byts = wl.alg_work(spec, plan.nnz)
        L = dict(spec=spec, plan=plan, x=x, b=b, y=y, flops=flops, bytes=byts, csr=csr, w=d["w"], bias=d["bias"], li=li)
        if train:
            g = torch.Generator(device="cuda").manual_seed(1701 + li)
            L["dy"] = torch.rand(y.shape, device="cuda", generator=g) * 2 - 1
            L["dx"] = torch.empty_like(x)
            L["db"] = torch.zeros(spec.Cout, device="cuda") if b is not None else None
            # CSR-ordered gradient in the reference's blob layout: group g at offset weight_offset * g
            Mg, Cg = spec.Cout // spec.group, spec.Cin // spec.group
            wo = Mg * Cg * spec.k * spec.k
            rp = csr["rowptr"].cpu().numpy()
            L["gsize"] = wo * (spec.group - 1) + int(rp[(Mg + 1) * spec.group - 1])
        layers.append(L)
    return layers, dirty


def check_parity(layers, train, torch, capi):
    """Outside the timed region: the outputs of the plans that were just timed (same kernels, same tilings) against
    the reference's own CPU direct sparse conv on the first images of the batch (oracle/_ref; the C port if the
    reference is not built).  Masked backward: against the oracle's restatement (the reference's backward is dense)."""
    from oracle import pyoracle as po
    po.build()
    out = {}
    for L in layers:
        spec, plan = L["spec"], L["plan"]
        n = min(8, spec.N)
        g = po.Geom(n, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
        ocsr = po.weight_align(L["w"], g)
        x8 = L["x"][:n].cpu().numpy()
        relu = not train
        if po.have_ref():
            ref, _ = po.ref_conv_forward(x8, ocsr, g, L["bias"], relu=relu)
        else:
            ref = po.conv_forward(x8, ocsr, g, L["bias"], relu=relu)
        res = {"fwd": po.rel_l2(L["y"][:n].cpu().numpy(), ref)}   # L["y"]: written by the last timed step
        if train:
            nb = min(2, n)
            g2 = po.Geom(nb, spec.Cin, spec.H, spec.H, spec.Cout, spec.k, spec.stride, spec.pad, 1, spec.group)
            dy2 = L["dy"][:nb].contiguous()
            wd = torch.zeros(L["w"].shape, device="cuda")
            plan.backward_weight(L["x"][:nb].contiguous(), dy2, wd_dense=wd)
            dx = plan.backward_data(dy2)
            torch.cuda.synchronize()
            wd_o, _, dx_o = po.conv_backward(x8[:nb], dy2.cpu().numpy(), L["w"], g2, mask_only=True, want_b=False)
            res["bwd_weight"] = po.rel_l2(wd.cpu().numpy(), wd_o)
            res["bwd_data"] = po.rel_l2(dx.cpu().numpy(), dx_o)
        out[spec.name] = res
    worst = max(v for r in out.values() for v in r.values())
    return out, worst


def run_train_leg(args, world, rank, torch, dist, capi, wl, barrier):
    """BASELINE.json configs[3] under the driver's clock: ResNet-50 branch2b convs, forward + masked backward, the
    gradient exchange issued per layer on a side stream through the library's own entry point with a real ncclComm_t
    (NCCL<Dtype>::on_gradients_ready / run, src/caffe/parallel.cpp:202-256), weights broadcast once (:189-199)."""
    specs = wl.RESNET50
    layers, _ = _make_layers(specs, args, capi, wl, torch, True, {})
    N = specs[0].N
    offs, tot = [], 0
    for L in layers:
        offs.append(tot)
        tot += (L["gsize"] + 3) // 4 * 4
    flat = torch.zeros(tot, device="cuda")   # the flat diff buffer of the exchange (parallel.cpp:75-108)
    for L, o in zip(layers, offs):
        L["wd"] = flat[o:o + L["gsize"]]
        L["wslice"] = flat[o:o + (L["gsize"] + 3) // 4 * 4]
    comm = None
    if world > 1:
        def exchange_id(raw):
            t = torch.tensor(list(raw) if raw is not None else [0] * 128, dtype=torch.uint8, device="cuda")
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        comm = capi.NcclComm(world, rank, exchange_id)
        # weights from rank 0 once, then the plans re-gather their values (every rank generated the same weights; the
        # broadcast is the reference's start-up step, not a correction)
        for L in layers:
            wdev = torch.from_numpy(L["w"]).cuda()
            capi.broadcast(wdev, 0, comm)
            L["plan"].refresh_values(wdev)
        torch.cuda.synchronize()
    side = torch.cuda.Stream()
    main = torch.cuda.current_stream()
    ev_bwd = torch.cuda.Event(enable_timing=True)
    ev_all = torch.cuda.Event(enable_timing=True)

    def step(mark=False, layerwise=True):
        for L in layers:
            L["plan"].forward(L["x"], L["b"], relu=False, top=L["y"])
        for L in reversed(layers):
            L["plan"].backward_weight(L["x"], L["dy"], wd_csr=L["wd"], accumulate=False)
            if L["db"] is not None:
                capi.bias_backward(L["dy"], L["db"])
            if comm is not None and layerwise:
                # this layer's gradient is final: reduce it on the side stream while the next layers' backward runs
                # (solver param layer_wise_reduce: true -> NCCL::run(layer), parallel.cpp:202-235)
                e = torch.cuda.Event()
                e.record(main)
                side.wait_event(e)
                capi.allreduce_grads(L["wslice"], 1.0 / world, comm, stream=side)
            L["plan"].backward_data(L["dy"], L["dx"])
        if mark:
            ev_bwd.record(main)
        if comm is not None:
            if layerwise:
                main.wait_stream(side)
            else:
                # layer_wise_reduce: false -> one all-reduce of the flat diff buffer (on_gradients_ready, parallel.cpp:246-255)
                capi.allreduce_grads(flat, 1.0 / world, comm)
        if mark:
            ev_all.record(main)

    def timed(n, layerwise, mark_last=False):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for s in range(n):
            step(mark=(mark_last and s == n - 1), layerwise=layerwise)
        t1.record()
        barrier()
        ms_ = t0.elapsed_time(t1) / n
        if world > 1:
            t = torch.tensor([ms_], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t[0].item())
        return ms_

    steps, warm = 3, 2
    for _ in range(warm):
        step()
    # Both of the reference's exchange modes are measured (2 steps each, max over ranks) and the faster one is timed:
    # the CSR-ordered gradients are small (13.6 MB for the 16 layers), and a persistent compute kernel that finds an SM
    # taken by an NCCL kernel runs one CTA late, so "overlap" can cost more than the flat exchange it hides.
    mode_ms = {"layerwise": None, "flat": None}
    layerwise = True
    if comm is not None:
        step(layerwise=False)
        mode_ms["layerwise"] = timed(2, True)
        mode_ms["flat"] = timed(2, False)
        layerwise = mode_ms["layerwise"] <= mode_ms["flat"]
    ms = timed(steps, layerwise, mark_last=True)
    exposed = ev_bwd.elapsed_time(ev_all)
    # the exchange alone (same buffers, nothing to overlap with): bus bandwidth of the flat all-reduce
    bus = None
    exch_ms = 0.0
    if comm is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        capi.allreduce_grads(flat, 1.0, comm)
        barrier()
        e0.record()
        for _ in range(5):
            capi.allreduce_grads(flat, 1.0, comm)
        e1.record()
        torch.cuda.synchronize()
        exch_ms = e0.elapsed_time(e1) / 5
        bus = 2.0 * (world - 1) / world * flat.numel() * 4 / (exch_ms * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([exposed], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exposed = float(t[0].item())
    parity, worst = check_parity(layers, True, torch, capi) if rank == 0 else ({}, 0.0)
    if comm is not None:
        torch.cuda.synchronize()
        comm.destroy()
    flops = sum(L["flops"] for L in layers) * 3
    return {"workload": "resnet50_branch2b_3x3_pruned70_fwd+masked-bwd_b256", "value": world * N / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms, "steps": steps, "warmup": warm, "tflops": flops / ms / 1e9,
            "exchange": {"through": ("escort_allreduce_grads(ncclComm_t) " + ("per layer on a side stream" if layerwise else
                                     "once on the flat CSR-ordered diff buffer after the backward") + ", 1/N fused behind it")
                                    if comm is not None else "single rank: no exchange",
                         "mode": ("layerwise" if layerwise else "flat") if comm is not None else None,
                         "mode_ms_per_step": mode_ms,
                         "allreduce_bytes_per_step": flat.numel() * 4 if comm is not None else 0,
                         "exposed_ms": exposed if comm is not None else 0.0,
                         "standalone_ms": exch_ms, "bus_gbs": bus},
            "kernels": {"fwd": layers[0]["plan"].kernel_names()["fwd"], "bwd_data": layers[0]["plan"].kernel_names()["bwd_data"],
                        "bwd_weight": layers[0]["plan"].kernel_names()["bwd_weight"]},
            "parity": parity, "parity_worst": worst}


def run_ours(args, specs, label):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the sparse-conv path has no CPU fallback "
                         "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    full_affinity = _pin_to_gpu_numa(local)   # before any pinned allocation: host buffers local to this rank's GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the first communicator comes up; stdout carries ONE JSON line,
        # so the bring-up (init + first collective) runs with fd 1 pointed at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    if not os.path.exists(os.path.join(ROOT, "caffe_escoin_b200", "libescort_b200.so")):
        if rank == 0:
            ge.build()
        if world > 1:
            dist.barrier()
    from caffe_escoin_b200 import capi, workloads as wl

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP32 FMA peak, measured live by this run (builder-measured: MEASURED_PEAKS.json has no fp32 entry) ----
    fp32_peaks = {}
    for v, name in ((0, "ffma_shared_operand"), (1, "ffma2_packed"), (2, "ffma_3reg")):
        fp32_peaks[name] = capi.measure_fp32_peak(v, 8192)[0]
    fp32_peak = max(fp32_peaks.values())
    hbm_peak, hbm_src = _peaks()

    # ---- set-up (untimed) ----
    tune_cache = {}
    if args.tune_cache and os.path.exists(args.tune_cache):
        tune_cache = json.load(open(args.tune_cache))
    layers, tune_dirty = _make_layers(specs, args, capi, wl, torch, args.train, tune_cache)
    if args.tune_cache and tune_dirty and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(args.tune_cache)), exist_ok=True)
        json.dump(tune_cache, open(args.tune_cache, "w"))
    N = specs[0].N
    nl = len(layers)
    host_in, host_out = [], []
    for L in layers:
        host_in.append(L["x"].cpu().pin_memory())
        host_out.append(torch.empty(L["y"].shape, dtype=torch.float32).pin_memory())
        if args.train:
            host_in.append(L["dy"].cpu().pin_memory())
            host_out.append(torch.empty(L["x"].shape, dtype=torch.float32).pin_memory())

    # the step as a list of timed operations: (layer index, kind, callable); every operation is one launch of ours
    flat = None
    if args.train:
        offs, tot = [], 0
        for Lr in layers:
            offs.append(tot)
            tot += (Lr["gsize"] + 3) // 4 * 4
        flat = torch.zeros(tot, device="cuda")   # the flat gradient buffer the exchange step reduces (parallel.cpp:75-108)
        for Lr, o in zip(layers, offs):
            Lr["wd"] = flat[o:o + Lr["gsize"]]
    ops = []
    for i, L in enumerate(layers):
        if not args.train:
            ops.append((i, "fwd", lambda L=L: L["plan"].forward(L["x"], L["b"], relu=True, top=L["y"])))
        else:
            ops.append((i, "fwd", lambda L=L: L["plan"].forward(L["x"], L["b"], relu=False, top=L["y"])))
            ops.append((i, "bwd_weight", lambda L=L: L["plan"].backward_weight(L["x"], L["dy"], wd_csr=L["wd"],
                                                                               accumulate=False)))
            ops.append((i, "bwd_data", lambda L=L: L["plan"].backward_data(L["dy"], L["dx"])))
    nops = len(ops)
    comm = None
    if args.train and world > 1:
        def exchange_id(raw):
            t = torch.tensor(list(raw) if raw is not None else [0] * 128, dtype=torch.uint8, device="cuda")
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        comm = capi.NcclComm(world, rank, exchange_id)

    def exchange():
        # the path's one exchange step: sum over ranks, then 1/N (NCCL<Dtype>::on_gradients_ready, parallel.cpp:238-256),
        # through the library's own entry point on a real ncclComm_t
        if not args.train:
            return
        for L in layers:
            if L["db"] is not None:
                capi.bias_backward(L["dy"], L["db"])
        if comm is not None:
            capi.allreduce_grads(flat, 1.0 / world, comm)

    def step(events=None):
        for k, (i, kind, fn) in enumerate(ops):
            if events is not None:
                events[k][0].record()
            fn()
            if events is not None:
                events[k][1].record()
        exchange()

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nops)]
           for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for s in range(args.steps):
        step(evs[s])
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * N * args.steps / (ms_total * 1e-3)
    op_ms = [float(np.mean([evs[s][k][0].elapsed_time(evs[s][k][1]) for s in range(args.steps)])) for k in range(nops)]

    # ---- parity of what was just timed (rank 0; outside the timed region) ----
    parity, parity_worst = (check_parity(layers, args.train, torch, capi) if rank == 0 and not args.no_parity else ({}, 0.0))

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region, chunk-pipelined over streams ----
    nchunk = 4 if N % 4 == 0 else 1
    cs = N // nchunk
    streams = [torch.cuda.Stream() for _ in range(4)]
    h2d = sum(h.numel() * 4 for h in host_in)
    d2h = sum(h.numel() * 4 for h in host_out)

    if args.train:
        h2d = sum(L["x"].numel() * 4 + L["dy"].numel() * 4 for L in layers)
        d2h = sum(L["dx"].numel() * 4 for L in layers) + flat.numel() * 4
        host_flat = torch.empty(flat.shape, dtype=torch.float32).pin_memory()

    def e2e_step(kernels=True):
        if args.train:
            # host bottom + top_diff in, bottom_diff + the reduced flat gradient out; one stream per layer, round robin
            for i, L in enumerate(layers):
                st = streams[i % len(streams)]
                with torch.cuda.stream(st):
                    L["x"].copy_(host_in[2 * i], non_blocking=True)
                    L["dy"].copy_(host_in[2 * i + 1], non_blocking=True)
                    if kernels:
                        L["plan"].forward(L["x"], L["b"], relu=False, top=L["y"], stream=st)
                        L["plan"].backward_weight(L["x"], L["dy"], wd_csr=L["wd"], accumulate=False, stream=st)
                        L["plan"].backward_data(L["dy"], L["dx"], stream=st)
                    host_out[2 * i + 1].copy_(L["dx"], non_blocking=True)
            for st in streams:
                st.synchronize()
            if kernels:
                exchange()
            host_flat.copy_(flat, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return
        k = 0
        for i, L in enumerate(layers):
            for c in range(nchunk):
                st = streams[k % len(streams)]
                k += 1
                with torch.cuda.stream(st):
                    sl = slice(c * cs, (c + 1) * cs)
                    L["x"][sl].copy_(host_in[i][sl], non_blocking=True)
                    if kernels:
                        L["plan"].forward(L["x"][sl], L["b"], relu=True, top=L["y"][sl], stream=st)
                    host_out[i][sl].copy_(L["y"][sl], non_blocking=True)
        for st in streams:
            st.synchronize()

    def timed_e2e(kernels):
        for _ in range(max(1, min(args.warmup, 3))):
            e2e_step(kernels)
        barrier()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step(kernels)
        barrier()
        dt = time.perf_counter() - w0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * N * args.steps / dt

    e2e_value = timed_e2e(True)
    copy_ceiling = timed_e2e(False)   # the same host <-> device bytes with no kernel at all: what the host links allow

    # ---- config 4 under the same clock: a short ResNet-50 training step with the library-issued exchange ----
    train_leg = None
    if not args.train and args.workload == "alexnet" and not args.no_train_leg and args.variant is None:
        for L in layers:          # free the forward workload's device tensors first
            L["x"] = L["y"] = None
        del host_in, host_out
        torch.cuda.empty_cache()
        train_leg = run_train_leg(args, world, rank, torch, dist, capi, wl, barrier)

    if comm is not None:
        torch.cuda.synchronize()
        comm.destroy()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of the step) ----
    kernel_of = {k: (lambda L, k=k: L["plan"].kernel_names()[k]) for k in ("fwd", "bwd_weight", "bwd_data")}
    dom = int(np.argmax(op_ms))
    L = layers[ops[dom][0]]
    dom_kernel = kernel_of[ops[dom][1]](L)
    t_meas = op_ms[dom] * 1e-3
    t_fma = L["flops"] / (fp32_peak * 1e12)
    t_hbm = L["bytes"] / (hbm_peak * 1e9)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(L["spec"].name, {}).get(dom_kernel)
        except Exception:
            traffic = None
    roofline = {"bound": "fp32_fma" if t_fma >= t_hbm else "hbm", "kernel": dom_kernel, "op": ops[dom][1],
                "layer": L["spec"].name, "achieved": L["flops"] / t_meas / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": max(t_fma, t_hbm) / t_meas,
                "peak_source": "FP32 FMA: builder-measured, live in this run, by escort_measure_fp32_peak (best of FFMA / "
                               "FFMA2 register-resident loops; MEASURED_PEAKS.json has no fp32 entry); HBM: " + hbm_src,
                "alg_flops_per_launch": L["flops"], "alg_bytes_per_launch": L["bytes"],
                "launch_ms": op_ms[dom], "hbm_achieved_gbs": L["bytes"] / t_meas / 1e9, "hbm_peak_gbs": hbm_peak,
                "traffic": traffic, "fp32_peaks_tflops": fp32_peaks,
                "step_weighted_frac": sum(max(Lr["flops"] / (fp32_peak * 1e12), Lr["bytes"] / (hbm_peak * 1e9))
                                          for (i, _, _) in ops for Lr in [layers[i]]) / (sum(op_ms) * 1e-3)}
    per_layer = []
    for k, (i, kind, _) in enumerate(ops):
        Lr = layers[i]
        tf, th = Lr["flops"] / (fp32_peak * 1e12), Lr["bytes"] / (hbm_peak * 1e9)
        per_layer.append({"layer": Lr["spec"].name, "op": kind, "kernel": kernel_of[kind](Lr), "ms": op_ms[k],
                          "images_per_s": N / (op_ms[k] * 1e-3), "tflops": Lr["flops"] / op_ms[k] / 1e9,
                          "roofline_frac": max(tf, th) / (op_ms[k] * 1e-3), "nnz": int(Lr["plan"].nnz)})

    # ---- cpu_baseline (N = 1 only): the reference's CPU path on this host, the reference arm's own protocol ----
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            if full_affinity:
                os.sched_setaffinity(0, full_affinity)   # the CPU baseline gets the whole host
            ips, ms, threads, kind, desc = cpu_reference_bounded(specs, args.steps, args.warmup, budget_s=60.0)
            cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": desc + ("; forward only -- the reference's CPU backward is im2col + BLAS GEMM, which this "
                                     "image cannot build" if args.train else "")}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)[:200]}

    train_launches = (nops + (sum(1 for L in layers if L["b"] is not None) + nl + (1 if world > 1 else 0)
                              if args.train else 0)) * args.steps
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": _config(label, N, world),
            "detail": {"parallelism": "batch-shard x%d, no data-path collective" % world if not args.train else
                       "batch-shard x%d, ncclAllReduce of the flat CSR-ordered gradient (%d floats) + 1/N scale through "
                       "escort_allreduce_grads" % (world, flat.numel()),
                       "epilogue": "bias+ReLU fused" if not args.train else "bias fused (training step: no ReLU)",
                       "l2": "inputs+outputs of one step (%.0f MB) exceed the 126 MB L2, so every step re-reads HBM"
                             % (sum(Lr["bytes"] for Lr in layers) / 1e6)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "copy_ceiling": copy_ceiling,
                    "how": ("pinned host (allocated on the GPU's NUMA node) -> device, C-ABI forward, device -> pinned host; "
                            "%d-image chunks over 4 streams; copy_ceiling = the same copies with no kernels" % cs)
                           if not args.train else "pinned host bottom + top_diff -> device, C-ABI forward + backward_weight + "
                           "backward_data per layer (4 streams), gradient exchange, bottom_diff + flat gradient -> pinned host"},
            "gpu_launches": train_launches, "roofline": roofline, "parity": parity, "parity_worst": parity_worst,
            "cpu_baseline": cpu, "clocks": clocks, "layers": per_layer, "train": train_leg}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    bad = max(parity_worst, train_leg["parity_worst"] if train_leg else 0.0)
    if bad > 1e-4:
        raise SystemExit("bench.py: parity FAILED: worst relative L2 %.3e > 1e-4 against the reference CPU kernels" % bad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="alexnet", choices=["alexnet", "googlenet", "resnet50", "lenet"])
    ap.add_argument("--variant", type=int, default=None, help="force a forward variant instead of autotuning")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed outputs")
    ap.add_argument("--no-train-leg", action="store_true", help="default workload: skip the short ResNet-50 training step")
    ap.add_argument("--train", action="store_true",
                    help="step = forward + masked backward (weight, data, bias) of every layer + the gradient all-reduce "
                         "(BASELINE.json configs[3]); default: forward only")
    ap.add_argument("--tune-cache", default=None, help="JSON file caching the autotuned (variant, layout) per layer")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    specs, label = _workload(args.workload, args.train)
    if args.impl == "reference":
        run_reference(args, specs, label)
    else:
        run_ours(args, specs, label)


if __name__ == "__main__":
    main()
