#!/usr/bin/env python
"""Benchmark of the sparse-convolution hot path: MinkUNet34C forward+backward(+SGD) over
synthetic 100k-voxel clouds, batch 8 per rank (BASELINE.json configs[3]), reporting
active-voxels/s.  Contract: see the one JSON line printed by rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = fresh SparseTensor (all coordinate maps and kernel maps rebuilt, as in the
reference's examples/multigpu_ddp.py:103-106) -> forward -> cross-entropy -> backward
(DDP gradient all-reduce when N>1) -> SGD step.
  value : inputs already resident in HBM when the timed region starts
  e2e   : the same step driven through the public API from PINNED HOST buffers — H2D copy of
          coordinates/features/labels and a D2H read of the loss inside the timed region
--impl reference times the reference's own CPU implementation (oracle/_ref, compiled from
the reference sources by oracle/build_ref.py) on the host cores, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "active-voxels/sec MinkUNet34C fwd+bwd @100k pts, 1/2/4/8 B200 vs CPU ref"
UNIT = "voxels/s"
MODEL = "MinkUNet34C"
VOXELS_PER_CLOUD = 100_000
CLOUDS_PER_RANK = 8
# algorithmic conv FLOPs per voxel, fwd+bwd (601.6 GFLOP / 100k voxels)
FLOP_PER_VOXEL_FWD_BWD = 601.6e9 / 100_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default=os.environ.get("MEB200_BENCH_DTYPE", "bf16"),
                    choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--clouds", type=int, default=CLOUDS_PER_RANK)
    ap.add_argument("--voxels", type=int, default=VOXELS_PER_CLOUD)
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
def make_batch(n_clouds, n_voxels, seed0):
    """Synthetic 'surface' clouds (SURVEY.md §8d generator), one batch index per cloud."""
    import torch
    from examples.synthetic import surface_cloud
    coords = torch.cat([surface_cloud(n_voxels, seed0 + j, batch=j) for j in range(n_clouds)], 0)
    g = torch.Generator().manual_seed(seed0)
    feats = torch.rand(len(coords), 3, generator=g)
    labels = torch.zeros(len(coords), dtype=torch.int64)
    return coords.contiguous(), feats, labels


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback"}


class NvmlClockSampler:
    """SM clock + clock-event reasons read through NVML (nvidia_ml_py) every 20 ms from a
    thread, so that even a half-second timed region gets tens of samples.  `start()` returns
    False when NVML is not usable; the caller then falls back to the nvidia-smi sampler."""
    # nvmlClocksEventReason* bit masks (nvml.h)
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index=0, period_s=0.02):
        self.index, self.period, self.sm, self.mask, self.mx = index, period_s, [], 0, None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:   # CUDA_VISIBLE_DEVICES may renumber devices: match on the PCI address
                import torch
                pr = torch.cuda.get_device_properties(self.index)
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._sample()
        except Exception:
            return False
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()
        return True

    def _sample(self):
        self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
        get = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            self.nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.mask |= int(get(self.h))

    def _loop(self):
        while not self._stop.wait(self.period):
            try:
                self._sample()
            except Exception:
                return

    def stop(self):
        self._stop.set()
        self.th.join(timeout=2)
        sm = sorted(self.sm[1:] or self.sm)      # the first sample predates the timed region
        reasons = sorted(nm for bit, nm in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(sm), "source": "nvml"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import ref
    from examples.minkunet import minkunet
    cores = os.cpu_count() or 1
    threads = min(cores, 16)  # the reference caps itself at 16 (MinkowskiEngine/__init__.py:35-46)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    torch.set_num_threads(threads)
    REF = ref.import_reference()
    total = args.steps + args.warmup
    # bounded sample: ~16k voxels/s on 8 cores -> keep the whole run within a few minutes
    n_vox = int(max(10_000, min(args.voxels, 150 * 16_000 / max(total, 1))))
    torch.manual_seed(0)
    net = minkunet(args.model, REF, 3, 20, 3)
    opt = torch.optim.SGD(net.parameters(), lr=1e-2)
    crit = torch.nn.CrossEntropyLoss()
    coords, feats, labels = make_batch(1, n_vox, seed0=0)

    def step():
        opt.zero_grad()
        x = REF.SparseTensor(feats, coords)
        out = net(x)
        loss = crit(out.F, labels)
        loss.backward()
        opt.step()
        return float(loss)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = n_vox * args.steps / dt
    sample = f"1 cloud x {n_vox} voxels per step, {args.steps} steps after {args.warmup} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} fwd+bwd+SGD, synthetic surface clouds, maps rebuilt "
                               f"every step (BASELINE configs[3]); reference CPU path",
                   "voxels_per_step": n_vox},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": sample, "host_cores": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    """Bounded CPU sample of the same workload through the compiled reference."""
    import torch
    from oracle import ref
    from examples.minkunet import minkunet
    cores = os.cpu_count() or 1
    threads = min(cores, 16)
    torch.set_num_threads(threads)
    try:
        REF = ref.import_reference()
    except Exception as e:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"unavailable: {e}"}
    n_vox = min(args.voxels, 100_000)
    torch.manual_seed(0)
    net = minkunet(args.model, REF, 3, 20, 3)
    crit = torch.nn.CrossEntropyLoss()
    coords, feats, labels = make_batch(1, n_vox, seed0=0)
    times = []
    for it in range(3):
        t0 = time.perf_counter()
        x = REF.SparseTensor(feats, coords)
        loss = crit(net(x).F, labels)
        loss.backward()
        times.append(time.perf_counter() - t0)
    best = min(times[1:])
    return {"value": n_vox / best, "unit": UNIT, "cores": threads, "kind": "reference",
            "host_cores": cores,
            "sample": f"1 cloud x {n_vox} voxels, fwd+bwd, min of 2 after 1 warm-up "
                      f"({best:.2f} s/step)"}


# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import minkowskiengine_b200 as ME
    from examples.minkunet import minkunet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]

    torch.manual_seed(0)
    net = minkunet(args.model, ME, 3, 20, 3).to(dev)
    if world > 1:
        net = ME.MinkowskiSyncBatchNorm.convert_sync_batchnorm(net)
        # Default = the reference example's plain DDP (examples/multigpu_ddp.py), the setting the
        # committed 2-GPU number was measured with.  MEB200_DDP_TUNE=1: SyncBN keeps the running
        # statistics identical on every rank, so the per-step buffer broadcast is redundant, and
        # gradients can live in the all-reduce buckets (no extra copy) - to be measured.
        ddp_kw = {}
        if os.environ.get("MEB200_DDP_TUNE", "0") not in ("", "0"):
            ddp_kw = dict(broadcast_buffers=False, gradient_as_bucket_view=True)
        net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local_rank], **ddp_kw)
    opt = torch.optim.SGD(net.parameters(), lr=1e-2)
    crit = torch.nn.CrossEntropyLoss()

    coords_h, feats_h, labels_h = make_batch(args.clouds, args.voxels, seed0=rank * args.clouds)
    coords_h, feats_h, labels_h = coords_h.pin_memory(), feats_h.pin_memory(), labels_h.pin_memory()
    n_vox = coords_h.shape[0]
    coords_d, feats_d, labels_d = coords_h.to(dev), feats_h.to(dev).to(dtype), labels_h.to(dev)

    def step_resident():
        opt.zero_grad(set_to_none=True)
        x = ME.SparseTensor(feats_d, coords_d)
        out = net(x)
        loss = crit(out.F.float(), labels_d)
        loss.backward()
        opt.step()
        return loss

    def step_e2e():
        c = coords_h.to(dev, non_blocking=True)
        f = feats_h.to(dev, non_blocking=True).to(dtype)
        l = labels_h.to(dev, non_blocking=True)
        opt.zero_grad(set_to_none=True)
        x = ME.SparseTensor(f, c)
        out = net(x)
        loss = crit(out.F.float(), l)
        loss.backward()
        opt.step()
        return float(loss.item())           # D2H read of the step's result

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ME._lib.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ME._lib.launch_count() - l0

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = None
    if rank == 0:
        sampler = NvmlClockSampler(local_rank)
        if not sampler.start():
            sampler = ClockSampler(local_rank)
            sampler.start()
    ms, launches = timed(step_resident, args.steps)
    clocks = None
    if rank == 0:
        try:
            clocks = sampler.stop()
        except Exception as exc:                      # never lose the measurement to the sampler
            clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"sampler failed: {exc}"]}
    step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)

    total_vox = torch.tensor([float(n_vox)], device=dev)
    if world > 1:
        dist.all_reduce(total_vox)
    total_vox = float(total_vox.item())
    value = total_vox * args.steps / (ms * 1e-3)
    e2e_value = total_vox * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel family, measured live --------------------------
    # Two EXTRA steps after the timed region with every convolution launch bracketed by CUDA
    # events on the launch stream (torch's current stream, the one libmeb200 launches on).
    # achieved = algorithmic FLOPs (2*P*Cin*Cout per launch) / summed durations.
    roof = None
    # every rank runs the extra steps (the backward contains DDP's gradient all-reduce)
    prof = ME.backend.profile_conv_kernels(step_resident, steps=2)
    if rank == 0:
        peaks = load_peaks()
        dom = prof["conv_fwd_dgrad"]
        wg = prof["conv_wgrad"]
        step_ms = ms / args.steps

        def tf(d):
            return d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0

        def gb(d):
            return d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0

        peak = peaks["bf16_tflops_sustained"]   # kernel timed inside a long step
        roof = {"bound": "tensor", "achieved": tf(dom), "peak": peak, "unit": "TFLOP/s",
                "frac": tf(dom) / peak, "traffic": None,
                "peak_source": f"{peaks['source']} (MEASURED_PEAKS.json bf16_tflops_sustained)",
                "kernel": "k_conv_rs (tcgen05 sparse-conv forward/dgrad)",
                "launches_per_step": dom["launches"] / dom["steps"],
                "ms_per_step": dom["ms"] / dom["steps"],
                "share_of_step": dom["ms"] / dom["steps"] / step_ms,
                "flops_per_step": dom["flops"] / dom["steps"],
                "hbm": {"achieved_gbs": gb(dom), "peak_gbs": peaks["hbm_gbs"],
                        "frac": gb(dom) / peaks["hbm_gbs"],
                        "note": "compulsory bytes (SURVEY.md 8d) / same durations"},
                "wgrad": {"kernel": "k_wgrad_tc (tcgen05 wgrad)", "achieved": tf(wg),
                          "frac": tf(wg) / peak, "ms_per_step": wg["ms"] / wg["steps"],
                          "share_of_step": wg["ms"] / wg["steps"] / step_ms,
                          "launches_per_step": wg["launches"] / wg["steps"]}}
        # DRAM traffic comes from an ncu --set full capture of ONE launch (the largest layer),
        # not from this run: copy the committed figure and say which launch it describes.
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))["k_conv_rs"]
            roof["traffic"] = cap["dram_bytes_read"] + cap["dram_bytes_write"]
            roof["traffic_launch"] = cap["launch"]
            roof["traffic_algorithmic_bytes"] = cap["algorithmic_bytes"]
            roof["traffic_source"] = cap["capture"]
        except (OSError, KeyError, ValueError):
            pass

    if rank == 0:
        cpu = None if (args.no_cpu_baseline or world > 1) else cpu_baseline(args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{args.model} fwd+bwd+SGD on {args.clouds} synthetic surface "
                                   f"clouds x {args.voxels} voxels per rank (BASELINE configs[3]), "
                                   "coordinate/kernel maps rebuilt every step",
                       "global_batch_clouds": args.clouds * world, "voxels_per_step": total_vox,
                       "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations + neighbour tables, several GB) "
                             "exceeds the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(coords_h.nbytes + feats_h.nbytes + labels_h.nbytes) * 1,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
