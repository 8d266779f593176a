#!/usr/bin/env python
"""bench.py -- residue-frames/sec of the TIMED 3D-CNN inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

Workload (BASELINE.json configs[1]): TIMED 20-class stand-in graph (the
reference ships no network, predict.py:121 loads an opaque .h5), synthetic 21^3 x 6 frames,
batch 4096 per GPU.  A *step* is one forward of one batch: every kernel from the raw frame
tensor to the (batch, 20) softmax probabilities (input conversion, six fused
conv+bias+ELU+BatchNorm implicit GEMMs, two max-pools, global average pool, softmax).

One JSON line on stdout (rank 0):
  value      frames/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        frames/s through Model.predict() (the C-ABI host call): pinned HOST frames in,
             H2D + forward + D2H of the probabilities inside the timed region
  roofline   dominant kernel (the conv op with the most FLOPs): algorithmic TFLOP/s from
             per-op CUDA events recorded inside the timed region vs MEASURED_PEAKS.json
  cpu_baseline  the oracle's torch-CPU restatement of the same graph ("port": TensorFlow is
             not installable here) on a bounded sample, all host threads
`--impl reference` times that CPU restatement alone (the reference's own path is CPU Keras).
Under torchrun each rank runs its own shard (frames are independent: no data-path
collective) and one NCCL all-gather per step reassembles the probability block.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "residue-frames/sec (TIMED stand-in, 21^3x6 frames, 20 classes)"
UNIT = "frames/s"
UNIQUE_FRAMES = 1024         # distinct synthetic frames generated per rank, tiled to the batch
CPU_BATCH = 32               # Keras predict() default batch size [EXTERNAL]


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML is polled every
    ~5 ms (the timed region of the default run is a fraction of a second: nvidia-smi, ~0.15 s per call, would
    return one or two samples); nvidia-smi is the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []            # (sm_mhz, max_mhz, power_w, [flags in NAMES order])
        self._halt = threading.Event()
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative index -> NVML handle through the PCI bus id torch reports
            import torch
            bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(gpu_index), "pci_bus_id") else None
            self._h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self._h = h
                        break
            if self._h is None:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        flags = [bool(r & n.nvmlClocksEventReasonHwSlowdown), bool(r & n.nvmlClocksEventReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksEventReasonSwThermalSlowdown), bool(r & n.nvmlClocksEventReasonSwPowerCap)]
        self.samples.append((float(sm), float(mx), pw, flags))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
        parts = [x.strip() for x in out.strip().split(",")]
        if len(parts) >= 7:
            self.samples.append((float(parts[0]), float(parts[1]), float(parts[2]),
                                 [p.lower().startswith("active") for p in parts[3:7]]))

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.005 if self._nvml is not None else 0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[3][i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.samples[0][1],
                "power_w_max": max(s[2] for s in self.samples), "reasons": reasons,
                "samples": len(self.samples), "source": self.source}


def cpu_port_frames_per_s(cfg, weights, frames, threads):
    """Time the oracle's torch-CPU restatement (oneDNN conv3d) on `frames`, batch CPU_BATCH."""
    from oracle import keras_oracle as ko
    ko.forward_torch(cfg, weights, frames[:CPU_BATCH], threads=threads)        # warm-up
    t0 = time.perf_counter()
    for i in range(0, len(frames), CPU_BATCH):
        ko.forward_torch(cfg, weights, frames[i:i + CPU_BATCH], threads=threads)
    return len(frames) / (time.perf_counter() - t0)


def run_reference(args, rank, world):
    """--impl reference: the reference path is CPU Keras (predict.py:142); TensorFlow cannot be
    installed here, so this times the oracle's CPU port of the same graph on the host cores."""
    if rank != 0:
        return
    from timed_design_b200 import standins
    cfg, weights = standins.timed_standin(20)
    threads = os.cpu_count() or 1
    per_step = 4 * CPU_BATCH
    frames = standins.synthetic_frames(per_step, seed=1234)
    from oracle import keras_oracle as ko
    for _ in range(max(args.warmup, 1)):
        ko.forward_torch(cfg, weights, frames[:CPU_BATCH], threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in range(0, per_step, CPU_BATCH):
            ko.forward_torch(cfg, weights, frames[i:i + CPU_BATCH], threads=threads)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    sample = f"{per_step} frames/step in batches of {CPU_BATCH} (bounded sample of the batch-{args.batch} workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "TIMED 20-class stand-in, 21^3x6 synthetic frames", "batch_per_step": per_step,
                   "note": "CPU port of the reference path (torch-CPU fp32, oneDNN); TensorFlow 2.13 is not installable offline"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--model", default="timed", choices=["timed", "densecpd"],
                    help="side measurements only: the bench line the driver reads is the default (timed, 20 classes)")
    ap.add_argument("--e2e-chunk", type=int, default=1024)
    ap.add_argument("--precise", action="store_true", help="side measurement: Model(precise=True)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from timed_design_b200 import standins
    from timed_design_b200.model import Model

    torch.cuda.set_device(local_rank)
    # fixture generation (stand-in calibration, synthetic frames) is CPU work: keep N ranks from
    # oversubscribing the host cores
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    cfg, weights = standins.timed_standin(args.classes) if args.model == "timed" else standins.densecpd_standin(args.classes)
    model = Model(cfg, weights, device=local_rank, max_chunk_frames=args.e2e_chunk, precise=args.precise)
    B = args.batch
    # frames are indexed globally so the data does not depend on the rank count
    uniq = standins.synthetic_frames(UNIQUE_FRAMES, seed=1234, first_index=rank * B)
    reps = -(-B // UNIQUE_FRAMES)
    frames = torch.from_numpy(uniq).to(dev).repeat(reps, 1, 1, 1, 1)[:B].contiguous()
    probs = torch.empty((B, model.n_classes), dtype=torch.float32, device=dev)
    ws = torch.empty(model.workspace_bytes(B), dtype=torch.uint8, device=dev)
    gathered = torch.empty((world * B, model.n_classes), dtype=torch.float32, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        model.forward_device(frames, probs, ws, stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, probs)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    model.set_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    op_times, n_fw = model.read_op_times()
    model.set_timing(False)
    if rank == 0:
        sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max / 1e3)

    # ---------------- e2e: the host call a user makes (Model.predict) with pinned host frames
    e2e = None
    if not args.no_e2e:
        h_frames = torch.from_numpy(uniq).repeat(reps, 1, 1, 1, 1)[:B].contiguous().pin_memory()
        x = h_frames.numpy()
        model.predict(x)                                   # warm-up (allocates staging)
        e2e_steps = max(2, min(args.steps, 5))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = model.predict(x)                         # synchronous: returns host probabilities
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * e2e_steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(x.nbytes), "d2h_bytes_per_step": int(out.nbytes),
               "steps": e2e_steps, "chunk_frames": args.e2e_chunk, "host_dtype": "float32 (pinned)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (most FLOPs) from the live per-op events
    peaks = load_peaks()
    convs = [o for o in op_times if o["flops_per_frame"] > 0]
    top = max(convs, key=lambda o: o["flops_per_frame"])
    top_ms = top["ms"] / max(n_fw, 1)
    achieved = top["flops_per_frame"] * B / (top_ms / 1e3) / 1e12
    peak = peaks["tflops_sustained"] or peaks["tflops_burst"]
    step_ms_ops = sum(o["ms"] for o in op_times) / max(n_fw, 1)
    traffic = None
    tp = ROOT / "profiles" / "r1i_dominant_kernel_ncu.json"
    if tp.exists() and args.model == "timed" and args.classes == 20 and B == 4096:
        traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "kernel": f"{model.op_kernel(top['index'], B)}[{top['name']}]",
        "peak_source": peaks["source"] + ", bf16 sustained (kernel timed inside a long step)",
        "mma_passes": 3,
        "issued_frac": 3 * achieved / peak,
        "note": "achieved counts ALGORITHMIC FLOPs; each K-step issues 3 bf16 tcgen05.mma (hi*hi+lo*hi+hi*lo split for the 1e-4 parity contract), so tensor-pipe work is 3x",
        "kernel_ms": top_ms, "kernel_share_of_step": top_ms / step_ms_ops,
        "whole_graph": {"achieved": model.flops_per_frame * B / (ms_max / args.steps / 1e3) / 1e12,
                        "frac": model.flops_per_frame * B / (ms_max / args.steps / 1e3) / 1e12 / peak},
        "per_op_ms": {f"{o['index']}:{o['name']}": round(o["ms"] / max(n_fw, 1), 4) for o in op_times},
        "per_op_kernel": {f"{o['index']}:{o['name']}": model.op_kernel(o["index"], B) for o in op_times},
    }

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_cpu = 16 * CPU_BATCH
        v = cpu_port_frames_per_s(cfg, weights, uniq[:n_cpu], threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_cpu} of the same synthetic frames, batches of {CPU_BATCH}, torch-CPU fp32 restatement "
                         f"of the graph (oracle/keras_oracle.py); TensorFlow 2.13 not installable offline"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 split operands (hi/lo planes), fp32 accumulate/epilogue",
        "data": "synthetic",
        "config": {"workload": f"{'TIMED' if args.model == 'timed' else 'DenseCPD'} {args.classes}-class stand-in inference, 21^3x6 synthetic frames",
                   "batch_per_gpu": B, "global_batch": world * B, "flops_per_frame": model.flops_per_frame,
                   "parallelism": f"frames sharded over {world} GPU(s), all-gather of probabilities" if world > 1
                   else "single GPU",
                   "l2": f"inputs larger than L2: {frames.numel() * 4 / 1e6:.0f} MB of frames per step "
                         f"({UNIQUE_FRAMES} distinct frames tiled to the batch)"},
        "e2e": e2e, "gpu_launches": model.launches_per_forward * args.steps,
        "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
