#!/usr/bin/env python
"""bench.py -- throughput of timed-design's inference / sampling hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config timed20|timed338|densecpd|sampler] [--batch B] [--frames F]

Default (what the driver reads): BASELINE.json configs[1] -- TIMED 20-class stand-in graph (the
reference ships no network, predict.py:121 loads an opaque .h5), synthetic 21^3 x 6 frames, batch 4096 per GPU.
A *step* is one forward of one batch: every kernel from the raw frame tensor to the (batch, classes) softmax
probabilities.  `--config` selects the other BASELINE configs with their own metric / workload labels:
  timed338  configs[2]  TIMED-rotamer 338-class head, same frames
  densecpd  configs[3]  DenseCPD dense-block stand-in (a step of 4096 frames runs as macro-chunks sized to the workspace)
  sampler   configs[4]  sample.py Monte-Carlo: 59 chains x 10 000 sequences x 20 temperatures per step

One JSON line on stdout (rank 0):
  value      units/s (frames or sampled residues), inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        the same through the host API a user calls (Model.predict / sampling_utils.sample_chains): pinned HOST
             inputs, H2D + kernels + D2H of the results inside the timed region
  roofline   dominant kernel: algorithmic FLOP/s (or bytes/s) from CUDA events recorded inside the timed region vs
             MEASURED_PEAKS.json
  cpu_baseline  the oracle's CPU restatement of the same workload ("port": TensorFlow is not installable here) on a
             bounded sample, all host threads
`--impl reference` times that CPU restatement alone (the reference's own path is CPU Keras / a numpy loop).
Under torchrun each rank runs its own shard (frames / sample blocks are independent: no data-path collective) and one
NCCL all-gather per step reassembles the results.  `--frames F` sets steps = ceil(F / (batch * ranks)): a long run
that shows the sustained clock.
"""
from __future__ import annotations

import argparse
import json
import re
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

UNIT = "frames/s"
UNIQUE_FRAMES = 1024         # distinct synthetic frames generated per rank, tiled to the batch
CPU_BATCH = 32               # Keras predict() default batch size [EXTERNAL]
WS_BUDGET = 64 << 30         # macro-chunking: largest device workspace one forward may take

CONFIGS = {
    "timed20": {"metric": "residue-frames/sec (TIMED stand-in, 21^3x6 frames, 20 classes)",
                "workload": "TIMED 20-class stand-in inference, 21^3x6 synthetic frames", "baseline_config": 1},
    "timed338": {"metric": "residue-frames/sec (TIMED-rotamer stand-in, 21^3x6 frames, 338 classes)",
                 "workload": "TIMED-rotamer 338-class stand-in inference, 21^3x6 synthetic frames", "baseline_config": 2},
    "densecpd": {"metric": "residue-frames/sec (DenseCPD stand-in, 21^3x6 frames, 20 classes)",
                 "workload": "DenseCPD dense-block stand-in inference (3 blocks x 6 layers, growth 32), 21^3x6 synthetic frames",
                 "baseline_config": 3},
    "sampler": {"metric": "sampled residues/sec (sample.py Monte-Carlo, 59 chains x 10k sequences x temperatures 0.1-2.0)",
                "workload": "temperature -> cumsum -> categorical draw for 59 chains (lengths ~U(60,400)), 10 000 sequences per "
                            "chain, 20 temperatures", "baseline_config": 4},
}


def build_graph(config: str):
    from timed_design_b200 import standins
    if config == "timed20":
        return standins.timed_standin(20)
    if config == "timed338":
        return standins.timed_standin(338, seed=8)
    if config == "densecpd":
        return standins.densecpd_standin(20)
    raise ValueError(config)


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


# ----------------------------------------------------------------------------- sampler workload (BASELINE configs[4])
def sampler_chains(classes: int):
    """59 chains (scripts/README.md:38), lengths ~U(60,400) seed 5, rows = fp16-rounded softmax of N(0,2^2) logits
    (SURVEY.md 8(d) config 5)."""
    rng = np.random.default_rng(5)
    lengths = rng.integers(60, 401, size=59)
    chains = []
    for n in lengths:
        z = rng.standard_normal((n, classes)) * 2
        e = np.exp(z - z.max(1, keepdims=True))
        chains.append((e / e.sum(1, keepdims=True)).astype(np.float16).astype(np.float64))
    return chains


SAMPLER_TEMPS = [round(0.1 * (i + 1), 1) for i in range(20)]


def _cpu_sampler_task(a):
    from oracle import sampler_oracle as so
    probs, n, t = a
    if t != 1:
        probs = so.apply_temp_to_probs(probs, t)
    np.random.seed()                      # forked workers would otherwise share the parent's generator state
    cats = None if probs.shape[1] == 20 else ["A"] * probs.shape[1]
    return sum(len(x) for x in so.sample_loop_numpy(probs, n, cats))


def cpu_sampler_residues_per_s(chains, samples_per_chain: int, workers: int, temperature: float = 0.5):
    """The reference's fan-out (sampling_utils.py:181-190): Pool(workers).starmap over chains, each running the verbatim
    numpy loop (cumsum recomputed per sample; the ampal metrics call removed)."""
    import multiprocessing as mp
    tasks = [(p, samples_per_chain, temperature) for p in chains]
    with mp.get_context("fork").Pool(workers) as pool:
        pool.map(_cpu_sampler_task, tasks[:workers])                      # warm-up (fork + imports)
        t0 = time.perf_counter()
        n = sum(pool.map(_cpu_sampler_task, tasks))
        dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference path is CPU Keras (predict.py:142) / a numpy loop under a process pool
    (sampling_utils.py:93-197); TensorFlow cannot be installed here, so the conv arm times the oracle's CPU port of the
    same graph; the sampler arm times the reference's loop verbatim (it is pure numpy)."""
    if rank != 0:
        return
    C = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    if args.config == "sampler":
        chains = sampler_chains(20)
        workers = min(8, threads)
        per_chain = 200
        cpu_sampler_residues_per_s(chains[:workers], 20, workers)          # warm-up
        t0 = time.perf_counter()
        n = 0
        for _ in range(args.steps):
            v, dt = cpu_sampler_residues_per_s(chains, per_chain, workers)
            n += v * dt
        dt = time.perf_counter() - t0
        v = n / dt
        sample = f"{per_chain} sequences per chain x 59 chains at T=0.5 per step (of 10 000 x 20 temperatures), Pool({workers})"
        print(json.dumps({
            "impl": "reference", "metric": C["metric"], "value": v, "unit": "residues/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": C["workload"], "note": "the reference's numpy loop verbatim (oracle/sampler_oracle.py), ampal metrics call removed"},
            "cpu_baseline": {"value": v, "unit": "residues/s", "cores": workers, "kind": "reference", "sample": sample},
            "e2e": {"value": v, "unit": "residues/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }), flush=True)
        return
    from timed_design_b200 import standins
    cfg, weights = build_graph(args.config)
    per_step = 4 * CPU_BATCH if args.config != "densecpd" else CPU_BATCH
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
        "impl": "reference", "metric": C["metric"], "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": C["workload"], "batch_per_step": per_step,
                   "note": "CPU port of the reference path (torch-CPU fp32, oneDNN); TensorFlow 2.13 is not installable offline"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_sampler(args, rank, world, local_rank):
    """configs[4]: one step = the whole temperature sweep for all chains (60 launches: temperature, cumsum,
    sample_chains per temperature), letters left in HBM.  Under N ranks every rank draws its own block of the sample
    index (first_sample = rank * samples: weak scaling, N x 10 000 sequences per chain in total)."""
    import ctypes as Ct
    import torch
    import torch.distributed as dist
    from timed_design_b200 import _lib, sampling_utils as su
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()
    C = CONFIGS["sampler"]
    classes = args.classes
    samples = args.samples
    chains = sampler_chains(classes)
    lens = np.array([p.shape[0] for p in chains], dtype=np.int64)
    row_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    blocks = (lens * samples + 3) // 4 * 4
    seq_off = np.concatenate([[0], np.cumsum(blocks)]).astype(np.int64)
    letters = torch.from_numpy(np.frombuffer(("ACDEFGHIKLMNPQRSTVWY" * 17)[:classes].encode(), np.uint8).copy()).to(dev)
    d_all = torch.from_numpy(np.concatenate(chains, axis=0)).to(dev)
    d_tmp, d_cdf = torch.empty_like(d_all), torch.empty_like(d_all)
    d_row, d_off = torch.from_numpy(row_off).to(dev), torch.from_numpy(seq_off).to(dev)
    d_seq = torch.empty(int(seq_off[-1]), dtype=torch.uint8, device=dev)
    st = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: Ct.c_void_p(t.data_ptr())
    n_rows, n_cls = d_all.shape
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    draw_ms = []

    def sweep(record=False):
        for ti, t in enumerate(SAMPLER_TEMPS):
            src = d_all
            if t != 1:
                _lib.check(lib.timed_b200_apply_temperature(P(d_all), n_rows, n_cls, float(t), P(d_tmp), st))
                src = d_tmp
            _lib.check(lib.timed_b200_cumsum_rows(P(src), n_rows, n_cls, P(d_cdf), st))
            if record and ti == len(SAMPLER_TEMPS) // 2:
                ev[2].record()
            _lib.check(lib.timed_b200_sample_chains(P(d_cdf), P(d_row), P(d_off), len(chains), int(seq_off[-1]), n_cls,
                                                    samples, rank * samples, 42, ti * 1000, P(letters), P(d_seq), st))
            if record and ti == len(SAMPLER_TEMPS) // 2:
                ev[3].record()

    for _ in range(max(args.warmup, 3)):
        sweep()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev[0].record()
    for k in range(args.steps):
        sweep(record=True)
    ev[1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    draw = ev[2].elapsed_time(ev[3])            # the draw launch of the middle temperature, last step
    if rank == 0:
        sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    res_per_sweep = int(lens.sum()) * samples * len(SAMPLER_TEMPS)
    value = world * res_per_sweep * args.steps / (ms_max / 1e3)

    # e2e: the host API (sampling_utils.sample_chains): probabilities from host memory, letters back to the host
    e2e = None
    if not args.no_e2e:
        cats = None if classes == 20 else ["A"] * classes
        su.sample_chains(chains, samples, cats, seed=42, temperature=0.5)        # warm-up
        torch.cuda.synchronize()
        e2e_temps = SAMPLER_TEMPS[:4]
        t0 = time.perf_counter()
        nbytes = 0
        for tt in e2e_temps:
            seqs = su.sample_chains(chains, samples, cats, seed=42, first_sample=rank * samples, temperature=tt)
            nbytes += sum(x.nbytes for x in seqs)
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        e2e = {"value": world * int(lens.sum()) * samples * len(e2e_temps) / float(td.item()), "unit": "residues/s",
               "h2d_bytes_per_step": int(sum(p.nbytes for p in chains)) * len(SAMPLER_TEMPS),
               "d2h_bytes_per_step": int(nbytes // len(e2e_temps)) * len(SAMPLER_TEMPS),
               "sample": f"{len(e2e_temps)} of the 20 temperatures through sampling_utils.sample_chains "
                         "(host float64 probabilities in, host uint8 letters out)"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    res_per_draw = int(lens.sum()) * samples
    achieved = res_per_draw / (draw / 1e3) / 1e9             # GB/s written: 1 byte per sampled residue
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
        "traffic": None, "kernel": "sample_tiled_kernel", "kernel_ms": draw,
        "kernel_share_of_step": draw * len(SAMPLER_TEMPS) / (ms_max / args.steps),
        "peak_source": peaks["source"] + ", STREAM-style copy",
        "algorithmic_bytes": "1 B written per sampled residue; the (rows, classes) fp64 CDF (2-36 MB) is L2-resident and "
                             "uniforms are generated in registers (Philox4x32-10), so reads do not reach HBM",
        "note": "far below the HBM roofline by construction (1 B per residue): ncu (profiles/r2h_sampler_tiled_ncu.json) shows "
                "the tiled kernel bound by instruction issue -- issue slots 83 % busy, ALU pipe 57 % -- one Philox4x32-10 "
                "block per two residues plus a lower-bound search of the CDF row staged in shared memory",
    }
    cpu = None
    if not args.no_cpu_baseline:
        workers = min(8, os.cpu_count() or 1)
        v, dt = cpu_sampler_residues_per_s(sampler_chains(classes), 200, workers)
        cpu = {"value": v, "unit": "residues/s", "cores": workers, "kind": "reference",
               "sample": f"200 sequences per chain x 59 chains at T=0.5 ({dt:.1f} s), the reference's numpy loop verbatim under "
                         f"Pool({workers}) (sampling_utils.py:181-190), ampal metrics call removed"}
    print(json.dumps({
        "metric": C["metric"], "value": value, "unit": "residues/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (probabilities, CDF, uniforms), u8 letters", "data": "synthetic",
        "config": {"workload": C["workload"], "classes": classes, "samples_per_chain_per_gpu": samples,
                   "residues_per_step_per_gpu": res_per_sweep, "sequences_per_s": value / (int(lens.sum()) / 59.0),
                   "parallelism": f"sample index sharded over {world} GPU(s) (counter-based RNG: union == one device's draw)",
                   "l2": f"outputs larger than L2: {int(seq_off[-1]) / 1e6:.0f} MB of letters written per temperature"},
        "e2e": e2e, "gpu_launches": 3 * len(SAMPLER_TEMPS) * args.steps - args.steps,   # T = 1 skips the temperature launch
        "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
    }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="timed20", choices=sorted(CONFIGS),
                    help="BASELINE.json config; the line the driver reads is the default (timed20)")
    ap.add_argument("--batch", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--frames", type=int, default=0, help="long run: steps = ceil(frames / (batch * ranks))")
    ap.add_argument("--classes", type=int, default=20, help="sampler config only: 20 or 338 categories")
    ap.add_argument("--samples", type=int, default=10000, help="sampler config only: sequences per chain")
    ap.add_argument("--e2e-chunk", type=int, default=1024)
    ap.add_argument("--fast-accum", action="store_true",
                    help="A/B: corrections of the bf16 split into the main accumulator (Model(precise=False))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.frames > 0:
        args.steps = max(1, -(-args.frames // (args.batch * world)))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.config == "sampler":
        run_sampler(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    from timed_design_b200 import standins
    from timed_design_b200.model import Model

    C = CONFIGS[args.config]
    torch.cuda.set_device(local_rank)
    # fixture generation (stand-in calibration, synthetic frames) is CPU work: keep N ranks from
    # oversubscribing the host cores
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    cfg, weights = build_graph(args.config)
    model = Model(cfg, weights, device=local_rank, max_chunk_frames=args.e2e_chunk, precise=not args.fast_accum)
    B = args.batch
    # macro-chunks: a step of B frames runs as forwards of `chunk` frames, the largest power-of-two split whose
    # workspace fits WS_BUDGET (DenseCPD keeps 256-channel tensors at 21^3: ~40 MB of activations per frame)
    chunk = B
    while chunk > 64 and model.workspace_bytes(chunk) > WS_BUDGET:
        chunk //= 2
    # frames are indexed globally so the data does not depend on the rank count
    uniq = standins.synthetic_frames(UNIQUE_FRAMES, seed=1234, first_index=rank * B)
    reps = -(-B // UNIQUE_FRAMES)
    frames = torch.from_numpy(uniq).to(dev).repeat(reps, 1, 1, 1, 1)[:B].contiguous()
    probs = torch.empty((B, model.n_classes), dtype=torch.float32, device=dev)
    ws = torch.empty(model.workspace_bytes(chunk), dtype=torch.uint8, device=dev)
    gathered = torch.empty((world * B, model.n_classes), dtype=torch.float32, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        for f0 in range(0, B, chunk):
            model.forward_device(frames[f0:f0 + chunk], probs[f0:f0 + chunk], ws, stream)
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
        # the float32 host path is PCIe-bound (222 KB per frame); the same call with the frame dtypes that carry fewer bytes:
        # float16 (a caller whose frames are exact in half) and uint8 (boolean voxels, voxels_as_gaussian = False datasets)
        alt = {}
        for name, conv in (("float16", lambda t: t.to(torch.float16)), ("uint8", lambda t: (t > 0.2).to(torch.uint8))):
            hx = conv(h_frames).contiguous().pin_memory().numpy()
            model.predict(hx)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                model.predict(hx)
            ta = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ta, op=dist.ReduceOp.MAX)
            alt[name] = {"value": world * B * e2e_steps / float(ta.item()), "unit": UNIT, "h2d_bytes_per_step": int(hx.nbytes)}
        e2e["by_host_dtype"] = alt

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (most FLOPs) from the live per-op events
    peaks = load_peaks()
    fw_per_step = -(-B // chunk)
    n_steps_timed = max(n_fw / fw_per_step, 1e-9)           # op times are summed over the recorded forwards
    convs = [o for o in op_times if o["flops_per_frame"] > 0]
    top = max(convs, key=lambda o: o["flops_per_frame"])
    top_ms = top["ms"] / n_steps_timed
    achieved = top["flops_per_frame"] * B / (top_ms / 1e3) / 1e12
    peak = peaks["tflops_sustained"] or peaks["tflops_burst"]
    step_ms_ops = sum(o["ms"] for o in op_times) / n_steps_timed
    traffic, traffic_source = None, None
    tp = ROOT / "profiles" / "r2p_dominant_kernel_ncu.json"
    if tp.exists() and args.config == "timed20" and B == 4096:
        prof = json.loads(tp.read_text())
        traffic = prof.get("dram_bytes_per_launch")
        traffic_source = (f"from profile ({tp.relative_to(ROOT)}: {prof.get('source', 'ncu capture of this kernel at this batch')}), "
                          f"not measured in this run")
    top_kernel = model.op_kernel(top["index"], chunk)
    mv = re.search(r"valid taps ([0-9.]+)", top_kernel)
    valid_taps = float(mv.group(1)) if mv else 1.0
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_source,
        "kernel": f"{top_kernel}[{top['name']}]",
        "peak_source": peaks["source"] + ", bf16 sustained (kernel timed inside a long step)",
        "mma_passes": 3,
        "valid_tap_fraction": valid_taps,
        "issued_frac": 3 * valid_taps * achieved / peak,
        "note": "achieved counts ALGORITHMIC FLOPs (the dense conv, zero padding included, as the reference's Keras layer "
                "does them); each K-step issues 3 bf16 tcgen05.mma (hi*hi+lo*hi+hi*lo split for the 1e-4 parity contract) and "
                "voxel-stationary tiles skip the taps that fall into the zero padding, so tensor-pipe work is "
                "3 x valid_tap_fraction of the algorithmic FLOPs (issued_frac)",
        "kernel_ms": top_ms, "kernel_share_of_step": top_ms / step_ms_ops,
        "whole_graph": {"achieved": model.flops_per_frame * B / (ms_max / args.steps / 1e3) / 1e12,
                        "frac": model.flops_per_frame * B / (ms_max / args.steps / 1e3) / 1e12 / peak},
        "per_op_ms": {f"{o['index']}:{o['name']}": round(o["ms"] / n_steps_timed, 4) for o in op_times},
        "per_op_kernel": {f"{o['index']}:{o['name']}": model.op_kernel(o["index"], chunk) for o in op_times},
    }
    if "col2im over kw in the epilogue" in top_kernel:
        # DenseCPD's growth convs (N = 3 x 32 columns): the tensor roofline is kept as the yardstick north_star names, but
        # what paces the kernel is the SM's shared memory (profiles/r2c_densecpd_ncu_kernels.csv)
        roofline["paced_by"] = ("shared-memory bandwidth: per 126-row tile 1.2 MB of tcgen05 operand reads (N <= 192 MMAs read "
                                "17 KB per K = 16 step for 144 cycles of math) + 1.0 MB of TMA fills against 128 B/clk/SM; "
                                "ncu: tensor pipe 53 %, L2 53 %, xbar->SM 40 %, DRAM bytes = algorithmic")

    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_cpu = (16 if args.config != "densecpd" else 4) * CPU_BATCH
        v = cpu_port_frames_per_s(cfg, weights, uniq[:n_cpu], threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_cpu} of the same synthetic frames, batches of {CPU_BATCH}, torch-CPU fp32 restatement "
                         f"of the graph (oracle/keras_oracle.py); TensorFlow 2.13 not installable offline"}

    line = {
        "metric": C["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 split operands (hi/lo planes), fp32 accumulate/epilogue",
        "data": "synthetic",
        "config": {"workload": C["workload"], "baseline_config": C["baseline_config"],
                   "batch_per_gpu": B, "global_batch": world * B, "forward_chunk_frames": chunk,
                   "flops_per_frame": model.flops_per_frame,
                   "accumulation": "corrections in the main accumulator (A/B)" if args.fast_accum else "separate correction accumulator (default)",
                   "parallelism": f"frames sharded over {world} GPU(s), all-gather of probabilities" if world > 1
                   else "single GPU",
                   "l2": f"inputs larger than L2: {frames.numel() * 4 / 1e6:.0f} MB of frames per step "
                         f"({UNIQUE_FRAMES} distinct frames tiled to the batch)"},
        "e2e": e2e, "gpu_launches": model.launches_per_forward * fw_per_step * args.steps,
        "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
