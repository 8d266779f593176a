"""Sampler benchmark (BASELINE.json configs[4]): 59 chains, lengths ~U(60,400) (seed 5), rows =
fp16-rounded softmax of N(0,2^2) logits, 10 000 sequences per chain, temperature sweep.

    python tools/bench_sampler.py [--classes 20|338] [--samples 10000] [--temps 20]

Reports sampled residues/s and sequences/s for the device pipeline (temperature -> cumsum -> draw,
letters left in HBM), the same through the host API (sample_block: probabilities up, letters down),
and the reference numpy loop (oracle, bounded sample) on the host cores."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--classes", type=int, default=20)
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--temps", type=int, default=20)
    args = ap.parse_args()
    import ctypes as C
    import torch
    from oracle import sampler_oracle as so
    from timed_design_b200 import _lib, sampling_utils as su
    lib = _lib.load()
    rng = np.random.default_rng(5)
    lengths = rng.integers(60, 401, size=59)
    chains = []
    for n in lengths:
        z = rng.standard_normal((n, args.classes)) * 2
        e = np.exp(z - z.max(1, keepdims=True))
        chains.append((e / e.sum(1, keepdims=True)).astype(np.float16).astype(np.float64))
    temps = [round(0.1 * (i + 1), 1) for i in range(args.temps)]
    letters = torch.from_numpy(np.frombuffer(("ACDEFGHIKLMNPQRSTVWY" * 17)[:args.classes].encode(), np.uint8).copy()).cuda()
    d_probs = [torch.from_numpy(p).cuda() for p in chains]
    d_tmp = [torch.empty_like(p) for p in d_probs]
    d_cdf = [torch.empty_like(p) for p in d_probs]
    d_seq = [torch.empty((args.samples * p.shape[0] + 3) // 4 * 4, dtype=torch.uint8, device="cuda") for p in d_probs]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())

    def sweep():
        for ti, t in enumerate(temps):
            for ci, p in enumerate(d_probs):
                n, c = p.shape
                src = p
                if t != 1:
                    _lib.check(lib.timed_b200_apply_temperature(P(p), n, c, float(t), P(d_tmp[ci]), st))
                    src = d_tmp[ci]
                _lib.check(lib.timed_b200_cumsum_rows(P(src), n, c, P(d_cdf[ci]), st))
                _lib.check(lib.timed_b200_sample(P(d_cdf[ci]), n, c, args.samples, 0, 42, ci * 1000 + ti, None,
                                                 P(letters), P(d_seq[ci]), None, st))
    # batched: the chains concatenated, three launches per temperature (temperature, cumsum, sample_chains)
    lens = np.array([p.shape[0] for p in chains], dtype=np.int64)
    row_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    blocks = (lens * args.samples + 3) // 4 * 4
    seq_off = np.concatenate([[0], np.cumsum(blocks)]).astype(np.int64)
    d_all = torch.from_numpy(np.concatenate(chains, axis=0)).cuda()
    d_all_tmp, d_all_cdf = torch.empty_like(d_all), torch.empty_like(d_all)
    d_row, d_off = torch.from_numpy(row_off).cuda(), torch.from_numpy(seq_off).cuda()
    d_all_seq = torch.empty(int(seq_off[-1]), dtype=torch.uint8, device="cuda")

    def sweep_batched():
        n, c = d_all.shape
        for ti, t in enumerate(temps):
            src = d_all
            if t != 1:
                _lib.check(lib.timed_b200_apply_temperature(P(d_all), n, c, float(t), P(d_all_tmp), st))
                src = d_all_tmp
            _lib.check(lib.timed_b200_cumsum_rows(P(src), n, c, P(d_all_cdf), st))
            _lib.check(lib.timed_b200_sample_chains(P(d_all_cdf), P(d_row), P(d_off), len(chains), int(seq_off[-1]), c,
                                                    args.samples, 0, 42, ti * 1000, P(letters), P(d_all_seq), st))

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    ms_per_chain = timed(sweep)
    ms = timed(sweep_batched)
    residues = int(lengths.sum()) * args.samples * len(temps)
    seqs = 59 * args.samples * len(temps)
    out = {"classes": args.classes, "chains": 59, "residues_per_chain_total": int(lengths.sum()),
           "samples_per_chain": args.samples, "temperatures": len(temps),
           "device": {"ms": ms, "residues_per_s": residues / ms * 1e3, "sequences_per_s": seqs / ms * 1e3,
                      "bytes_written_GBps": residues / ms * 1e3 / 1e9, "launches": 3 * len(temps),
                      "path": "timed_b200_sample_chains: all chains per launch"},
           "device_per_chain_launches": {"ms": ms_per_chain, "residues_per_s": residues / ms_per_chain * 1e3,
                                         "launches": 3 * 59 * len(temps)}}
    # host API: one temperature, all chains, letters copied back
    t0 = time.perf_counter()
    for ci, p in enumerate(chains):
        su.sample_block(p, args.samples, None if args.classes == 20 else ["A"] * args.classes, seed=42, stream_id=ci,
                        temperature=0.5)
    dt = time.perf_counter() - t0
    out["host_api_one_temperature"] = {"s": dt, "residues_per_s": int(lengths.sum()) * args.samples / dt}
    # reference numpy loop, bounded sample: 3 chains x 200 samples
    t0 = time.perf_counter()
    nres = 0
    for p in chains[:3]:
        so.sample_loop_numpy(p, 200, None if args.classes == 20 else ["A"] * args.classes)
        nres += p.shape[0] * 200
    dt = time.perf_counter() - t0
    out["cpu_reference_loop"] = {"sample": "3 chains x 200 samples, 1 core, metrics call removed",
                                 "residues_per_s": nres / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
