"""``sample.py`` drop-in: same flags and ``main_sample(args)`` entry point as
/root/reference/sample.py, drawing the sequences on the GPU.

Flow (sample.py:19-93): read the prediction matrix (``{model}.csv``) and the dataset map,
apply the temperature (skipped when it is exactly 1, sample.py:40-41), group rows per chain,
draw ``sample_n`` sequences per chain, write ``.json`` / ``.fasta`` / ``_metrics.csv`` into the
CWD under ``{csv stem}_temp_{T}_n_{N}_{first chain}``.

Deviations: ``--seed`` is honoured (the reference builds a generator and discards it,
sample.py:21, so its output is unseeded); ``--workers`` is accepted and ignored.
"""
from __future__ import annotations

import argparse
from pathlib import Path

import numpy as np

from . import sampling_utils
from .postprocess import (_THREE_TO_ONE, extract_sequence_from_pred_matrix, get_rotamer_codec,
                          load_datasetmap, load_matrix_csv)
from .predict import _flag


def main_sample(args):
    sampling_utils.set_seed(args.seed)
    args.path_to_pred_matrix = Path(args.path_to_pred_matrix)
    args.path_to_datasetmap = Path(args.path_to_datasetmap)
    assert args.path_to_pred_matrix.exists(), f"Prediction Matrix file {args.path_to_pred_matrix} does not exist"
    assert args.path_to_datasetmap.exists(), f"Dataset Map file {args.path_to_datasetmap} does not exist"
    if args.path_to_pred_matrix.suffix == ".npy":      # binary fast path written by predict.py --binary_outputs
        prediction_matrix = np.load(args.path_to_pred_matrix).astype(np.float64)
    else:
        prediction_matrix = load_matrix_csv(args.path_to_pred_matrix)
    if prediction_matrix.ndim == 1:
        prediction_matrix = prediction_matrix[None, :]
    datasetmap = load_datasetmap(args.path_to_datasetmap, is_old=args.support_old_datasetmap)
    if args.temperature != 1:
        prediction_matrix = sampling_utils.apply_temp_to_probs(prediction_matrix, t=args.temperature)
    if args.predict_rotamers:
        flat_categories = [_THREE_TO_ONE[res.split("_")[0]] for res in get_rotamer_codec()[1]]
    else:
        flat_categories = None
    _, pdb_to_probability, _, _, _ = extract_sequence_from_pred_matrix(
        datasetmap, prediction_matrix, rotamers_categories=flat_categories,
        old_datasetmap=args.support_old_datasetmap)
    pdb_codes = list(pdb_to_probability.keys())
    print(f"Ready to sample {args.sample_n} for each of the {len(pdb_codes)} proteins from "
          f"{args.path_to_pred_matrix.stem}.")
    pdb_to_sample = sampling_utils.sample_with_multiprocessing(
        args.workers, pdb_codes, args.sample_n, pdb_to_probability, flat_categories)
    import os
    if int(os.environ.get("RANK", "0")) != 0:          # torchrun: every rank holds the gathered samples, rank 0 writes
        return []
    return sampling_utils.save_as(
        pdb_to_sample,
        filename=f"{args.path_to_pred_matrix.stem}_temp_{args.temperature}_n_{args.sample_n}_{pdb_codes[0]}",
        mode=args.save_as)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Monte-Carlo sequence sampling from TIMED predictions (B200-native)")
    p.add_argument("--path_to_pred_matrix", type=str, help="Prediction matrix ({model}.csv, or {model}.npy from predict.py --binary_outputs)")
    p.add_argument("--path_to_datasetmap", default="datasetmap.txt", type=str, help="Dataset map (.txt)")
    p.add_argument("--predict_rotamers", nargs="?", const=True, default=False, type=_flag,
                   help="The matrix holds 338 rotamer classes instead of 20 residues")
    p.add_argument("--sample_n", type=int, default=100, help="Sequences drawn per chain")
    p.add_argument("--save_as", type=str, default="all", const="all", nargs="?", choices=["fasta", "json", "all"],
                   help="Output container(s) (default: all)")
    p.add_argument("--workers", type=int, default=8, help="Accepted for compatibility; ignored on the GPU")
    p.add_argument("--temperature", type=float, default=1, help="Softmax temperature (default 1: unchanged)")
    p.add_argument("--support_old_datasetmap", nargs="?", const=True, default=False, type=_flag,
                   help="The dataset map is the old 4-column datasetmap.txt")
    p.add_argument("--seed", type=int, default=42, help="Seed of the counter-based generator (default: 42)")
    return p


def cli(argv=None):
    return main_sample(build_parser().parse_args(argv))


if __name__ == "__main__":
    cli()
