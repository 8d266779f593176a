"""``predict.py`` drop-in: same command-line flags, same importable
``load_dataset_and_predict(...)`` signature and 6-tuple, same output files as
/root/reference/predict.py -- with the Keras load/predict pair (predict.py:121,142) replaced by
the B200 graph executor and the per-batch text round trip (predict.py:145-163) kept only as an
output format, not as the data path.

Files written into ``path_to_output``: ``{model}.csv`` (float16-cast
probabilities, ``%.18e``), ``encoded_labels.csv``, ``datasetmap.txt``, ``{model}.txt``,
``{model}.fasta``, ``dataset.fasta``; rotamer mode adds ``{model}_rot.csv`` (raw float32
338-wide rows) and turns ``{model}.csv`` into residue one-hots; NMR consensus files go to the
CWD like the reference's.

Deviations, all additive: the rotamer dump is named ``{model}_rot.csv`` (the
reference's missing f-prefix produces a literal ``{model_name}_rot.csv``); ``--predict_rotamers``
/ ``--is_structure_nmr`` also accept an explicit ``True``/``False`` as the README writes them;
``--yes`` creates a missing output directory without the interactive prompt.
"""
from __future__ import annotations

import argparse
from math import ceil
from pathlib import Path

import numpy as np

from .frames import create_flat_dataset_map, load_batch, load_batch_device
from .model import load_model
from .postprocess import (convert_dataset_map_for_srb, extract_sequence_from_pred_matrix,
                          get_pdb_keys_to_filter, get_rotamer_codec, rotamer_class_to_residue,
                          save_consensus_probs, save_dict_to_fasta, save_outputs_to_file, savetxt_e18,
                          standard_amino_acids)


def _dist_context():
    """(rank, world, local_rank) under torchrun (one process per GPU); (0, 1, 0) otherwise.  The process
    group is NCCL unless TIMED_B200_DIST_BACKEND says otherwise (the CPU tests use gloo)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if not dist.is_initialized():
        backend = os.environ.get("TIMED_B200_DIST_BACKEND", "nccl")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def _gather(arr: np.ndarray, n_total: int, local_rank: int) -> np.ndarray:
    """All-gather the per-rank row blocks (contiguous frame ranges, dist.shard_range) into the full matrix."""
    import torch
    import torch.distributed as dist
    from .dist import gather_rows
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == "nccl":
        t = t.cuda(local_rank)
    return gather_rows(t, n_total).cpu().numpy()


def _structure_files(path):
    """[structure files] when `path` is a PDB file (.pdb / .pdb1 / .ent, optionally .gz) or a directory of them, else None."""
    path = Path(path)
    is_pdb = lambda f: any(x in f.name.lower() for x in (".pdb", ".ent")) and f.is_file()
    if path.is_dir():
        files = sorted(f for f in path.iterdir() if is_pdb(f))
        return files or None
    return [path] if is_pdb(path) and path.suffix.lower() not in (".hdf5", ".h5") else None


class _StructureSource:
    """Frames of a set of structure files, voxelised on the GPU on demand in blocks of ~`block_frames` frames (a directory of
    thousands of structures must not sit in HBM as float32 frames all at once); rows() hands out device-resident slices."""

    def __init__(self, files, codec, all_states, device, filter_pdb_list, block_frames: int = 16384):
        from . import voxelise
        files = [f for f in files if f.name.split(".pdb")[0][:4] not in filter_pdb_list]
        if not files:
            raise ValueError("no structure left to voxelise")
        self._vx = voxelise
        self.device = device
        self.states = voxelise.load_states(files, codec or "CNOCBCA", voxelise_all_states=all_states)
        self.flat = np.array(voxelise.flat_map_of(self.states), dtype=str)
        counts = np.array([len(st.order) for st in self.states], dtype=np.int64)
        self.first = np.concatenate([[0], np.cumsum(counts)])          # first flat row of every state
        order = list(standard_amino_acids.values())
        self.labels = np.eye(20, dtype=np.float64)[[order.index(r[3]) for r in self.flat]]
        self.block_frames = block_frames
        self._lo = self._hi = 0
        self._frames = None

    def _load(self, a: int, b: int):
        s0 = int(np.searchsorted(self.first, a, side="right") - 1)
        s1 = s0
        while s1 < len(self.states) and (self.first[s1 + 1] < b or self.first[s1 + 1] - self.first[s0] < self.block_frames):
            s1 += 1
        s1 = min(s1 + 1, len(self.states))
        self._frames = None                                            # release the previous block first
        self._frames = self._vx.voxelise_states(self.states[s0:s1], device=self.device, return_device=True)
        self._lo, self._hi = int(self.first[s0]), int(self.first[s1])

    def rows(self, a: int, b: int):
        if self._frames is None or a < self._lo or b > self._hi:
            self._load(a, b)
        return self._frames[a - self._lo:b - self._lo]


def _forward_device_rows(model, d_frames):
    """Probabilities of device-resident frames through Model.forward_device (no host round trip of the frames)."""
    import torch
    n = d_frames.shape[0]
    if tuple(d_frames.shape[1:]) != tuple(model.input_shape):
        raise ValueError(f"voxelised frames have shape {tuple(d_frames.shape[1:])} but the model expects "
                         f"{tuple(model.input_shape)} (choose the codec with --codec)")
    out = torch.empty((n, model.n_classes), dtype=torch.float32, device=d_frames.device)
    chunk = max(1, min(n, model.max_chunk_frames))
    ws = torch.empty(model.workspace_bytes(chunk), dtype=torch.uint8, device=d_frames.device)
    stream = torch.cuda.current_stream(d_frames.device).cuda_stream
    for a in range(0, n, chunk):
        model.forward_device(d_frames[a:a + chunk].contiguous(), out[a:a + chunk], ws, stream)
    return out.cpu().numpy()


def load_dataset_and_predict(
    models: list,
    dataset_path: Path,
    batch_size: int = 20,
    start_batch: int = 0,
    dataset_map_path: Path = "datasetmap.txt",
    blacklist: Path = None,
    predict_rotamers: bool = False,
    model_name_suffix: str = "",
    is_consensus: bool = False,
    path_to_output: Path = Path.cwd(),
    binary_outputs: bool = False,
    codec: str = None,
):
    """predict.py:28-194.  Returns (flat_dataset_map, pdb_to_sequence, pdb_to_probability,
    pdb_to_real_sequence, pdb_to_consensus, pdb_to_consensus_prob) of the LAST model."""
    path_to_output = Path(path_to_output)
    n_classes = 338 if predict_rotamers else 20
    print(f"Running model on {n_classes} classes. Rotamer Mode is {predict_rotamers}")
    filter_pdb_list = get_pdb_keys_to_filter(blacklist) if blacklist else []
    rank, world, local_rank = _dist_context()
    # Extension: `dataset_path` may be a structure file or a directory of them instead of an
    # aposteriori .hdf5 -- the frames are then voxelised on the GPU (voxelise.py) and never touch the disk or the host.
    structures = _structure_files(dataset_path)
    source = None
    if structures is not None:
        source = _StructureSource(structures, codec, is_consensus, local_rank, filter_pdb_list)
        flat_dataset_map = source.flat
    elif Path(dataset_map_path).exists():          # a stale map is reused, as the reference does
        flat_dataset_map = np.genfromtxt(dataset_map_path, delimiter=",", dtype="str")
        if flat_dataset_map.ndim == 1:
            flat_dataset_map = flat_dataset_map[None, :]
    else:
        flat_dataset_map, _ = create_flat_dataset_map(dataset_path, filter_pdb_list)
    old_datasetmap = len(flat_dataset_map[0]) == 4
    flat_categories = get_rotamer_codec()[1] if predict_rotamers else None
    cls_to_res = rotamer_class_to_residue() if predict_rotamers else None
    n_batches = ceil(len(flat_dataset_map) / batch_size)
    out = None
    for i, m in enumerate(models):
        model_name = (m.stem if isinstance(m, Path) else str(m)) + model_name_suffix
        frame_model = load_model(Path(m), device=local_rank)
        if frame_model.n_classes != n_classes:
            raise ValueError(f"{m}: model has {frame_model.n_classes} outputs but "
                             f"{'--predict_rotamers' if predict_rotamers else 'residue mode'} expects {n_classes}")
        def predicted(ranges):
            """Yields (probabilities, one-hot labels) of the frame ranges [a, b) of the flat map, in order.  On the .hdf5
            route consecutive ranges are read and predicted together in groups of >= 2048 frames -- one index call, one copy
            of the stored chunks, one inflate launch, one forward for many of the CLI's (small: default 12) batches; the
            probabilities of a frame do not depend on how frames are grouped (tests/test_properties_gpu.py) -- and the NEXT
            group is read while the GPU predicts the current one."""
            ranges = list(ranges)
            if source is not None:
                for a, b in ranges:
                    yield _forward_device_rows(frame_model, source.rows(a, b)), source.labels[a:b]
                return
            from concurrent.futures import ThreadPoolExecutor
            group_frames = max(int(batch_size), 2048)
            groups = []                                   # lists of consecutive (a, b) ranges
            for a, b in ranges:
                if groups and groups[-1][-1][1] == a and b - groups[-1][0][0] <= group_frames:
                    groups[-1].append((a, b))
                else:
                    groups.append([(a, b)])

            def load(grp):
                """Frames of a group: stored chunks inflated on the device when the file allows it
                (frames.load_batch_device), else load_batch on the host threads."""
                rows = flat_dataset_map[grp[0][0]:grp[-1][1]]
                dev = load_batch_device(dataset_path, rows, local_rank)
                return (True, *dev) if dev is not None else (False, *load_batch(dataset_path, rows))

            with ThreadPoolExecutor(max_workers=1) as pool:
                nxt = pool.submit(load, groups[0]) if groups else None
                for k, grp in enumerate(groups):
                    on_device, X_grp, y_grp = nxt.result()
                    nxt = pool.submit(load, groups[k + 1]) if k + 1 < len(groups) else None
                    p_grp = _forward_device_rows(frame_model, X_grp) if on_device else frame_model.predict(X_grp)
                    a0 = grp[0][0]
                    for a, b in grp:
                        yield p_grp[a - a0:b - a0], y_grp[a - a0:b - a0]

        rot_out = path_to_output / f"{model_name}_rot.csv"
        model_out = rot_out if predict_rotamers else path_to_output / f"{model_name}.csv"
        rows_before = sum(1 for _ in open(model_out)) if model_out.exists() else 0
        raw_rows = []
        gathered = None
        if world > 1:
            # frames shard by flat index into contiguous per-rank ranges (each chain's rows stay adjacent); every rank
            # predicts its range, ONE all-gather reassembles probabilities and labels, rank 0 writes the reference's files
            from .dist import shard_range
            # resume (predict.py:32,54-57,126): only the frames from batch `start_batch` on are predicted; they are the
            # range that is sharded, and the gathered rows are indexed relative to its first frame
            first = min(start_batch * batch_size, len(flat_dataset_map))
            n_total = len(flat_dataset_map) - first
            lo, hi = shard_range(n_total, rank, world)
            lo, hi = lo + first, hi + first
            preds, labels = [], []
            for y_pred_b, y_true_b in predicted((b0, min(b0 + batch_size, hi)) for b0 in range(lo, hi, batch_size)):
                preds.append(y_pred_b)
                labels.append(np.asarray(y_true_b, dtype=np.float64))
            local_p = np.concatenate(preds) if preds else np.zeros((0, n_classes), np.float32)
            local_y = np.concatenate(labels) if labels else np.zeros((0, 20), np.float64)
            gathered = (_gather(local_p.astype(np.float32), n_total, local_rank), _gather(local_y, n_total, local_rank))
        local = None if gathered is not None else predicted(
            (index * batch_size, min((index + 1) * batch_size, len(flat_dataset_map))) for index in range(start_batch, n_batches))
        pend_pred, pend_true, pend_rows = [], [], 0
        for index in range(start_batch, n_batches):
            if gathered is not None:
                r0 = (index - start_batch) * batch_size
                y_pred_batch = gathered[0][r0:r0 + batch_size]
                y_true_batch = gathered[1][r0:r0 + batch_size]
            else:
                y_pred_batch, y_true_batch = next(local)
            raw_rows.append(y_pred_batch)
            if rank != 0:
                continue                               # only rank 0 touches the output directory
            # The files are appended row by row, so writing several of the CLI's batches at once leaves the same bytes; the
            # rows are flushed every >= 2048 frames (whole batches: the files always end on a batch boundary, which is what
            # --start_batch resumes from) instead of once per 12-frame batch.
            pend_pred.append(np.asarray(y_pred_batch))
            pend_true.append(np.asarray(y_true_batch))
            pend_rows += len(pend_pred[-1])
            if pend_rows >= 2048 or index == n_batches - 1:
                yp, yt = np.concatenate(pend_pred), np.concatenate(pend_true)
                pend_pred, pend_true, pend_rows = [], [], 0
                if predict_rotamers:
                    with open(rot_out, "a") as f:
                        savetxt_e18(f, yp)
                    yp = np.eye(20, dtype=int)[cls_to_res[np.argmax(yp, axis=1)]]
                save_outputs_to_file(list(yt), {i: list(yp)}, flat_dataset_map, i, model_name, path_to_output)
        frame_model.close()
        flat_dataset_map = np.array(flat_dataset_map)
        if rank == 0:
            convert_dataset_map_for_srb(flat_dataset_map, model_name, path_to_output)
        # predict.py:163 re-parses the whole CSV as float16.  The rows written above are these
        # arrays printed with 18 significant digits (already float16-cast in residue mode), so
        # casting them to float16 reproduces the parsed matrix bit for bit without the text round
        # trip -- unless the file already held rows (append mode / start_batch): then honour it.
        if rows_before == 0 and start_batch == 0 and raw_rows:
            prediction_matrix = np.concatenate(raw_rows).astype(np.float16)
        else:
            if world > 1:                     # rank 0 has appended the resumed rows: every rank parses the complete file
                import torch.distributed as dist
                dist.barrier()
            prediction_matrix = np.genfromtxt(model_out, delimiter=",", dtype=np.float16)
        if prediction_matrix.ndim == 1:
            prediction_matrix = prediction_matrix[None, :]
        if binary_outputs and rank == 0:     # the %.18e text costs ~25 bytes per probability; .npy is 2
            np.save(path_to_output / f"{model_name}.npy", prediction_matrix)
            if predict_rotamers and raw_rows:
                np.save(path_to_output / f"{model_name}_rot.npy", np.concatenate(raw_rows).astype(np.float32))
        out = extract_sequence_from_pred_matrix(
            flat_dataset_map, prediction_matrix,
            rotamers_categories=flat_categories if predict_rotamers else None,
            old_datasetmap=old_datasetmap, is_consensus=False)
        if is_consensus:       # NMR ensembles: running float16 mean over the states + argmax, on the device
            from .device_post import nmr_consensus
            cons, cons_prob = nmr_consensus(out[1], flat_categories if predict_rotamers else None)
            out = (out[0], out[1], out[2], cons, cons_prob)
        if rank != 0:
            continue
        save_dict_to_fasta(out[0], model_name, path_to_output)
        save_dict_to_fasta(out[2], "dataset", path_to_output)
        if out[3]:
            save_dict_to_fasta(out[3], model_name + "_consensus")        # CWD, as the reference
            save_consensus_probs(out[4], model_name, path_to_output)
    return (flat_dataset_map, *out)


def _flag(value):
    """store_true-compatible flag that also accepts the README's explicit ``True``/``False``."""
    if value is None or value is True:
        return True
    v = str(value).strip().lower()
    if v in ("true", "1", "yes", "y"):
        return True
    if v in ("false", "0", "no", "n"):
        return False
    raise argparse.ArgumentTypeError(f"expected True or False, got {value!r}")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Predict with TIMED (B200-native path)")
    p.add_argument("--batch_size", type=int, default=12, help="Frames predicted per batch (default: 12)")
    p.add_argument("--path_to_dataset", type=str, help="Frame dataset (.hdf5)")
    p.add_argument("--path_to_datasetmap", default="datasetmap.txt", type=str, help="Dataset map (.txt)")
    p.add_argument("--path_to_model", type=str, help="Keras model file (.h5) or .npz container")
    p.add_argument("--path_to_blacklist", type=str, default=None, help="Directory of PDB-code lists to exclude")
    p.add_argument("--path_to_output", type=str, default=".", help="Output directory (default: CWD)")
    p.add_argument("--output_analysis", action="store_true", help="Accepted for compatibility; unused")
    p.add_argument("--predict_rotamers", nargs="?", const=True, default=False, type=_flag,
                   help="Model predicts 338 rotamer classes instead of 20 residues")
    p.add_argument("--is_structure_nmr", nargs="?", const=True, default=False, type=_flag,
                   help="NMR ensemble: also build a consensus over the states")
    p.add_argument("--yes", action="store_true", help="Create a missing output directory without asking")
    p.add_argument("--codec", type=str, default=None,
                   help="Atom codec when --path_to_dataset is a structure file / directory voxelised on the GPU "
                        "(CNOCBCA, CNOCBCAQ, CNOCBCAP, ...; default CNOCBCA)")
    p.add_argument("--binary_outputs", action="store_true",
                   help="Also write {model}.npy (the float16 matrix of {model}.csv; sample.py reads it directly)")
    return p


def main(args) -> None:
    """predict.py:197-247."""
    args.path_to_dataset = Path(args.path_to_dataset)
    args.path_to_model = Path(args.path_to_model)
    args.path_to_datasetmap = Path(args.path_to_datasetmap)
    args.path_to_output = Path(args.path_to_output)
    if not args.path_to_output.exists():
        if not getattr(args, "yes", False):
            print(f"Output directory at {args.path_to_output} does not exist. Do you want to create it? (y/n)")
            if input() != "y":
                print("Exiting...")
                raise SystemExit(0)
        args.path_to_output.mkdir(parents=True, exist_ok=True)
    if args.path_to_blacklist:
        args.path_to_blacklist = Path(args.path_to_blacklist)
        assert args.path_to_blacklist.exists(), f"Path to blacklist at {args.path_to_blacklist} does not exists."
    assert args.path_to_model.exists(), f"Path to model at {args.path_to_model} does not exists."
    assert args.path_to_dataset.exists(), f"Path to dataset at {args.path_to_dataset} does not exists."
    assert args.batch_size > 0, f"Batch size must be higher than 0 but got {args.batch_size}"
    load_dataset_and_predict(
        [args.path_to_model], args.path_to_dataset, batch_size=args.batch_size, start_batch=0,
        blacklist=args.path_to_blacklist, dataset_map_path=args.path_to_datasetmap,
        predict_rotamers=args.predict_rotamers, is_consensus=args.is_structure_nmr,
        path_to_output=args.path_to_output, binary_outputs=getattr(args, "binary_outputs", False),
        codec=getattr(args, "codec", None))


def cli(argv=None) -> None:
    main(build_parser().parse_args(argv))


if __name__ == "__main__":
    cli()
