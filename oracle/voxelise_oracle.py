"""ORACLE (test infrastructure, never shipped on the product path): numpy restatement of the residue-frame voxeliser.

What it restates: aposteriori 2.4.0's ``make-frame-dataset`` as the reference uses it (/root/reference/README.md:84-97
command line and :242 prose; /root/reference/ui.py:63-87 keyword arguments; dataset schema
/root/reference/design_utils/utils.py:238-251).  aposteriori is an un-vendored dependency (requirements.txt) that is
absent offline, so this is a restatement of its PUBLISHED behaviour, not of its source: **parity unpinned** for the voxel
values.  One constant of the reference pins the frame convention: with C-alpha at the origin, N on the +y axis and C in
the xy plane at x > 0, the mean C-beta of 1ubq's residues is (-0.771, -0.527, -1.216), 0.03 A from the hard-coded
C-beta (-0.741287356, -0.53937931, -1.224287356) the README describes; every other axis assignment is >= 0.29 A away
(tests/test_voxelise_cpu.py).

Arithmetic (the same in csrc/voxelise.cuh): float32 atom coordinates and frame matrices, transformed in float64; voxel
index = round-half-even(local / edge) + V // 2; atoms whose voxel is outside the grid are dropped; gaussian frames add a
3x3x3 stamp exp(-|n - d|^2 / (2 sigma^2)) (d = sub-voxel offset, sigma in voxels) normalised over its 27 cells, clipped
at the grid border, accumulated as round(w * 2^24) integers; boolean frames set the atom's voxel.
"""
from __future__ import annotations

import numpy as np

FIXED = float(1 << 24)


def voxelise(atoms_xyzs: np.ndarray, atom_channel: np.ndarray, atom_residue: np.ndarray, atom_is_cb: np.ndarray,
             res_frame: np.ndarray, res_property, residues, voxels_per_side: int, voxel_edge: float, n_channels: int,
             as_gaussian: bool, encode_cb: bool, ideal_cb_xyz_sigma, cb_channel: int, property_channel: int) -> np.ndarray:
    """-> (len(residues), V, V, V, C) float32 (gaussian) or uint8 (boolean)."""
    V, C = int(voxels_per_side), int(n_channels)
    half = V // 2
    inv_edge = np.float64(np.float32(1.0) / np.float32(voxel_edge))
    out = np.zeros((len(residues), V, V, V, C), dtype=np.int64)
    xyz = atoms_xyzs[:, :3].astype(np.float64)

    def add(frame, local, sigma, ch, prop, with_prop):
        u = local * inv_edge
        idx = np.rint(u).astype(np.int64) + half
        if (idx < 0).any() or (idx >= V).any():
            return
        if not as_gaussian:
            frame[idx[0], idx[1], idx[2], ch] = int(FIXED)
            if with_prop:
                frame[idx[0], idx[1], idx[2], property_channel] += int(np.rint(np.float64(np.float32(prop)) * FIXED))
            return
        d = u - (idx - half)
        inv2s2 = 1.0 / (2.0 * float(np.float32(sigma)) * float(np.float32(sigma)))
        w = [np.exp(-((np.arange(3) - 1) - d[k]) ** 2 * inv2s2) for k in range(3)]
        norm = 1.0 / (w[0].sum() * w[1].sum() * w[2].sum())
        for a in range(3):
            for b in range(3):
                for c in range(3):
                    x, y, z = idx[0] + a - 1, idx[1] + b - 1, idx[2] + c - 1
                    if min(x, y, z) < 0 or max(x, y, z) >= V:
                        continue
                    wt = w[0][a] * w[1][b] * w[2][c] * norm
                    frame[x, y, z, ch] += int(np.rint(wt * FIXED))
                    if with_prop:
                        frame[x, y, z, property_channel] += int(np.rint(wt * float(np.float32(prop)) * FIXED))

    has_prop = property_channel >= 0 and res_property is not None
    for k, r in enumerate(residues):
        f = res_frame[r].astype(np.float64)
        origin, R = f[:3], f[3:].reshape(3, 3)
        t = xyz - origin
        # the kernel rounds every product and sum separately (no fused multiply-add)
        local = np.stack([(R[i, 0] * t[:, 0] + R[i, 1] * t[:, 1]) + R[i, 2] * t[:, 2] for i in range(3)], axis=1)
        # cheap reject of far atoms (exact test repeated inside add)
        near = np.abs(local * inv_edge).max(axis=1) < half + 1.0
        for a in np.nonzero(near)[0]:
            ch = int(atom_channel[a])
            if ch < 0:
                continue
            cb = bool(atom_is_cb[a])
            if encode_cb and cb and int(atom_residue[a]) == r:
                continue
            with_prop = has_prop and cb
            add(out[k], local[a], atoms_xyzs[a, 3], ch, res_property[atom_residue[a]] if with_prop else 0.0, with_prop)
        if encode_cb:
            cbx = np.asarray(ideal_cb_xyz_sigma, dtype=np.float32)
            add(out[k], cbx[:3].astype(np.float64), cbx[3], cb_channel, res_property[r] if has_prop else 0.0, has_prop)
    if as_gaussian:
        return (out.astype(np.float32) * np.float32(1.0 / FIXED)).astype(np.float32)
    return (out != 0).astype(np.uint8)
