"""ORACLE (test infrastructure, never shipped on the product path).

Literal, loop-by-loop restatement of the reference's prediction post-processing so that the
product's vectorised versions can be checked row for row:
  * get_rotamer_codec                  /root/reference/design_utils/utils.py:410-465
  * compress_rotamer_predictions_to_20 /root/reference/design_utils/utils.py:468-484
  * extract_sequence_from_pred_matrix  /root/reference/design_utils/utils.py:616-723
  * convert_dataset_map_for_srb        /root/reference/design_utils/utils.py:533-566
  * save_outputs_to_file               /root/reference/design_utils/utils.py:726-771

Pinned against tests/golden/{rotamer_codec.json,postprocess.*,files.json}, produced by running
the reference's own functions (tests/golden/make_golden.py).  The two ampal tables the
reference imports (absent here) are restated below and pinned by the rotamer offsets the
reference quotes at utils.py:425.
"""
from __future__ import annotations

import io
from itertools import product

import numpy as np

STANDARD_AMINO_ACIDS = {
    "A": "ALA", "C": "CYS", "D": "ASP", "E": "GLU", "F": "PHE", "G": "GLY", "H": "HIS", "I": "ILE",
    "K": "LYS", "L": "LEU", "M": "MET", "N": "ASN", "P": "PRO", "Q": "GLN", "R": "ARG", "S": "SER",
    "T": "THR", "V": "VAL", "W": "TRP", "Y": "TYR"}
N_CHI = {"ARG": 4, "ASN": 2, "ASP": 2, "CYS": 1, "GLN": 3, "GLU": 3, "HIS": 2, "ILE": 2, "LEU": 2,
         "LYS": 4, "MET": 3, "PHE": 2, "PRO": 2, "SER": 1, "THR": 1, "TRP": 2, "TYR": 2, "VAL": 1}


def get_rotamer_codec():
    """utils.py:432-461: residues in standard order, 3**n_chi rotamers each, labels RES_chi."""
    flat, rot_to_20, guide = [], {}, []
    r_count = 0
    for i, (_, res) in enumerate(STANDARD_AMINO_ACIDS.items()):
        guide.append(r_count)
        if res in N_CHI:
            rots = list(product([1, 2, 3], repeat=N_CHI[res]))
            for r, rota in enumerate(rots):
                flat.append(f"{res}_{''.join(str(x) for x in rota)}")
                v = np.array([0] * 20)
                v[i] = 1
                rot_to_20[r_count + r] = v
            r_count += len(rots)
        else:
            flat.append(f"{res}_0")
            v = np.array([0] * 20)
            v[i] = 1
            rot_to_20[r_count] = v
            r_count += 1
    return rot_to_20, flat, guide


def compress_rotamer_predictions_to_20(pm):
    _, _, guide = get_rotamer_codec()
    return np.add.reduceat(pm, guide, axis=1)


def extract_sequence_from_pred_matrix(flat_dataset_map, prediction_matrix, rotamers_categories=None,
                                      is_consensus=False):
    """utils.py:639-723 (loops kept literal)."""
    res_to_r = {v: k for k, v in STANDARD_AMINO_ACIDS.items()}
    if rotamers_categories:
        if len(rotamers_categories[0]) == 1:
            res_dic = rotamers_categories
        else:
            res_dic = [res_to_r[r.split("_")[0]] for r in rotamers_categories]
    else:
        res_dic = list(STANDARD_AMINO_ACIDS.keys())
    max_idx = np.argmax(prediction_matrix, axis=1)
    seqs, probs, real = {}, {}, {}
    previous = 0
    old = len(flat_dataset_map[0]) == 4
    for i in range(len(flat_dataset_map)):
        if old:
            pdb_chain, chain, _, res = flat_dataset_map[i]
            count = 1
        else:
            pdb_chain, count = flat_dataset_map[i]
            count = int(count)
            chain = ""
        pdb_chain = pdb_chain + chain
        if pdb_chain not in seqs:
            seqs[pdb_chain] = ""
            real[pdb_chain] = ""
            probs[pdb_chain] = []
        for n in range(previous, previous + count):
            idx = i if old else n
            probs[pdb_chain].append(list(prediction_matrix[idx]))
            seqs[pdb_chain] += res_dic[max_idx[idx]]
            if old:
                real[pdb_chain] += res_to_r[res]
        if not old:
            previous += count
    if not is_consensus:
        return seqs, probs, real, None, None
    cons_prob, cons = {}, {}
    last = ""
    for pdb_chain in seqs:
        cur = pdb_chain.split("_")[0]
        if last != cur:
            cons_prob[cur] = np.array(probs[pdb_chain])
            last = cur
        else:
            cons_prob[cur] = (cons_prob[cur] + np.array(probs[pdb_chain])) / 2   # running pairwise mean
    for k, v in cons_prob.items():
        cons[k] = "".join(res_dic[m] for m in np.argmax(v, axis=1))
    return seqs, probs, real, cons, cons_prob


def srb_map_text(flat_dataset_map) -> str:
    """utils.py:549-566 -> the text of {model}.txt."""
    counts = {}
    for pdb, chain, _, _ in flat_dataset_map:
        if "_0" in pdb:
            pdb = pdb.split("_0")[0]
        if len(pdb) == 4:
            pdb += chain
        counts[pdb] = counts.get(pdb, 0) + 1
    return "ignore_uncommon False\ninclude_pdbs\n##########\n" + "".join(f"{k} {v}\n" for k, v in counts.items())


def predictions_csv_text(y_pred) -> str:
    """utils.py:768-771: float16 cast, then np.savetxt default '%.18e', comma separated."""
    buf = io.StringIO()
    np.savetxt(buf, np.array(y_pred, dtype=np.float16), delimiter=",")
    return buf.getvalue()


def labels_csv_text(y_true) -> str:
    buf = io.StringIO()
    np.savetxt(buf, np.asarray(y_true), delimiter=",", fmt="%i")
    return buf.getvalue()


def fasta_text(d) -> str:
    """utils.py:610-613."""
    return "".join(f">{k}\n{v}\n" for k, v in d.items())
