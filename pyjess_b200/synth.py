"""Deterministic synthetic workloads of the BASELINE.json configs 3-5.

Used by bench.py and the parity tests; there is no network for real template
libraries or structure sets, so the inputs are generated:

* proteins: confined self-avoiding C-alpha walk (step 3.8 A, sphere of volume
  150 A^3 per residue), sequence iid from UniProt-like background frequencies,
  every residue a rigid copy of a real residue geometry (data/residue_library.json,
  cut from tests/data/1AMY.pdb by tools/make_residue_library.py) under a uniformly
  random rotation, coordinates rounded to 3 decimals (PDB precision);
* templates: 3-6 (cfg 5: 8) neighbouring residues of a source protein, 3 atoms per
  residue with the match modes and alternate residues of the M-CSA residue table
  (restated from vendor/jess/filter_output.py:8-31 of the reference; VAL uses CG1
  because valine has no CG), coordinates jittered by N(0, 0.3 A);
* multimers (cfg 5): C4 homo-tetramers of one chain, chains A-D.

Everything is seeded; the same (kind, index) always yields the same object.
All generation is vectorised across structures.
"""
from __future__ import annotations

import json
import os
from typing import List, Sequence, Tuple

import numpy as np

from ._model import ATOM_DTYPE, Molecule, Template, TemplateAtom, _encode_token

_HERE = os.path.dirname(os.path.abspath(__file__))

AA = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE",
      "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"]
# UniProtKB/Swiss-Prot amino-acid composition (percent), same order as AA
AA_FREQ = np.array([8.25, 5.53, 4.06, 5.45, 1.37, 3.93, 6.75, 7.07, 2.27, 5.96,
                    9.66, 5.84, 2.42, 3.86, 4.70, 6.56, 5.34, 1.08, 2.92, 6.87])
AA_FREQ = AA_FREQ / AA_FREQ.sum()

#: residue -> [(atom, match mode, alternate one-letter residues)]
MCSA_ATOMS = {
    "ALA": [("C", 0, ""), ("CA", 0, ""), ("CB", 0, "")],
    "CYS": [("CA", 0, ""), ("CB", 0, ""), ("SG", 0, "")],
    "ASP": [("CG", 3, "E"), ("OD1", 3, "E"), ("OD2", 3, "E")],
    "GLU": [("CD", 3, "D"), ("OE1", 3, "D"), ("OE2", 3, "D")],
    "PHE": [("CE1", 3, ""), ("CE2", 3, ""), ("CZ", 0, "")],
    "GLY": [("CA", 0, ""), ("C", 0, ""), ("O", 0, "")],
    "HIS": [("CG", 0, ""), ("CD2", 8, ""), ("ND1", 8, "")],
    "ILE": [("CA", 0, "VL"), ("CB", 0, "VL"), ("CG1", 3, "VL")],
    "LYS": [("CD", 0, ""), ("CE", 0, ""), ("NZ", 0, "")],
    "LEU": [("CA", 0, "VI"), ("CB", 0, "VI"), ("CG", 3, "VI")],
    "MET": [("CG", 0, ""), ("SD", 0, ""), ("CE", 0, "")],
    "ASN": [("CG", 3, "Q"), ("OD1", 1, "Q"), ("ND2", 1, "Q")],
    "PRO": [("C", 0, ""), ("CA", 0, ""), ("O", 0, "")],
    "GLN": [("CD", 3, "N"), ("OE1", 1, "N"), ("NE2", 1, "N")],
    "ARG": [("CZ", 0, ""), ("NH1", 3, ""), ("NH2", 3, "")],
    "SER": [("CA", 0, "TY"), ("CB", 0, "TY"), ("OG", 1, "TY")],
    "THR": [("CA", 0, "SY"), ("CB", 0, "SY"), ("OG1", 1, "SY")],
    "TRP": [("NE1", 0, ""), ("CZ2", 0, ""), ("CH2", 0, "")],
    "TYR": [("CE1", 3, "ST"), ("CZ", 3, "ST"), ("OH", 1, "ST")],
    "VAL": [("CA", 0, "LI"), ("CB", 0, "LI"), ("CG1", 3, "LI")],
}
_ONE = dict(zip("ACDEFGHIKLMNPQRSTVWY", ["ALA", "CYS", "ASP", "GLU", "PHE", "GLY", "HIS", "ILE", "LYS", "LEU",
                                            "MET", "ASN", "PRO", "GLN", "ARG", "SER", "THR", "VAL", "TRP", "TYR"]))

_LIB = None


def residue_library():
    """name tokens + C-alpha centred coordinates per residue type."""
    global _LIB
    if _LIB is None:
        with open(os.path.join(_HERE, "data", "residue_library.json")) as f:
            raw = json.load(f)
        lib = []
        for res in AA:
            names = raw[res]["names"]
            toks = np.array([_encode_token(b"_" + n.encode() if len(n) < 4 else n.encode(), 4) for n in names], "S4")
            lib.append({"res": res, "names": names, "tokens": toks,
                        "xyz": np.asarray(raw[res]["xyz"], dtype=np.float64),
                        "restok": _encode_token(res.encode(), 3)})
        _LIB = lib
    return _LIB


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], 1)


def _ca_walks(rng, lengths: np.ndarray, trials: int = 10, fast: bool = False) -> np.ndarray:
    """(P, Lmax, 3) confined self-avoiding walks, vectorised over the P chains.
    `fast` evaluates the self-avoidance distances through a matmul (different rounding,
    hence its own deterministic stream of structures: used for the ranged-length cfg 4)."""
    P, Lmax = len(lengths), int(lengths.max())
    radius = (3.0 * lengths * 150.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    ca = np.zeros((P, Lmax, 3), dtype=np.float32)
    sq = np.zeros((P, Lmax), dtype=np.float32)
    for i in range(1, Lmax):
        d = rng.normal(size=(P, trials, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=2, keepdims=True)
        cand = ca[:, i - 1, None, :] + np.float32(3.8) * d                      # (P, trials, 3)
        inside = np.linalg.norm(cand, axis=2) <= radius[:, None]
        if i >= 2 and fast:
            # |c - p|^2 = |c|^2 + |p|^2 - 2 c.p through one batched matmul (cfg 4's long chains)
            prev = ca[:, : i - 1, :]
            d2 = (np.einsum("ptc,ptc->pt", cand, cand)[:, :, None] + sq[:, None, : i - 1]
                  - np.float32(2.0) * np.matmul(cand, prev.transpose(0, 2, 1)))
            mind = np.sqrt(np.maximum(np.min(d2, axis=2), np.float32(0.0)))
        elif i >= 2:
            prev = ca[:, : i - 1, :]                                            # non-adjacent earlier C-alphas
            diff = cand[:, :, None, :] - prev[:, None, :, :]
            mind = np.sqrt(np.min(np.einsum("ptkc,ptkc->ptk", diff, diff), axis=2))
        else:
            mind = np.full((P, trials), 1e9, dtype=np.float32)
        score = np.where(inside, np.minimum(mind, np.float32(4.5)), np.float32(-1.0) + 0.0 * mind)
        score = score + np.where(inside, 0.0, -np.linalg.norm(cand, axis=2) * 1e-3)
        pick = np.argmax(score, axis=1)                                           # first valid, else best of trials
        ca[:, i, :] = cand[np.arange(P), pick]
        sq[:, i] = np.einsum("pc,pc->p", ca[:, i, :], ca[:, i, :])
    return ca


def _assemble(rng, ca: np.ndarray, seq: np.ndarray, chain: str = "A", resseq0: int = 1,
              serial0: int = 1) -> np.ndarray:
    """Atoms of ONE chain: ca (L,3), seq (L,) residue-type indices -> record array."""
    lib = residue_library()
    L = len(seq)
    rots = _random_rotations(rng, L)
    counts = np.array([len(lib[s]["names"]) for s in seq])
    total = int(counts.sum())
    arr = np.zeros(total, dtype=ATOM_DTYPE)
    start = np.concatenate([[0], np.cumsum(counts)])
    for t in range(20):                                  # vectorised per residue type
        where = np.nonzero(seq == t)[0]
        if len(where) == 0:
            continue
        rel = lib[t]["xyz"]                              # (k,3)
        k = len(rel)
        xyz = np.einsum("nij,kj->nki", rots[where], rel) + ca[where][:, None, :].astype(np.float64)
        idx = (start[where][:, None] + np.arange(k)[None, :]).reshape(-1)
        arr["x"][idx] = np.round(xyz.reshape(-1, 3), 3)
        arr["name"][idx] = np.tile(lib[t]["tokens"], len(where))
        arr["resName"][idx] = lib[t]["restok"]
        arr["resSeq"][idx] = np.repeat(where + resseq0, k)
    arr["serial"] = np.arange(serial0, serial0 + total)
    arr["occupancy"] = 1.0
    arr["tempFactor"] = 20.0
    arr["altLoc"] = 0x20
    arr["iCode"] = 0x20
    arr["chainID1"] = 0x20
    arr["chainID2"] = ord(chain)
    arr["segID"] = b"____"
    arr["element"] = b"__"
    return arr


def _protein_group(args):
    """Worker of the ranged-length generator: one group of chains from its own RNG stream."""
    seed, first, count, g0, lens, ids = args
    rng = np.random.default_rng([seed, first, count, g0])
    ca = _ca_walks(rng, lens, fast=True)
    out = []
    for j in range(len(lens)):
        L = int(lens[j])
        seq = rng.choice(20, size=L, p=AA_FREQ)
        out.append(_assemble(rng, ca[j, :L], seq))
    return ids, out


def protein_groups(count: int, length, seed: int, first: int = 0):
    """The independent generation units of a ranged-length set: groups of 64 chains of similar
    length, each with its own RNG stream -- so ANY subset of the groups can be generated (by any
    process) and yields the same structures.  Returns [(job, ids, lengths)]: `ids` = positions of
    the group's structures in the set of `count`, `job` = argument of `_protein_group`."""
    rng = np.random.default_rng([seed, first, count])
    lengths = rng.integers(length[0], length[1] + 1, size=count)
    order = np.argsort(lengths, kind="stable")
    group = 64 if lengths.max() > 600 else 256
    out = []
    for g0 in range(0, count, group):
        ids = order[g0:g0 + group]
        out.append(((seed, first, count, g0, lengths[ids], ids), ids, lengths[ids]))
    return out


def proteins_of_groups(groups, first: int = 0, workers=None):
    """Generate the structures of `groups` (entries of `protein_groups`) on a process pool.
    Returns (ids, molecules), group by group."""
    jobs = [g[0] for g in groups]
    if workers is None:
        workers = min(len(jobs), len(os.sched_getaffinity(0)))
    n = sum(len(g[1]) for g in groups)
    if n >= 512 and workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            results = pool.map(_protein_group, jobs, chunksize=1)
    else:
        results = list(map(_protein_group, jobs))
    ids, mols = [], []
    for gi, arrs in results:
        for i, arr in zip(gi, arrs):
            ids.append(int(i))
            mols.append(Molecule._from_records(arr, id=f"S{(first + int(i)) % 1000:03d}"))
    return np.asarray(ids, dtype=np.int64), mols


def proteins(count: int, length=300, seed: int = 1000, first: int = 0) -> List[Molecule]:
    """`count` single-chain proteins; `length` is an int or a (lo, hi) range
    (cfg 3: 300; cfg 4: (300, 1500)).  Ranged lengths are generated group by group from
    independent RNG streams (deterministic whatever the worker count), on a process pool
    when the set is large (cfg 4: 6,250 structures per GPU)."""
    ranged = isinstance(length, (tuple, list))
    if ranged:
        ids, mols = proteins_of_groups(protein_groups(count, length, seed, first), first)
        out: List[Molecule] = [None] * count
        for i, m in zip(ids.tolist(), mols):
            out[i] = m
        return out
    rng = np.random.default_rng([seed, first, count])
    lengths = np.full(count, int(length))
    out = [None] * count
    order = np.argsort(lengths, kind="stable")
    # walk chains of similar length together to bound the (P, trials, L) temporaries
    group = 64 if lengths.max() > 600 else 256
    for g0 in range(0, count, group):
        ids = order[g0:g0 + group]
        ca = _ca_walks(rng, lengths[ids])
        for j, i in enumerate(ids):
            L = int(lengths[i])
            seq = rng.choice(20, size=L, p=AA_FREQ)
            arr = _assemble(rng, ca[j, :L], seq)
            out[i] = Molecule._from_records(arr, id=f"S{(first + i) % 1000:03d}")
    return out


def multimers(count: int, chain_length: int = 500, copies: int = 4, seed: int = 3000) -> List[Molecule]:
    """cfg 5 structures: C4 homo-multimers (chains A, B, C, ...) of a generated chain,
    the copies rotated about z and pushed out to ~0.9 R so that interfaces exist."""
    rng = np.random.default_rng([seed, count, chain_length])
    lengths = np.full(count, chain_length)
    ca = _ca_walks(rng, lengths)
    radius = (3.0 * chain_length * 150.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    out = []
    for i in range(count):
        seq = rng.choice(20, size=chain_length, p=AA_FREQ)
        base = _assemble(rng, ca[i], seq)
        parts = []
        for c in range(copies):
            ang = 2.0 * np.pi * c / copies
            rot = np.array([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]])
            part = base.copy()
            shifted = base["x"] + np.array([0.9 * radius, 0.0, 0.0])
            part["x"] = np.round(shifted @ rot.T, 3)
            part["chainID2"] = ord("A") + c
            part["serial"] = base["serial"] + c * len(base)
            parts.append(part)
        out.append(Molecule._from_records(np.concatenate(parts), id=f"M{i % 1000:03d}"))
    return out


def _template_from_residues(rng, mol: Molecule, residues: Sequence[Tuple[int, int]], jitter: float,
                            weights: bool, tid: str) -> Template:
    a = mol._atoms
    atoms = []
    for chain, resseq in residues:
        sel = np.nonzero((a["chainID2"] == chain) & (a["resSeq"] == resseq))[0]
        res = a["resName"][sel[0]].decode()[:3]
        w = float(np.round(rng.uniform(0.0, 1.5), 2)) if weights else 0.0
        for name, mode, alts in MCSA_ATOMS[res]:
            tok = _encode_token(b"_" + name.encode() if len(name) < 4 else name.encode(), 4)
            hit = sel[a["name"][sel] == tok]
            xyz = np.round(a["x"][hit[0]] + rng.normal(0.0, jitter, size=3), 3)
            atoms.append(TemplateAtom(chain_id=chr(chain), residue_number=int(resseq), x=float(xyz[0]),
                                      y=float(xyz[1]), z=float(xyz[2]),
                                      residue_names=[res] + [_ONE[c] for c in alts], atom_names=[name],
                                      distance_weight=w, match_mode=mode))
    return Template(atoms, id=tid)


def templates(count: int, sources: Sequence[Molecule], residues=(3, 6), seed: int = 5000, jitter: float = 0.3,
              weights: bool = False, cross_chain: bool = False) -> List[Template]:
    """M-CSA-shaped templates cut out of `sources` (template i from source i mod len)."""
    out = []
    for i in range(count):
        rng = np.random.default_rng([seed, i])
        mol = sources[i % len(sources)]
        a = mol._atoms
        is_ca = a["name"] == b"_CA_"
        ca_xyz, ca_chain, ca_res = a["x"][is_ca], a["chainID2"][is_ca], a["resSeq"][is_ca]
        r = int(rng.integers(residues[0], residues[1] + 1))
        for _ in range(100):
            s = int(rng.integers(len(ca_xyz)))
            d = np.linalg.norm(ca_xyz - ca_xyz[s], axis=1)
            near = np.nonzero((d <= 10.0) & (d > 0))[0]
            if cross_chain and i % 2 == 1:
                other = near[ca_chain[near] != ca_chain[s]]
                if len(other) == 0:
                    continue
                # at least one residue from another chain
                rest = rng.permutation(near[near != other[0]])[: r - 2]
                pick = np.concatenate([[s, other[0]], rest])
            else:
                same = near[ca_chain[near] == ca_chain[s]] if cross_chain else near
                pick = np.concatenate([[s], rng.permutation(same)[: r - 1]])
            if len(pick) == r:
                break
        else:
            raise RuntimeError("could not place a template")
        res = [(int(ca_chain[p]), int(ca_res[p])) for p in pick]
        out.append(_template_from_residues(rng, mol, res, jitter, weights, f"T{i:04d}"))
    return out


# ---- named workloads ---------------------------------------------------------------

def workload(name: str, scale: float = 1.0, rank: int = 0):
    """Returns (templates, molecules, (rmsd, distance, max_dynamic), kwargs).

    cfg3       1,000 templates x 1,000 proteins of 300 residues      (2.0, 1.5, 1.5)
    cfg3dense  100 templates x 1,000 proteins, dense-hit thresholds   (2.0, 3.0, 3.0)
    cfg4shard  1,000 templates x 6,250 proteins of 300-1,500 residues (one GPU's share
               of the 50,000-structure config 4)                      (2.0, 1.5, 1.5)
    cfg5       1,000 8-residue templates with distance weights x 250 2,000-residue
               tetramers                                              (2.0, 1.5, 3.0)
    `scale` shrinks the counts (tests); `rank` selects a disjoint set of structures.
    Templates never depend on the rank (replicated on every GPU).
    """
    n = lambda v: max(1, int(round(v * scale)))  # noqa: E731
    src = proteins(n(1000), 300, seed=1000, first=0)
    if name in ("cfg3", "cfg3dense"):
        tpl = templates(n(1000 if name == "cfg3" else 100), src, (3, 6), seed=5000)
        mols = src if rank == 0 else proteins(n(1000), 300, seed=1000 + rank, first=0)
        params = (2.0, 1.5, 1.5) if name == "cfg3" else (2.0, 3.0, 3.0)
        return tpl, mols, params, {}
    if name == "cfg4shard":
        tpl = templates(n(1000), src, (3, 6), seed=5000)
        mols = proteins(n(6250), (300, 1500), seed=20000 + rank, first=0)
        return tpl, mols, (2.0, 1.5, 1.5), {}
    if name == "cfg5":
        mols = multimers(n(250), 500, 4, seed=3000 + rank)
        srcm = mols if rank == 0 else multimers(n(250), 500, 4, seed=3000)
        tpl = templates(n(1000), srcm, (8, 8), seed=7000, weights=True, cross_chain=True)
        return tpl, mols, (2.0, 1.5, 3.0), {"ignore_chain": None}
    raise ValueError(f"unknown workload {name!r}")
