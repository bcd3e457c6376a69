"""GPU parity tests proper: the CUDA engine, called through the C ABI, against
(a) the reference's golden vectors and (b) the CPU oracle on the same inputs.
Hit key sets must be bit-exact; RMSD / rotation / centroids within TOL = 1e-6."""
import os

import numpy as np
import pytest

from _util import TOL, assert_matches_golden, case_inputs, compare_tables, golden_cases
from pyjess_b200._tables import MoleculeBatch, TemplateTable

pytestmark = pytest.mark.gpu

CASES = sorted(golden_cases())
#: single-atom cases: the C ABI reports every candidate, the reference's skipped first
#: candidate is removed by the host layer (tests/test_api_gpu.py pins that to the golden)
ABI_CASES = [c for c in CASES if not golden_cases()[c].get("single_atom")]
COUNTERS = ("pairs", "live_pairs", "candidate_atoms", "frontier_items", "frontier_ints", "candidate_tests",
            "tuples", "nan_dropped")


def run_engine(engine, templates, mols, params):
    from pyjess_b200 import _capi
    tt = TemplateTable(templates)
    dev = _capi.DeviceTemplates(engine, tt)
    try:
        return _capi.query_host(engine, dev, MoleculeBatch(mols), params)
    finally:
        dev.close()


@pytest.mark.parametrize("name", ABI_CASES)
def test_engine_matches_reference_golden(engine, name):
    case = golden_cases()[name]
    templates, mols, params = case_inputs(case)
    table, stats = run_engine(engine, templates, mols, params)
    assert table.count == case["n_hits"]
    assert_matches_golden(table, case, best_match=bool(params.best_match))


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_oracle(engine, oracle_lib, name):
    case = golden_cases()[name]
    templates, mols, params = case_inputs(case)
    table, stats = run_engine(engine, templates, mols, params)
    ours, ostats = oracle_lib.query(TemplateTable(templates), MoleculeBatch(mols), params, threads=8)
    if params.best_match:
        assert table.count == ours.count
        assert np.abs(np.sort(table.rmsd) - np.sort(ours.rmsd)).max() <= TOL
        return
    assert compare_tables(table, ours, params.rmsd_threshold) == []
    # the algorithm-independent work counters must agree exactly (they define the roofline bytes)
    for key in COUNTERS:
        assert stats[key] == ostats[key], key
    assert stats["hits"] == table.count


# ---- synthetic BASELINE configs 3-5 at oracle-friendly sizes --------------------------
SYNTH = [("cfg3", 0.04, {}, 0), ("cfg3", 0.04, {}, 1), ("cfg3", 0.04, {}, 2),
         ("cfg3dense", 0.03, {}, 1), ("cfg3dense", 0.03, {}, 2), ("cfg4shard", 0.012, {}, 1), ("cfg4shard", 0.012, {}, 2),
         ("cfg5", 0.012, {"ignore_chain": None}, 1), ("cfg5", 0.012, {"ignore_chain": "residues"}, 1),
         ("cfg5", 0.006, {"ignore_chain": "atoms"}, 1)]


@pytest.mark.parametrize("name,scale,kw,reorder", SYNTH)
def test_synthetic_matches_oracle(engine, oracle_lib, name, scale, kw, reorder):
    from pyjess_b200 import synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw0 = synth.workload(name, scale=scale)
    kw0.update(kw)
    params = query_params(*thr, reorder=reorder, **kw0)
    table, stats = run_engine(engine, tpl, mols, params)
    ours, ostats = oracle_lib.query(TemplateTable(tpl), MoleculeBatch(mols), params, threads=8)
    assert table.count > 0
    ill = []
    assert compare_tables(table, ours, params.rmsd_threshold, tpl, mols, ill) == []
    assert len(ill) <= max(2, table.count // 100), ill      # near-coplanar sets only, listed separately
    for key in COUNTERS:
        assert stats[key] == ostats[key], key
    assert stats["scan_tests"] <= stats["candidate_tests"]


def test_scan_order_and_group_invariance_full_cfg3(engine):
    """Full-size property (no oracle): the hit set of BASELINE config 3 does not depend on
    the scan order, and every hit satisfies rmsd <= threshold."""
    from pyjess_b200 import synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload("cfg3", scale=0.3)
    tables = []
    for reorder in (1, 2):
        table, stats = run_engine(engine, tpl, mols, query_params(*thr, reorder=reorder, **kw))
        tables.append(table)
        assert stats["hits"] == table.count and (table.rmsd <= thr[0]).all()
    assert tables[0].keys() == tables[1].keys()
    assert np.abs(tables[0].rmsd - tables[1].rmsd).max() <= TOL


@pytest.mark.parametrize("mode", ["masks", "plain", "overflow"])
def test_every_classify_path_builds_the_same_candidate_lists(engine, oracle_lib, monkeypatch, mode):
    """k_classify has three ways to the candidate lists: the class-mask path (one 32 x 32 bit transpose per 32 atoms and
    32 specs), the plain loop over all specs (JB_CLASSIFY_PLAIN, or a template set whose masks do not fit), and the plain
    loop a structure falls back to at run time when it holds more distinct (atom name, residue name) pairs than the
    shared-memory class set (here: 900 made-up residue names on atoms no template matches).  Same hits, same work
    counters as the oracle on all three -- the counters see every list's length."""
    from pyjess_b200 import synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload("cfg3", scale=0.03)
    mols = [m.copy() for m in mols]
    if mode == "plain":
        monkeypatch.setenv("JB_CLASSIFY_PLAIN", "1")
    if mode == "overflow":
        rng = np.random.default_rng(3)
        letters = np.frombuffer(b"QWXZJ0123456789", dtype=np.uint8)
        for m in mols[::2]:
            at = m._atoms
            pick = rng.choice(len(at), size=min(900, len(at) // 2), replace=False)
            # residue names no template lists (digits / rare letters), all different: > 768 classes in one structure
            at["resName"][pick] = np.array([letters[[i % 15, (i // 15) % 15, (i // 225) % 15]].tobytes()
                                            for i in range(len(pick))], dtype="S3")
            m._pack = None
    params = query_params(*thr, reorder=2, **kw)
    table, stats = run_engine(engine, tpl, mols, params)
    ours, ostats = oracle_lib.query(TemplateTable(tpl), MoleculeBatch(mols), params, threads=8)
    assert table.count > 0
    ill = []
    assert compare_tables(table, ours, params.rmsd_threshold, tpl, mols, ill) == []
    for key in COUNTERS:
        assert stats[key] == ostats[key], key


@pytest.mark.parametrize("env", ["JB_ORDER_LSD", "JB_ORDER_NO_LSD", "JB_ORDER_COUNTING"])
@pytest.mark.parametrize("hit_order", ["canonical", "reference"])
def test_every_hit_ordering_path_gives_the_same_records(engine, monkeypatch, env, hit_order):
    """hits.cuh has three ways to establish the record order (keyed radix sort + tie ranking, its
    multi-pass LSD fallback for dense ties, the counting-sort grouping for huge structures):
    all must produce the same table."""
    from pyjess_b200 import synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload("cfg3dense", scale=0.05)
    tpl, mols = tpl[:5], mols[:20]
    params = query_params(*thr, hit_order=hit_order, **kw)

    def run():
        from pyjess_b200 import _capi
        dev = _capi.DeviceTemplates(engine, TemplateTable(tpl))
        try:
            return _capi.query_host(engine, dev, MoleculeBatch(mols, serial=True), params)[0]
        finally:
            dev.close()
    base = run()
    monkeypatch.setenv(env, "1")
    other = run()
    assert base.count == other.count > 1000
    for a, b in zip(base.arrays(), other.arrays()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("hit_order", ["canonical", "reference"])
def test_pipelined_host_path_gives_the_same_records(engine, monkeypatch, hit_order):
    """jb_query_host in slices (H2D / kernels / D2H overlapped on a copy stream) against the
    single-shot path: identical tables, for both record orders, incl. best_match."""
    from pyjess_b200 import _capi, synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload("cfg3", scale=0.1)
    tpl = tpl[:40]
    monkeypatch.setenv("JB_PIPELINE", "3")
    sliced = _capi.Context(0)            # reads JB_PIPELINE at creation
    try:
        for best in (False, True):
            params = query_params(*thr, hit_order=hit_order, best_match=best, **kw)
            tables = []
            for ctx in (engine, sliced):
                dev = _capi.DeviceTemplates(ctx, TemplateTable(tpl))
                try:
                    for _ in range(2):   # second call: the pinned block is sized from the first one
                        table, stats = _capi.query_host(ctx, dev, MoleculeBatch(mols, serial=True), params)
                    tables.append(table)
                finally:
                    dev.close()
            assert tables[0].count == tables[1].count > 50
            for a, b in zip(tables[0].arrays(), tables[1].arrays()):
                assert np.array_equal(a, b)
    finally:
        sliced.close()


@pytest.mark.parametrize("size", [2, 3, 4])
def test_short_templates_match_oracle(engine, oracle_lib, size):
    """Templates of 2-4 atoms: the fused level-0/1 kernel emits complete tuples (n == 2) or hands
    level-2 items on; engine vs oracle: tuple set and work counters exact.  With 2 or 3 atoms the
    covariance is rank deficient and the reference's rotation is 0/0 (vendor/jess/src/Super.c:
    501-520 divides by sqrt of a zero eigenvalue): whether it comes out NaN -- and the hit is
    dropped, _jess.pyx:2240-2251 -- depends on libm rounding noise, so for those sizes the HIT
    sets are only required to agree on tuples whose rotation is finite on both sides."""
    from _util import load_molecule, load_template
    from pyjess_b200 import Template
    from pyjess_b200._tables import query_params
    full = [load_template(n) for n in ("template_01.qry", "template_02.qry", "1.3.3.tpl", "4.1.2.tpl")]
    tpl = [Template(list(t)[k:k + size], id=f"{t.id}-{k}") for t in full for k in (0, 3)]
    mols = [load_molecule("pdb1lnb.ent"), load_molecule("1AMY.pdb")]
    for reorder in (1, 2):
        params = query_params(2.0, 1.0, 1.0, reorder=reorder)
        table, stats = run_engine(engine, tpl, mols, params)
        ours, ostats = oracle_lib.query(TemplateTable(tpl), MoleculeBatch(mols), params, threads=4)
        for key in ("pairs", "live_pairs", "candidate_atoms", "frontier_items", "frontier_ints", "candidate_tests",
                    "tuples"):
            assert stats[key] == ostats[key], key
        every, _ = run_engine(engine, tpl, mols, query_params(2.0, 1.0, 1.0, reorder=reorder, all_tuples=True))
        assert every.count == stats["tuples"] and set(table.keys()) <= set(every.keys()) >= set(ours.keys())
        if size >= 4:
            assert table.count == ours.count > 0
            assert compare_tables(table, ours, params.rmsd_threshold, tpl, mols, []) == []
            assert stats["nan_dropped"] == ostats["nan_dropped"]
        else:
            assert table.count + stats["nan_dropped"] == ours.count + ostats["nan_dropped"] > 0   # same RMSD filter
            mine = {k: table.rmsd[i] for i, k in enumerate(table.keys())}
            for i, k in enumerate(ours.keys()):
                if k in mine:
                    assert abs(mine[k] - ours.rmsd[i]) <= TOL


def test_small_workspace_splits_chunks_and_fails_loudly(engine, monkeypatch):
    """The frontier workspace is bounded (`workspace_bytes`): an overflowing chunk is re-run with
    larger buffers or split into smaller runs of structures -- same table -- and a single
    structure that cannot fit raises JB_ERR_CAPACITY instead of truncating.  (Level-synchronous
    sweep: the depth-first sweep of small chunks keeps no frontier, JB_DFS_PAIRS=0 turns it off.)"""
    from pyjess_b200 import _capi, synth
    from pyjess_b200._tables import query_params
    monkeypatch.setenv("JB_DFS_PAIRS", "0")
    tpl, mols, thr, kw = synth.workload("cfg3dense", scale=0.05)
    tpl, mols = tpl[:5], mols[:24]
    params = query_params(*thr, **kw)
    base, _ = run_engine(engine, tpl, mols, params)
    small = _capi.Context(0, workspace_bytes=1 << 20)
    try:
        table, stats = run_engine(small, tpl, mols, params)
        assert stats["retries"] > 0 and stats["chunks"] > 1
        assert base.count == table.count > 1000
        for a, b in zip(base.arrays(), table.arrays()):
            assert np.array_equal(a, b)
    finally:
        small.close()
    tiny = _capi.Context(0, workspace_bytes=1 << 12)
    try:
        with pytest.raises(_capi.EngineError) as err:
            run_engine(tiny, tpl, mols, params)
        assert err.value.status == -4
    finally:
        tiny.close()


def test_candidate_record_overflow_is_retried(engine, oracle_lib, monkeypatch):
    """Residue-independent match modes (100-107: most atoms are candidates of several specs) against
    the oracle, and the candidate-record buffer overflowing on the first attempt (forced small with
    JB_TEST_CREC): the chunk is re-run with a larger buffer, same table."""
    from pyjess_b200 import Template, TemplateAtom, _capi, synth
    from pyjess_b200._tables import query_params
    _, mols, _, _ = synth.workload("cfg3", scale=0.03)
    mols = mols[:12]
    src = mols[0]._atoms
    bb = [int(np.nonzero((src["resSeq"] == 5) & (src["name"] == n))[0][0]) for n in (b"_N__", b"_CA_", b"_C__", b"_O__")]
    # modes that accept N / CA / C / O respectively (mainchain = CA, N, O; eff:TessAtom.c:269-276)
    sets = [[(104, "N"), (105, "CA"), (106, "C"), (102, "O")], [(107, "N"), (100, "CA"), (103, "C"), (104, "O")],
            [(102, "N"), (107, "CA"), (103, "CA"), (105, "O")]]
    tpl = []
    for k, modes in enumerate(sets):
        atoms = []
        for (mode, name), a in zip(modes, bb):
            x, y, z = (float(v) for v in src["x"][a])
            atoms.append(TemplateAtom(chain_id="A", residue_number=5, x=x, y=y, z=z, residue_names=["ALA"],
                                      atom_names=[name], distance_weight=0.0, match_mode=mode))
        tpl.append(Template(atoms, id=f"any-{k}"))
    params = query_params(0.5, 0.3, 0.3)
    base, bstats = run_engine(engine, tpl, mols, params)
    ours, ostats = oracle_lib.query(TemplateTable(tpl), MoleculeBatch(mols), params, threads=8)
    assert base.count == ours.count > 1000
    assert compare_tables(base, ours, params.rmsd_threshold, tpl, mols, []) == []
    for key in COUNTERS:
        assert bstats[key] == ostats[key], key
    monkeypatch.setenv("JB_TEST_CREC", "4096")
    fresh = _capi.Context(0)             # its candidate buffer starts at 4,096 records
    try:
        table, stats = run_engine(fresh, tpl, mols, params)
    finally:
        fresh.close()
    assert stats["retries"] > 0 and stats["candidate_records"] > 4096
    for a, b in zip(base.arrays(), table.arrays()):
        assert np.array_equal(a, b)


def test_context_trim_releases_buffers_and_runs_again(engine, monkeypatch):
    """jb_context_trim: the grow-only device buffers go back to the driver, the next run
    re-allocates them and returns the same table; hits of a run that was not fetched before the
    trim are refused instead of being read from freed memory."""
    import torch
    from pyjess_b200 import _capi, synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload("cfg3", scale=0.05)
    params = query_params(*thr, **kw)
    base, _ = run_engine(engine, tpl, mols, params)
    free0, _ = torch.cuda.mem_get_info()
    engine.trim()
    free1, _ = torch.cuda.mem_get_info()
    assert free1 > free0
    again, _ = run_engine(engine, tpl, mols, params)
    assert again.count == base.count > 0
    for a, b in zip(base.arrays(), again.arrays()):
        assert np.array_equal(a, b)
    # un-fetched result across a trim
    dev = _capi.DeviceTemplates(engine, TemplateTable(tpl))
    dmol = _capi.DeviceMolecules(engine, MoleculeBatch(mols))
    try:
        # a small result already sits in its page-locked host block (k_hits_small): it survives the trim
        res = _capi.query_run(engine, dev, dmol, params)
        engine.trim()
        assert base.count <= 1024                                 # kSmallHits
        kept = res.fetch()
        assert kept.count == base.count and np.array_equal(kept.rmsd, base.rmsd)
        res.close()
        # ... a result that still lives in the device buffers does not: refused, never garbage
        monkeypatch.setenv("JB_ORDER_GENERAL", "1")
        res = _capi.query_run(engine, dev, dmol, params)
        engine.trim()
        with pytest.raises(_capi.EngineError, match="jb_context_trim"):
            res.fetch()
        res.close()
        later = _capi.query_run(engine, dev, dmol, params)
        _capi.query_run(engine, dev, dmol, params).close()
        with pytest.raises(_capi.EngineError, match="later run"):
            later.fetch()                                       # overwritten by the run after it (run epoch)
        later.close()
        fresh = _capi.query_run(engine, dev, dmol, params).fetch()
        assert fresh.count == base.count
    finally:
        dmol.close()
        dev.close()


def test_oversized_buffers_decay_after_a_few_small_runs(monkeypatch):
    """High-water marks are not forever (context_decay): after a large batch, a handful of small
    runs give the candidate / pair / hit buffers back, and a steady workload keeps its buffers.
    (JB_DECAY_BYTES / JB_DECAY_RUNS lower the 1 GiB / 64-call thresholds so that a test-sized workload shows it.)"""
    import torch
    from _util import load_molecule, load_template
    from pyjess_b200 import _capi, synth
    from pyjess_b200._tables import query_params
    monkeypatch.setenv("JB_DECAY_BYTES", str(1 << 20))
    monkeypatch.setenv("JB_DECAY_RUNS", "8")
    ctx = _capi.Context(0)
    used = lambda: (lambda f, t: t - f)(*torch.cuda.mem_get_info())  # noqa: E731
    try:
        tpl, mols, thr, kw = synth.workload("cfg3", scale=0.3)
        big = query_params(*thr, **kw)
        first, _ = run_engine(ctx, tpl, mols, big)
        steady = []
        for _ in range(18):                                     # steady large workload: nothing is released
            again, _ = run_engine(ctx, tpl, mols, big)
            steady.append(used())
        assert max(steady) - min(steady) < 4 << 20
        assert again.count == first.count and all(np.array_equal(a, b) for a, b in zip(first.arrays(), again.arrays()))
        few = [load_template("template_01.qry"), load_template("template_02.qry")]
        one = [load_molecule("pdb1lnb.ent")]
        small = query_params(2.0, 3.0, 3.0)
        base, _ = run_engine(ctx, few, one, small)
        low = min(used() for _ in range(18) if run_engine(ctx, few, one, small))
        assert steady[-1] - low > 16 << 20
        back, _ = run_engine(ctx, tpl, mols, big)               # and the large batch still runs
        assert back.count == first.count and all(np.array_equal(a, b) for a, b in zip(first.arrays(), back.arrays()))
    finally:
        ctx.close()


@pytest.mark.parametrize("shift", [0.0, 731.25])
def test_shell_boundary_candidates_take_the_exact_path(engine, oracle_lib, shift):
    """The drain of the cell-grid sweep decides in FP32 and hands every candidate whose squared distance
    lies within its error band of lo^2 / hi^2 to the reference's FP64 operations (expand_grid.cuh).  Here
    most candidates sit ON the shell boundaries -- exactly, one FP64 ulp inside and outside, and a few
    1e-7 A steps around them -- also with the structure origin far away (large FP32 error band).  The
    tuple set and the counters must equal the oracle's (decisions bit for bit)."""
    from pyjess_b200 import Molecule, Template, TemplateAtom
    from pyjess_b200._tables import query_params
    rng = np.random.default_rng(7)
    dirs = [np.array(v, dtype=np.float64) for v in ((1, 0, 0), (0, 1, 0), (0, 0, -1), (0.6, 0.8, 0), (0, -0.6, 0.8),
                                                      (2 / 3, 1 / 3, 2 / 3), (-3 / 7, 6 / 7, 2 / 7))]
    radii = []
    for r in (5.0, 7.0, 3.0, 9.0):
        radii += [r, np.nextafter(r, 0), np.nextafter(r, 100), r - 1e-7, r + 1e-7, r - 3e-6, r + 3e-6, r - 1e-3, r + 1e-3]
    pts = [np.zeros(3)]
    for d in dirs:
        for r in radii:
            pts.append(d * r)
    for _ in range(60):
        v = rng.normal(size=3)
        pts.append(v / np.linalg.norm(v) * rng.choice([5.0, 7.0, 6.0, 3.0, 9.0]))
    xyz = np.array(pts) + np.array([shift, -shift / 2, shift / 4])
    xyz = np.concatenate([[[0.0, 0.0, 0.0]], xyz]) if shift else xyz      # first atom = origin of the FP32 frame
    n = len(xyz)
    mol = Molecule.from_arrays(name=["CA"] * n, residue_name=["ALA"] * n, residue_number=np.arange(1, n + 1),
                               chain_id="A", x=xyz[:, 0], y=xyz[:, 1], z=xyz[:, 2], id="EDGE")

    def tpl(d, tid, extra=()):
        atoms = [TemplateAtom(chain_id="A", residue_number=1 + k, x=p[0], y=p[1], z=p[2], residue_names=["ALA"],
                              atom_names=["CA"], distance_weight=0.0, match_mode=0)
                 for k, p in enumerate([(0.0, 0.0, 0.0), (d, 0.0, 0.0), *extra])]
        return Template(atoms, id=tid)
    # 2-atom templates: every decision is one shell test; the oracle drops their (rank-deficient, NaN) hits,
    # so the engine's tuple count and work counters are compared; the 4-atom template gives real hits
    templates = [tpl(6.0, "D6"), tpl(4.0, "D4"), tpl(8.0, "D8"),
                 tpl(6.0, "Q6", extra=[(0.0, 6.0, 0.0), (0.0, 0.0, -6.0)])]
    for reorder in (0, 1, 2):
        params = query_params(50.0, 1.0, 10.0, reorder=reorder)
        table, stats = run_engine(engine, templates, [mol], params)
        ours, ostats = oracle_lib.query(TemplateTable(templates), MoleculeBatch([mol]), params, threads=2)
        assert stats["tuples"] == ostats["tuples"] > 200
        for key in ("frontier_items", "frontier_ints", "candidate_tests"):
            assert stats[key] == ostats[key], key
        mine = {k for k in table.keys() if k[1] == 3}
        assert mine == {k for k in ours.keys() if k[1] == 3} and len(mine) > 0


@pytest.mark.parametrize("workload,scale", [("cfg3", 0.05), ("cfg3dense", 0.04), ("cfg5", 0.012)])
def test_depth_first_sweep_equals_level_sweep(workload, scale, monkeypatch):
    """Small chunks take ONE depth-first launch (expand_dfs.cuh) instead of one launch per level:
    same hit table, same algorithm-independent work counters, all three ignore_chain modes and
    both scan orders; with single-atom templates in the set (a root is a complete tuple)."""
    from pyjess_b200 import Template, _capi, synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload(workload, scale=scale)
    tpl = list(tpl)[:40] + [Template(list(tpl[0])[:1], id="single"), Template(list(tpl[1])[:2], id="pair")]
    mols = mols[:3]
    ctxs = {}
    for name, pairs in (("bfs", "0"), ("dfs", "1000000")):
        monkeypatch.setenv("JB_DFS_PAIRS", pairs)
        ctxs[name] = _capi.Context(0)
    try:
        for ignore in (None, "residues", "atoms"):
            for reorder in (2, 1, 0):
                kw2 = dict(kw, ignore_chain=ignore)
                params = query_params(*thr, reorder=reorder, all_tuples=True, **kw2)
                got = {name: run_engine(c, tpl, mols, params) for name, c in ctxs.items()}
                (ta, sa), (tb, sb) = got["bfs"], got["dfs"]
                assert ta.count == tb.count > 0
                for a, b in zip(ta.arrays(), tb.arrays()):       # (every tuple: rank-deficient ones carry NaN rotations)
                    assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
                for key in ("live_pairs", "candidate_atoms", "candidate_tests", "frontier_items", "frontier_ints", "tuples", "hits"):
                    assert sa[key] == sb[key], key
                assert sb["launches"] < sa["launches"]
    finally:
        for c in ctxs.values():
            c.close()


@pytest.mark.parametrize("workload,scale", [("cfg3", 0.06), ("cfg3dense", 0.04), ("cfg5", 0.012)])
def test_tail_finish_equals_level_sweep(workload, scale, monkeypatch):
    """The level sweep may hand a SMALL frontier to a depth-first finish (JB_TAIL_ITEMS, expand_dfs.cuh: TAIL): same hit
    table and the same work counters whether the hand-over never happens (0), happens at a late level (a few thousand
    partial matches) or at the first level it may (any frontier qualifies); DFS chunks switched off so that the level
    kernels run."""
    from pyjess_b200 import _capi, synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload(workload, scale=scale)
    monkeypatch.setenv("JB_DFS_PAIRS", "0")
    ctxs = {}
    for name, items in (("levels", "0"), ("late", "3000"), ("always", "1000000000")):
        monkeypatch.setenv("JB_TAIL_ITEMS", items)
        ctxs[name] = _capi.Context(0)
    try:
        for ignore, reorder in ((kw.get("ignore_chain"), 2), ("residues", 1), ("atoms", 0)):
            params = query_params(*thr, reorder=reorder, all_tuples=True, **dict(kw, ignore_chain=ignore))
            got = {name: run_engine(c, tpl, mols, params) for name, c in ctxs.items()}
            ta, sa = got["levels"]
            assert ta.count > 0
            for name in ("late", "always"):
                tb, sb = got[name]
                assert ta.count == tb.count
                for a, b in zip(ta.arrays(), tb.arrays()):
                    assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f")
                for key in ("live_pairs", "candidate_atoms", "candidate_tests", "frontier_items", "frontier_ints", "tuples", "hits"):
                    assert sa[key] == sb[key], (name, key)
    finally:
        for c in ctxs.values():
            c.close()


@pytest.mark.parametrize("workload,scale,kwargs", [
    ("cfg3dense", 0.5, {}),                       # 50 templates x 500 structures, dense-hit thresholds: 2.2 M hits
    ("cfg4shard", 0.2, {}),                       # 200 templates x 1,250 structures of 300-1,500 residues
    ("cfg5", 0.2, {"ignore_chain": None}),        # 200 24-atom weighted templates x 50 tetramers
    ("cfg5", 0.1, {"ignore_chain": "residues"}),
    ("cfg5", 0.1, {"ignore_chain": "atoms"}),
])
def test_large_samples_match_the_unmodified_reference_library(engine, workload, scale, kwargs):
    """VERDICT r1 (parity, item 1): the BASELINE shapes at the largest scale the reference's CPU path
    finishes in ~10-30 s, against oracle/_ref (the UNMODIFIED reference C library, not the port):
    hit key sets bit-exact, RMSD / rotation / centroids within 1e-6, ill-conditioned superpositions
    and threshold-edge hits counted separately (bench.parity_report -- the same check the bench line
    carries for its timed result)."""
    import oracle
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libjess_ref.so not built")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import parity_report
    from pyjess_b200 import synth
    from pyjess_b200._tables import query_params
    tpl, mols, thr, kw = synth.workload(workload, scale=scale)
    kw = dict(kw, **kwargs)
    params = query_params(*thr, **kw)
    got, _ = run_engine(engine, tpl, mols, params)
    want, _ = oracle.ref_query(TemplateTable(tpl), oracle.numpy_batch(mols), params, threads=len(os.sched_getaffinity(0)))
    rep = parity_report(got, want, thr[0], tpl, mols)
    print(rep)
    assert rep["keys"].startswith("identical") and rep["values"] == "within tolerance", rep
    assert rep["hits_reference"] > 1000
    # listed separately, and rare: one-sided rank-deficient superpositions (found by this very test at scale 1.0: 1 of 8,953,863
    # cfg3dense hits, sigma_min / sigma_max = 4e-9, the reference's own rotation there has det R = 4.07),
    # ill-conditioned rotations, hits within 1e-6 of the RMSD threshold
    assert rep.get("rank_deficient_one_sided", 0) <= 1e-5 * rep["hits_reference"] + 1, rep
    assert rep["ill_conditioned"] <= 0.01 * rep["hits_gpu"], rep
    assert rep["threshold_edge"] <= 1e-5 * rep["hits_reference"] + 2, rep
