"""CPU-only behaviour of the drop-in surface that needs no engine call (reference:
src/pyjess/tests/test_jess.py:18-94, test_template.py:128-132, test_atom.py)."""
import pickle

import pytest

from _util import load_molecule, load_template
from pyjess_b200 import Atom, Jess, Molecule, Template, TemplateAtom


def test_empty_jess_and_empty_molecule_need_no_gpu():
    jess = Jess()
    assert len(jess) == 0 and not jess
    with pytest.raises(StopIteration):
        next(jess.query(Molecule(), 2, 2, 4))
    with pytest.raises(StopIteration):
        next(Jess([load_template("template_01.qry")]).query(Molecule(), 2, 2, 4))


def test_jess_container_protocol():
    t1, t2 = load_template("template_01.qry"), load_template("template_02.qry")
    j1, j2, j3 = Jess([t1, t2]), Jess([t1, t2]), Jess([t1])
    assert j1 == j2 and j1 is not j2 and hash(j1) == hash(j2)
    assert j1 != j3 and hash(j1) != hash(j3)
    assert j1.copy() == j1 and pickle.loads(pickle.dumps(j1)) == j1
    assert j1[0] is t1 and j1[-1] is t2 and len(j1[:1]) == 1 and len(j1[2:]) == 0
    with pytest.raises(IndexError):
        j1[2]
    with pytest.raises(TypeError):
        Jess([object()])
    assert Jess() == Jess() and hash(Jess()) == hash(Jess())


def test_query_argument_validation_matches_reference():
    jess = Jess([load_template("template_01.qry")])
    mol = load_molecule("pdb1a0p.ent")
    with pytest.raises(ValueError):
        jess.query(mol, 2, 2, 2, ignore_chain="all")
    with pytest.raises(ValueError):
        jess.query(mol, 2, 2, 2, max_candidates=-1)
    with pytest.raises(TypeError):
        jess.query(object(), 2, 2, 2)
    q = jess.query(mol, 2, 3, 4, best_match=True, ignore_chain="atoms")
    assert q.jess is jess and q.molecule is mol and q.rmsd_threshold == 2.0
    assert q.best_match is True and q.ignore_chain == "atoms" and q.max_candidates is None


def test_template_dimension_counts_residues_across_chains():
    assert load_template("1sur.qry").dimension == 4          # test_template.py:128-132
    assert load_template("template_01.qry").dimension == 4


def test_template_roundtrips():
    tpl = load_template("1.3.3.tpl")
    assert Template.loads(tpl.dumps(), id=tpl.id) == tpl
    assert pickle.loads(pickle.dumps(tpl)) == tpl and tpl.copy() == tpl
    assert [a.match_mode for a in tpl][:3] == [a.match_mode for a in Template.loads(tpl.dumps())][:3]


def test_atom_parse_and_token_rules():
    """Names shorter than 4 are right-shifted by one '_' (SURVEY 7.3-h)."""
    atom = Atom.loads("ATOM   1112  ND1 HIS E 142      47.130  25.140  -6.840  1.00 10.22           N")
    assert (atom.serial, atom.name, atom.residue_name, atom.chain_id, atom.residue_number) == \
        (1112, "ND1", "HIS", "E", 142)
    assert (atom.x, atom.y, atom.z) == (47.13, 25.14, -6.84)
    assert pickle.loads(pickle.dumps(atom)) == atom
    built = Atom(serial=1112, name="ND1", residue_name="HIS", chain_id="E", residue_number=142,
                 x=47.13, y=25.14, z=-6.84, occupancy=1.0, temperature_factor=10.22, element="N")
    assert built.name == "ND1" and bytes(built._rec["name"][0]) == b"_ND1"
    tatom = TemplateAtom(chain_id="A", residue_number=1, x=0, y=0, z=0, residue_names=["HIS"],
                         atom_names=["ND1"], distance_weight=0.0, match_mode=0)
    assert list(tatom.atom_names) == ["ND1"] and list(tatom.residue_names) == ["HIS"]
    with pytest.raises(ValueError):
        TemplateAtom(chain_id="A", residue_number=1, x=0, y=0, z=0, residue_names=["HIS"], atom_names=["ND1"],
                     distance_weight=0.0, match_mode=42)


def test_molecule_basics():
    mol = load_molecule("pdb1lnb.ent")
    assert mol.id == "1LNB" and len(mol) > 2000
    assert len(mol.conserved(10.0)) < len(mol)
    assert pickle.loads(pickle.dumps(mol)) == mol and mol.copy() == mol
    assert len(mol[:10]) == 10 and mol[0].serial == 1


def test_iteration_rows_mirror_equals_the_sorted_formulation():
    """`_iteration_rows` (runs of a structure mirrored in place) == a lexsort of the runs by
    (structure ascending, template descending), on random tables incl. filtered rows."""
    import numpy as np
    from pyjess_b200._jess import _iteration_rows
    rng = np.random.default_rng(3)
    for case in range(20):
        n = int(rng.integers(1, 3000))
        mol = np.sort(rng.integers(0, 40, n)).astype(np.int32)
        tpl = rng.integers(0, 12 if case % 2 else 400, n).astype(np.int32)
        order = np.lexsort((tpl, mol))
        mol, tpl = mol[order], tpl[order]
        rows = np.nonzero(rng.random(n) < 0.9)[0] if case % 3 == 0 else np.arange(n)
        got = _iteration_rows(mol, tpl, rows.astype(np.int64))
        want = rows[np.lexsort((np.arange(len(rows)), -tpl[rows].astype(np.int64), mol[rows]))]
        assert np.array_equal(got, want)
        if case % 3:
            assert np.array_equal(_iteration_rows(mol, tpl, None), want)       # None: every record
    assert len(_iteration_rows(mol, tpl, np.zeros(0, dtype=np.int64))) == 0
    assert len(_iteration_rows(mol[:0], tpl[:0], None)) == 0


def test_batch_run_slices_with_a_stubbed_engine(monkeypatch):
    """Host logic of `_BatchRun.run` (no engine call): per structure, the records come in
    descending template order, table order inside a (structure, template) run, the first record
    of single-atom templates dropped (the reference never reports its first candidate)."""
    import types
    import numpy as np
    from pyjess_b200 import Jess, Template, _capi, synth
    from pyjess_b200 import _jess
    tpl, mols, thr, _ = synth.workload("cfg3", scale=0.01)
    tpl = list(tpl) + [Template(list(tpl[0])[:1], id="single")]
    rng = np.random.default_rng(11)
    n = 5000
    mol = np.sort(rng.integers(0, len(mols), n)).astype(np.int32)
    t = rng.integers(0, len(tpl), n).astype(np.int32)
    order = np.lexsort((t, mol))
    mol, t = mol[order], t[order]
    table = types.SimpleNamespace(count=n, molecule=mol, template=t, atoms=np.zeros((n, 18), np.int32),
                                  rmsd=rng.uniform(0, 2, n), rotation=np.zeros((n, 3, 3)), centre0=np.zeros((n, 3)),
                                  centre1=np.zeros((n, 3)))
    monkeypatch.setattr(_capi, "default_context", lambda device=None: None)
    monkeypatch.setattr(Jess, "_device_templates", lambda self, ctx: None)
    monkeypatch.setattr(_capi, "query_records", lambda ctx, dev, molecules, params: (table, {"nan_dropped": 0}))
    jess = Jess(tpl)
    single = len(tpl) - 1
    seen = 0
    for m, query in enumerate(jess.query_many(mols, *thr)):
        _, rows = query.hit_table()
        want = [i for i in range(n) if mol[i] == m]
        first_single = next((i for i in want if t[i] == single), None)
        want = [i for i in want if i != first_single]
        want.sort(key=lambda i: (-int(t[i]), i))
        assert rows.tolist() == want
        hits = list(query)
        assert [h.rmsd for h in hits] == [float(table.rmsd[i]) for i in want]
        assert all(h.template() is tpl[t[i]] and h.molecule() is mols[m] for h, i in zip(hits, want))
        seen += len(rows)
    assert seen == n - len({int(mol[i]) for i in range(n) if t[i] == single})


def test_template_atom_residue_codes_are_reread_like_the_reference():
    """``TessAtom_create`` does not step over the one-letter residue codes it has just read
    (eff:TessAtom.c:236-252): while names are still owed, the SAME run of codes is read again.
    Expected values: the reference library (pyjess 0.10.1) on these lines; found by a differential
    fuzz of ``Template.loads`` (4,500 mutated templates, no other difference)."""
    from pyjess_b200 import TemplateAtom
    cases = {
        "ATOM      3  OE1ZGLU A  43      42.721 -29.948   8.832 E   C 2.08 ": ("GLU", "GLU", "GLU"),
        "ATOM      0  NE1ZTRP A 204      43.868   1.684  54.342 W     L0.00 ": ("TRP", "TRP", "TRP"),
        "ATOM      1  OG ZSER A  10       1.000   2.000   3.000 TY    1.50 ": ("SER", "THR", "TYR"),
        ("ATOM      0  CZ2ZTRP A 204      41.395   1.476  53.906 W     0.00 "
         "ATOM      0  CH2ZTRP A 204      40.173   1.626  54.507 W     0.00 "): ("TRP",) * 13,
    }
    for line, want in cases.items():
        assert TemplateAtom.loads(line).residue_names == want
    # an atom name still owed after a run of codes: the reference re-reads the codes for ever (and
    # writes past its allocation); refused here
    import pytest
    with pytest.raises(ValueError):
        TemplateAtom.loads("ATOM      1  OG ZSER A  10       1.000   2.000   3.000 T (OG1) 1.50 ")


# public members of the reference's classes (dir() of the PyJess 0.10.1 build under baseline/_ref)
REFERENCE_SURFACE = {
    "Atom": ["altloc", "chain_id", "charge", "copy", "dump", "dumps", "element", "insertion_code",
             "load", "loads", "name", "occupancy", "residue_name", "residue_number", "segment",
             "serial", "temperature_factor", "transform", "x", "y", "z"],
    "Molecule": ["conserved", "copy", "date", "dump", "dumps", "from_biopython", "from_biotite",
                 "from_gemmi", "id", "load", "loads", "name", "transform"],
    "TemplateAtom": ["atom_names", "chain_id", "copy", "distance_weight", "dump", "dumps", "load",
                     "loads", "match_mode", "residue_names", "residue_number", "transform", "x", "y", "z"],
    "Template": ["copy", "dimension", "dump", "dumps", "id", "load", "loads", "transform"],
    "Jess": ["copy", "query"],
    "Query": ["best_match", "ignore_chain", "jess", "max_candidates", "molecule", "rmsd_threshold"],
    "Hit": ["atoms", "determinant", "dump", "dumps", "evalue", "inverse_transformation",
            "log_evalue", "molecule", "rmsd", "template", "transformation"],
}


def test_every_public_member_of_the_reference_classes_exists():
    import pyjess_b200
    assert set(REFERENCE_SURFACE) <= set(dir(pyjess_b200))
    for cls, members in REFERENCE_SURFACE.items():
        missing = [m for m in members if not hasattr(getattr(pyjess_b200, cls), m)]
        assert not missing, (cls, missing)


def _stub_engine(monkeypatch, mols, table, calls=None, delay=0.0):
    """`_capi.query_records` answered from a global hit table (numbered by position in `mols`): the
    call's own structures are looked up by identity and the rows re-numbered to the call."""
    import time
    import types
    import numpy as np
    from pyjess_b200 import Jess, _capi
    where = {id(m): i for i, m in enumerate(mols)}

    def query_records(ctx, dev, molecules, params):
        ids = np.asarray([where[id(m)] for m in molecules], dtype=np.int64)
        local = np.full(len(mols), -1, dtype=np.int64)
        local[ids] = np.arange(len(ids))
        rows = np.nonzero(local[table.molecule] >= 0)[0]
        rows = rows[np.argsort(local[table.molecule[rows]], kind="stable")]
        if calls is not None:
            calls.append((ctx, ids.tolist()))
        if delay:
            time.sleep(delay)
        sub = types.SimpleNamespace(count=len(rows), molecule=local[table.molecule[rows]].astype(np.int32),
                                    template=table.template[rows].copy(), atoms=table.atoms[rows].copy(),
                                    rmsd=table.rmsd[rows].copy(), rotation=table.rotation[rows].copy(),
                                    centre0=table.centre0[rows].copy(), centre1=table.centre1[rows].copy(), coords=None)
        return sub, {"nan_dropped": 0, "hits": len(rows)}

    monkeypatch.setattr(_capi, "default_context", lambda device=None: device)
    monkeypatch.setattr(Jess, "_device_templates", lambda self, ctx: None)
    monkeypatch.setattr(_capi, "query_records", query_records)


def _random_table(n_mols, n_tpl, n, seed):
    import types
    import numpy as np
    rng = np.random.default_rng(seed)
    mol = np.sort(rng.integers(0, n_mols, n)).astype(np.int32)
    t = rng.integers(0, n_tpl, n).astype(np.int32)
    order = np.lexsort((t, mol))
    mol, t = mol[order], t[order]
    return types.SimpleNamespace(count=n, molecule=mol, template=t,
                                 atoms=rng.integers(0, 50, (n, 18)).astype(np.int32), rmsd=rng.uniform(0, 2, n),
                                 rotation=np.zeros((n, 3, 3)), centre0=np.zeros((n, 3)), centre1=np.zeros((n, 3)))


def test_sharded_runs_equal_the_single_run(monkeypatch):
    """`query_many(devices=[...])`: structures dealt to the devices by atom count (LPT), one hit
    table per device -- every Query must see exactly the records (and order) of the one-device
    run.  Host logic only: the engine is stubbed."""
    import numpy as np
    from pyjess_b200 import Jess, synth
    tpl, mols, thr, _ = synth.workload("cfg3", scale=0.04)
    table = _random_table(len(mols), len(tpl), 4000, 5)
    jess = Jess(tpl)
    _stub_engine(monkeypatch, mols, table)
    base = [(q.hit_table()[1].tolist(), [(h.rmsd, h.template()) for h in q]) for q in jess.query_many(mols, *thr)]
    want = [([int(table.molecule[i]) for i in rows], [float(table.rmsd[i]) for i in rows]) for rows, _ in base]
    for devices in ([0, 1, 2], [1, 0], [3], "2,5"):
        calls = []
        _stub_engine(monkeypatch, mols, table, calls)
        queries = jess.query_many(mols, *thr, devices=devices)
        if isinstance(devices, str):
            devices = [2, 5]
        for m, q in enumerate(queries):
            t, rows = q.hit_table()
            assert [int(v) for v in t.molecule[rows]] == want[m][0] == [m] * len(rows)   # caller's numbering
            assert [float(v) for v in t.rmsd[rows]] == want[m][1]
            assert [(h.rmsd, h.template()) for h in q] == base[m][1]
        assert sorted(i for _, ids in calls for i in ids) == list(range(len(mols)))          # a partition
        assert {c for c, _ in calls} == set(devices)
        assert len(calls) == len(devices) and all(ids == sorted(ids) for _, ids in calls)
        if len(devices) > 1:
            atoms = {d: sum(len(mols[i]) for c, ids in calls if c == d for i in ids) for d in devices}
            assert max(atoms.values()) - min(atoms.values()) <= max(len(m) for m in mols)    # LPT balance
        assert queries[0].stats["hits"] == table.count


def test_best_match_with_single_atom_templates_is_selected_after_the_first_candidate_drop(monkeypatch):
    """every single-atom tuple has RMSD 0, so an engine-side best_match keeps the first
    record of the run -- the candidate the reference never reports.  The host must drop that
    record FIRST and select afterwards (first of equal RMSDs in reference order)."""
    import numpy as np
    from pyjess_b200 import Jess, Template, synth
    tpl, mols, thr, _ = synth.workload("cfg3", scale=0.01)
    tpl = [Template(list(tpl[0])[:1], id="single"), tpl[1], Template(list(tpl[2])[:1], id="single2")]
    table = _random_table(len(mols), 3, 600, 9)
    table.rmsd[np.isin(table.template, (0, 2))] = 0.0
    seen = {}

    def spy(ctx, dev, batch, params, orig=None):
        seen["best"], seen["order"] = params.best_match, params.hit_order
        return orig(ctx, dev, batch, params)

    from pyjess_b200 import _capi
    _stub_engine(monkeypatch, mols, table)
    orig = _capi.query_records
    monkeypatch.setattr(_capi, "query_records", lambda *a: spy(*a, orig=orig))
    got = [[(jess_t, h.rmsd, h._row) for h in q for jess_t in [tpl.index(h.template())]]
           for q in Jess(tpl).query_many(mols, *thr, best_match=True, hit_order="canonical")]
    assert seen == {"best": 0, "order": 1}          # host-side selection, reference record order
    for m, hits in enumerate(got):
        want = []
        for t in (2, 1, 0):                          # reverse insertion order
            rows = [i for i in range(table.count) if table.molecule[i] == m and table.template[i] == t]
            if t != 1:
                rows = rows[1:]                      # the reference never reports its first candidate
            if rows:
                best = min(rows, key=lambda i: (table.rmsd[i], i))
                want.append((t, float(table.rmsd[best]), best))
        assert hits == want


def test_queries_of_one_batch_consumed_from_several_threads(monkeypatch):
    """`pool.map(list, jess.query_many(...))` -- the first thread runs the engine, the
    others must wait for it (not see a half-initialised run); an engine error is re-raised by
    every consumer, not replaced by a TypeError."""
    from multiprocessing.pool import ThreadPool
    import pytest
    from pyjess_b200 import Jess, _capi, synth
    tpl, mols, thr, _ = synth.workload("cfg3", scale=0.02)
    table = _random_table(len(mols), len(tpl), 500, 3)
    jess = Jess(tpl)
    _stub_engine(monkeypatch, mols, table, delay=0.05)
    with ThreadPool(8) as pool:
        counts = pool.map(lambda q: len(list(q)), jess.query_many(mols, *thr))
    assert sum(counts) == table.count

    def broken(*a):
        raise _capi.EngineError(-3, "out of memory")

    monkeypatch.setattr(_capi, "query_records", broken)
    queries = jess.query_many(mols, *thr)
    for q in queries[:3]:
        with pytest.raises(_capi.EngineError):
            next(q)


def test_single_structure_queries_from_a_thread_pool_combine_into_batches(monkeypatch):
    """The reference's recipe for many structures (README.md:147-162): `pool.map(lambda m: list(jess.query(m, ...)))`.
    Calls that find the engine busy are combined into ONE engine batch by the next free thread; every caller must
    still see exactly its structure's records, in order, and an engine error must reach every caller."""
    from multiprocessing.pool import ThreadPool
    import pytest
    from pyjess_b200 import Jess, _capi, synth
    tpl, mols, thr, _ = synth.workload("cfg3", scale=0.03)
    table = _random_table(len(mols), len(tpl), 900, 11)
    jess = Jess(tpl)
    _stub_engine(monkeypatch, mols, table)
    want = [[(h.rmsd, tpl.index(h.template()), h.atoms(transform=False)[0].serial) for h in jess.query(m, *thr)] for m in mols]
    assert sum(len(w) for w in want) == table.count
    calls = []
    _stub_engine(monkeypatch, mols, table, calls, delay=0.02)
    with ThreadPool(8) as pool:
        got = pool.map(lambda m: [(h.rmsd, tpl.index(h.template()), h.atoms(transform=False)[0].serial)
                                  for h in jess.query(m, *thr)], mols)
        # different parameters never share a batch
        mixed = pool.map(lambda im: len(list(jess.query(im[1], *thr, best_match=bool(im[0] % 2)))), list(enumerate(mols)))
    assert got == want
    served = [ids for _, ids in calls]
    assert any(len(ids) > 1 for ids in served), "no calls were combined"
    assert len(served) < 2 * len(mols)                                        # far fewer engine calls than the 2 x len(mols) queries
    assert sorted(i for ids in served for i in ids) == sorted(2 * list(range(len(mols))))   # every query served exactly once
    assert mixed == [len(w) for w in want]            # (the stub ignores best_match; the runs must not have been mixed up)

    def broken(*a):
        raise _capi.EngineError(-3, "out of memory")

    monkeypatch.setattr(_capi, "query_records", broken)
    with ThreadPool(4) as pool:
        def attempt(m):
            with pytest.raises(_capi.EngineError):
                next(jess.query(m, *thr))
            return True
        assert all(pool.map(attempt, mols[:12]))
    # the engine works again: nothing is left behind by the failed rounds
    _stub_engine(monkeypatch, mols, table)
    assert [(h.rmsd, tpl.index(h.template()), h.atoms(transform=False)[0].serial) for h in jess.query(mols[3], *thr)] == want[3]
    assert not jess._pending and not jess._leader


def test_query_properties_settable_only_while_they_can_be_honoured(monkeypatch):
    import pytest
    from pyjess_b200 import Jess, synth
    tpl, mols, thr, _ = synth.workload("cfg3", scale=0.01)
    table = _random_table(len(mols), len(tpl), 50, 1)
    _stub_engine(monkeypatch, mols, table)
    jess = Jess(tpl)
    q = jess.query(mols[0], *thr)
    q.ignore_chain = "atoms"                         # before the first next(): reaches the engine run
    assert q._run.args[3] == "atoms"
    list(q)
    q.ignore_chain = "atoms"                         # unchanged value: fine
    with pytest.raises(RuntimeError):
        q.ignore_chain = None
    with pytest.raises(RuntimeError):
        q.max_candidates = 3
    with pytest.raises(ValueError):
        q.ignore_chain = "all"
    many = jess.query_many(mols, *thr)
    with pytest.raises(RuntimeError):
        many[0].ignore_chain = "residues"            # one query of a shared batch
    with pytest.raises(ValueError):
        jess.query_many(mols, *thr, hit_order="random")


def test_templates_beyond_the_engine_limit_are_refused_at_construction():
    """the reference accepts templates of any size; the engine's pair record holds
    32 atoms (JB_MAX_TEMPLATE_ATOMS).  n = 32 is accepted, n = 33 raises when the Jess is built --
    with the template's index and size in the message -- never silently truncated."""
    import pytest
    from pyjess_b200 import Jess, Template, TemplateAtom
    atoms = [TemplateAtom(chain_id="A", residue_number=1 + k // 3, x=1.5 * k, y=0.3 * k, z=0.0, residue_names=["GLY"],
                          atom_names=["CA"], distance_weight=0.0, match_mode=0) for k in range(33)]
    assert len(Jess([Template(atoms[:32], id="ok")])) == 1
    with pytest.raises(ValueError, match="template 1 has 33 atoms.*at most 32"):
        Jess([Template(atoms[:5], id="small"), Template(atoms, id="big")])
