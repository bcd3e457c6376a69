"""The drop-in Python surface (`Jess.query` -> `Query` -> `Hit`) on the GPU, checked against the
known answers the reference's own test-suite pins for this path (src/pyjess/tests/test_jess.py,
test_hit.py -- cited per test).  Reads like the reference's tests on purpose: a PyJess user
switching packages runs the same assertions."""
import math
import os
import pickle

import numpy as np
import pytest

from _util import DATA, golden_cases, load_molecule, load_template
from pyjess_b200 import Jess, Molecule, Template

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lnb():
    return load_molecule("pdb1lnb.ent")


def test_query_thresholds_counts_and_scores(lnb):
    """test_jess.py:158-211: 1/2/3 hits, RMSD 0.555/1.440/1.644, det 1, logE -2.04/0.17/0.68."""
    template = load_template("template_01.qry")
    jess = Jess([template])
    expect = {(1, 2, 2): [(0.555, -2.04)], (2, 5, 3): [(0.555, -2.04), (1.440, 0.17)],
              (2, 5, 5): [(0.555, -2.04), (1.440, 0.17), (1.644, 0.68)]}
    for args, rows in expect.items():
        hits = sorted(jess.query(lnb, *args), key=lambda h: h.rmsd)
        assert len(hits) == len(rows)
        for hit, (rmsd, loge) in zip(hits, rows):
            assert hit.template() is template
            assert hit.rmsd == pytest.approx(rmsd, abs=5e-4)
            assert hit.determinant == pytest.approx(1.0, abs=1e-3)
            assert hit.log_evalue == pytest.approx(loge, abs=0.05)
            assert hit.evalue == pytest.approx(math.exp(loge), abs=0.05)
    hits = list(jess.query(lnb.conserved(10.0), 1, 2, 2))
    assert len(hits) == 1 and hits[0].log_evalue == pytest.approx(-2.10, abs=0.05)


def test_best_match_with_repeated_template(lnb):
    """test_jess.py:213-227: three copies of one template -> three hits, all the best RMSD."""
    template = load_template("template_01.qry")
    hits = list(Jess([template, template, template]).query(lnb, 2, 5, 5, best_match=True))
    assert len(hits) == 3
    for hit in hits:
        assert hit.template() is template
        assert hit.rmsd == pytest.approx(0.555, abs=5e-4)


def test_template_subclass_identity(lnb):
    """test_jess.py:118-133."""
    class MyTemplate(Template):
        pass
    template = MyTemplate.load(os.path.join(DATA, "template_01.qry"))
    jess = Jess([template])
    assert isinstance(jess[0], MyTemplate) and jess[0] == template
    hits = list(jess.query(lnb, 1, 2, 2))
    assert hits[0].template() is template


def _golden_blocks():
    with open(os.path.join(DATA, "1AMY+1.3.3.txt")) as f:
        return [b for b in f.read().split("REMARK") if b.strip()]


@pytest.mark.parametrize("reorder", [True, False])
def test_mcsa_query_matches_original_jess_output(reorder):
    """test_jess.py:229-335: 1.3.3.tpl x 1AMY at (2, 4, 4) against the 156 hit blocks written
    by the ORIGINAL Jess binary; hits are matched by their atom-serial tuples."""
    template = load_template("1.3.3.tpl")
    molecule = load_molecule("1AMY.pdb")
    blocks = {}
    for block in _golden_blocks():
        lines = block.strip().splitlines()
        blocks[tuple(int(line[6:11]) for line in lines[1:-1])] = lines
    hits = list(Jess([template]).query(molecule, 2, 4, 4, reorder=reorder))
    assert len(hits) == len(blocks) == 156
    seen = set()
    for hit in hits:
        serials = tuple(a.serial for a in hit.atoms(transform=False))
        lines = blocks[serials]
        seen.add(serials)
        query_id, rmsd, _tid, _, det, _, loge = lines[0].split()
        assert query_id == "1AMY"
        assert hit.rmsd == pytest.approx(float(rmsd), abs=5e-4)
        assert hit.determinant == pytest.approx(float(det), abs=0.05)
        assert hit.log_evalue == pytest.approx(float(loge), abs=0.05)
        for atom, line in zip(hit.atoms(), lines[1:-1]):
            assert atom.name == line[12:16].strip() and atom.residue_name == line[17:20].strip()
            assert atom.chain_id == line[20:22].strip() and atom.residue_number == int(line[22:26])
            assert (atom.x, atom.y, atom.z) == pytest.approx(
                (float(line[30:38]), float(line[38:46]), float(line[46:54])), abs=1.5e-3)
    assert len(seen) == 156


@pytest.mark.parametrize("mode,count", [(None, 0), ("residues", 2), ("atoms", 7)])
def test_ignore_chain_modes(mode, count):
    """test_jess.py:337-372."""
    template = load_template("4.1.2.tpl")
    molecule = load_molecule("5ayx.EF.pdb")
    hits = list(Jess([template]).query(molecule, 3, 3, 3, ignore_chain=mode))
    assert len(hits) == count
    if mode == "residues":
        for hit in hits:
            atoms = hit.atoms(transform=False)
            for a, b in zip(atoms, atoms[1:]):
                if a.residue_number == b.residue_number:
                    assert a.chain_id == b.chain_id


def test_ignore_chain_bool_is_deprecated(lnb):
    template = load_template("template_01.qry")
    with pytest.warns(DeprecationWarning):
        assert len(list(Jess([template]).query(lnb, 1, 2, 2, ignore_chain=True))) == 1
    with pytest.raises(ValueError):
        Jess([template]).query(lnb, 1, 2, 2, ignore_chain="all")


def test_hit_dump_text_and_pickle(lnb):
    """test_hit.py:27-57: exact PDB text of the first hit; pickle round trip."""
    template = Template.load(os.path.join(DATA, "template_01.qry"), id="T1")
    hit = next(Jess([template]).query(lnb, 1, 2, 2))
    expected = """REMARK 1LNB 0.555 T1 Det= 1.0 log(E)~ -2.04
ATOM   1112  ND1 HIS E 142       4.157  -2.574   4.091  1.00 10.22           N
ATOM   1115  NE2 HIS E 142       3.805  -2.316   1.890  1.00 16.00           N
ATOM   1123  OE1 GLU E 143       0.480  -4.544  -0.363  1.00 14.72           O
ATOM   1124  OE2 GLU E 143       1.870  -6.097   0.337  1.00 22.48           O
ATOM   1145  CG  HIS E 146       0.031   0.017   0.019  1.00 12.45           C
ATOM   1146  ND1 HIS E 146       0.823   1.160   0.000  1.00 15.79           N
ATOM   1147  CD2 HIS E 146       0.878  -1.064   0.006  1.00 20.85           C
ATOM   1148  CE1 HIS E 146       2.115   0.777  -0.013  1.00 13.50           C
ATOM   1149  NE2 HIS E 146       2.181  -0.553  -0.032  1.00 12.44           N
ATOM   1298  OE1 GLU E 166       6.432  -0.605   1.056  1.00 16.03           O
ATOM   1299  OE2 GLU E 166       4.839   0.070  -0.343  1.00 19.31           O
ENDMDL"""
    assert hit.dumps().strip() == expected
    copy = pickle.loads(pickle.dumps(hit))
    assert copy is not hit and copy.rmsd == hit.rmsd and copy.determinant == hit.determinant
    assert copy.evalue == hit.evalue
    assert copy.template(transform=False) == hit.template(transform=False)
    assert copy.atoms(transform=True) == hit.atoms(transform=True)
    assert copy.atoms(transform=False) == hit.atoms(transform=False)
    assert copy.molecule(transform=False) == hit.molecule(transform=False)


def test_hit_transformation_maps_atoms_onto_template(lnb):
    template = load_template("template_01.qry")
    hit = next(Jess([template]).query(lnb, 1, 2, 2))
    moved = np.array([[a.x, a.y, a.z] for a in hit.atoms(transform=True)])
    target = np.array([[a.x, a.y, a.z] for a in template])
    assert math.sqrt(((moved - target) ** 2).sum() / len(target)) == pytest.approx(hit.rmsd, abs=1e-9)
    m, inv = np.asarray(hit.transformation), np.asarray(hit.inverse_transformation)
    assert np.abs(m @ inv - np.eye(4)).max() < 1e-9
    back = hit.template(transform=True)
    orig = np.array([[a.x, a.y, a.z] for a in hit.atoms(transform=False)])
    assert math.sqrt(((np.array([[a.x, a.y, a.z] for a in back]) - orig) ** 2).sum() / len(orig)) == \
        pytest.approx(hit.rmsd, abs=1e-9)


def test_split_and_joint_template_sets_agree():
    """test_jess.py:96-116: Jess([t1, t2]) finds exactly the hits of Jess([t1]) and Jess([t2])."""
    t1, t2 = load_template("template_01.qry"), load_template("template_02.qry")
    molecule = load_molecule("1AMY.pdb")
    key = lambda h: (h.template().id, tuple(a.serial for a in h.atoms(False)))  # noqa: E731
    joint = list(Jess([t1, t2]).query(molecule, 2, 5, 5))
    split = list(Jess([t1]).query(molecule, 2, 5, 5)) + list(Jess([t2]).query(molecule, 2, 5, 5))
    assert sorted(map(key, joint)) == sorted(map(key, split))
    # templates come back in REVERSE insertion order (eff:Jess.c:109-117)
    order = [h.template() for h in joint]
    assert order == sorted(order, key=lambda t: 0 if t is t2 else 1)


def test_query_many_equals_single_queries():
    templates = [load_template(f"template_0{i}.qry") for i in (1, 2, 3)]
    from _util import cfg2_molecules
    mols = cfg2_molecules()[:12] + [Molecule()]
    jess = Jess(templates)
    batch = jess.query_many(mols, 2.0, 3.0, 3.0)
    assert len(batch) == len(mols)
    total = 0
    for mol, query in zip(mols, batch):
        assert query.molecule is mol
        one = list(jess.query(mol, 2.0, 3.0, 3.0))
        many = list(query)
        total += len(many)
        assert [(h.template().id, h.rmsd, tuple(a.serial for a in h.atoms(False))) for h in many] == \
               [(h.template().id, h.rmsd, tuple(a.serial for a in h.atoms(False))) for h in one]
    assert total > 0


def test_query_max_candidates(lnb):
    """The reference's own test (src/pyjess/tests/test_jess.py:135-156)."""
    jess = Jess([load_template("template_01.qry")])
    assert len(list(jess.query(lnb, 2, 5, 5))) == 3
    assert len(list(jess.query(lnb, 2, 5, 5, max_candidates=3))) == 3
    assert len(list(jess.query(lnb, 2, 5, 5, max_candidates=2))) == 2
    assert len(list(jess.query(lnb, 2, 5, 5, max_candidates=1))) == 1
    with pytest.raises(ValueError):
        list(jess.query(lnb, 2, 5, 5, max_candidates=-1))
    q = jess.query(lnb, 2, 5, 5)
    q.max_candidates = 2                 # settable before the first hit is pulled, like the reference's property
    assert q.max_candidates == 2 and len(list(q)) == 2


SINGLE = sorted(c for c in golden_cases() if golden_cases()[c].get("single_atom"))


@pytest.mark.parametrize("name", SINGLE)
def test_single_atom_templates_match_reference(name):
    """n == 1: the reference never reports the first entry of its candidate
    list and returns rmsd 0 with an all-zero rotation; golden hits come from the reference."""
    case = golden_cases()[name]
    templates = [load_template(t) for t in case["templates"]]
    mol = load_molecule(case["molecule"])
    hits = list(Jess(templates).query(mol, *case["params"], **case["kwargs"]))
    got = sorted((templates.index(h.template()), tuple(int(np.nonzero(mol._atoms["serial"] == a.serial)[0][0])
                                                       for a in h.atoms(False))) for h in hits)
    assert got == sorted((h["template"], tuple(h["atoms"])) for h in case["hits"])
    gold = {(h["template"], tuple(h["atoms"])): h for h in case["hits"]}
    for h in hits:
        g = gold[(templates.index(h.template()), tuple(int(np.nonzero(mol._atoms["serial"] == a.serial)[0][0])
                                                       for a in h.atoms(False)))]
        assert h.rmsd == pytest.approx(g["rmsd"], abs=1e-6)
        assert np.abs(np.asarray(h.transformation) - np.asarray(g["transformation"])).max() <= 1e-6
        assert h.determinant == pytest.approx(g["determinant"], abs=1e-6)
        assert h.log_evalue == pytest.approx(g["log_evalue"], abs=1e-9)


def test_thread_pool_recipe_of_the_reference_readme(lnb):
    """README.md:155-162 of the reference: one Jess, a thread pool over molecules.  Engine
    calls of concurrent threads are serialised per context; results must not mix."""
    from multiprocessing.pool import ThreadPool
    jess = Jess([load_template("template_01.qry"), load_template("1.3.3.tpl")])
    mols = [lnb, load_molecule("1AMY.pdb"), load_molecule("pdb1a0p.ent")] * 6
    expected = [[(h.template().id, round(h.rmsd, 9), tuple(a.serial for a in h.atoms(False)))
                 for h in jess.query(m, 2.0, 4.0, 4.0)] for m in mols[:3]]
    with ThreadPool(6) as pool:
        got = pool.map(lambda m: [(h.template().id, round(h.rmsd, 9), tuple(a.serial for a in h.atoms(False)))
                                  for h in jess.query(m, 2.0, 4.0, 4.0)], mols)
    assert got == expected * 6 and sum(map(len, expected)) > 100


def _all_hits_text(queries):
    import io
    buf = io.StringIO()
    for query in queries:
        for hit in query:
            hit.dump(buf)
            buf.write("\n")
    buf.seek(0)
    return buf


@pytest.mark.parametrize("all_instances", [True, False])
@pytest.mark.parametrize("workload", ["cfg2", "cfg3dense"])
def test_best_sites_equals_the_filter_script_on_the_dumped_hits(workload, all_instances):
    """`sites.best_sites` on the engine's hit table == `sites.filter_text` (pinned
    byte for byte to vendor/jess/filter_output.py by tests/test_sites.py) on the text of ALL
    hits.  cfg3dense: thousands of hits, several per site, self matches (a template renamed
    to the id of a structure)."""
    import io
    from pyjess_b200 import sites, synth
    if workload == "cfg2":
        from _util import cfg2_molecules
        templates = [Template.load(os.path.join(DATA, f"template_0{i}.qry"), id=f"template_0{i}") for i in (1, 2, 3)]
        mols, thr = cfg2_molecules(), (2.0, 3.0, 3.0)
    else:
        templates, mols, thr, _ = synth.workload("cfg3dense", scale=0.04)
        templates = [Template(list(t), id=mols[0].id if i == 1 else t.id) for i, t in enumerate(templates)]
    jess = Jess(templates)
    want = io.StringIO()
    n_want = sites.filter_text(_all_hits_text(jess.query_many(mols, *thr)), want, all_instances=all_instances)
    hits = sites.best_sites(jess.query_many(mols, *thr), all_instances=all_instances)
    got = io.StringIO()
    sites.write_hits(hits, got)
    assert len(hits) == n_want and n_want > 0
    assert got.getvalue() == want.getvalue()
    assert all(h.template().id != h.molecule().id for h in hits)


def _hit_keys(queries):
    return [[(h.template().id, tuple(a.serial for a in h.atoms(False)), h.rmsd) for h in q] for q in queries]


def test_device_coordinates_equal_the_host_transform():
    """`coordinates=True` returns, per hit, the matched atoms transformed into the
    template frame by the device (k_hit_coords: Hit.transformation + Atom._transform operation by
    operation).  They must equal the host transform of the same hit (1e-9: numpy's 4x4 product may
    contract differently), and `Hit.atoms` / `Hit.dump` must read them."""
    from pyjess_b200 import synth
    templates, mols, thr, _ = synth.workload("cfg3dense", scale=0.03)
    jess = Jess(templates)
    plain = jess.query_many(mols, *thr)
    rich = jess.query_many(mols, *thr, coordinates=True)
    n = 0
    for qp, qr in zip(plain, rich):
        table, rows = qr.hit_table()
        assert table.coords is not None and table.coords.shape == (table.count, table.stride, 3)
        hp, hr = list(qp), list(qr)
        assert len(hp) == len(hr) == len(rows)
        for a, b in zip(hp[:40], hr[:40]):
            assert a.rmsd == b.rmsd
            xa = np.asarray([[t.x, t.y, t.z] for t in a.atoms(transform=True)])
            xb = np.asarray([[t.x, t.y, t.z] for t in b.atoms(transform=True)])
            assert np.abs(xa - xb).max() <= 1e-9
            assert np.array_equal(xb, table.coords[b._row, :len(b.template())])
            assert a.dumps() == b.dumps()
            n += 1
        # padding columns of shorter templates are zero
        k = np.asarray([len(templates[t]) for t in table.template[rows]])
        for r, kk in zip(rows[:20], k[:20]):
            assert not table.coords[r, kk:].any()
    assert n > 100
    # best_match compacts the coordinates together with the other columns
    for qp, qr in zip(jess.query_many(mols[:10], *thr, best_match=True),
                      jess.query_many(mols[:10], *thr, best_match=True, coordinates=True)):
        for a, b in zip(list(qp), list(qr)):
            xa = np.asarray([[t.x, t.y, t.z] for t in a.atoms(transform=True)])
            xb = np.asarray([[t.x, t.y, t.z] for t in b.atoms(transform=True)])
            assert a.rmsd == b.rmsd and np.abs(xa - xb).max() <= 1e-9


def test_canonical_hit_order_is_the_same_set_in_tuple_order():
    from pyjess_b200 import synth
    templates, mols, thr, _ = synth.workload("cfg3dense", scale=0.03)
    jess = Jess(templates)
    ref = _hit_keys(jess.query_many(mols, *thr))
    can = _hit_keys(jess.query_many(mols, *thr, hit_order="canonical"))
    assert sum(map(len, ref)) > 200
    for a, b in zip(ref, can):
        assert sorted(a) == sorted(b)
        assert [k[0] for k in a] == [k[0] for k in b]          # same template runs, reverse insertion order
    assert any(a != b for a, b in zip(ref, can))               # ... and the order inside a run does differ


def test_records_call_equals_the_packed_call_sliced_or_not(monkeypatch):
    """`jb_query_records` (the engine packs the atom records itself, slice by slice behind the
    kernels) == `jb_query_host` on the batch packed by the caller, whatever the slicing."""
    from pyjess_b200 import _capi, synth
    from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params
    templates, mols, thr, kw = synth.workload("cfg3", scale=0.1)
    mols = mols + [Molecule()]
    ctx = _capi.default_context()
    dev = _capi.DeviceTemplates(ctx, TemplateTable(templates))
    for order in ("canonical", "reference"):
        qp = query_params(*thr, hit_order=order, **kw)
        want, _ = _capi.query_host(ctx, dev, MoleculeBatch(mols, serial=True), qp)
        assert want.count > 50
        for slices in ("1", "2", "5"):
            monkeypatch.setenv("JB_PIPELINE", slices)
            ctx2 = _capi.Context(0)                       # JB_PIPELINE is read when a context is created
            dev2 = _capi.DeviceTemplates(ctx2, TemplateTable(templates))
            for _ in range(2):                            # second call: staging and result blocks are re-used
                got, _ = _capi.query_records(ctx2, dev2, mols, qp)
                assert got.count == want.count
                for a, b in zip(got.arrays(), want.arrays()):
                    assert np.array_equal(a, b)
            del got
            dev2.close()
            ctx2.close()
    empty, _ = _capi.query_records(ctx, dev, [], query_params(*thr))
    assert empty.count == 0
    dev.close()


def test_query_many_over_several_devices_equals_one_device():
    """SURVEY 8e behind the drop-in API: structures dealt to the GPUs by atom count, one engine
    context and host thread per GPU, per-device hit tables (nothing merged)."""
    from pyjess_b200 import _capi, synth
    if _capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    templates, mols, thr, _ = synth.workload("cfg3", scale=0.2)
    jess = Jess(templates)
    one = _hit_keys(jess.query_many(mols, *thr, devices=[0]))
    for devices in ([0, 1], "all"):
        queries = jess.query_many(mols, *thr, devices=devices)
        assert _hit_keys(queries) == one
        assert len(queries[0]._run.tables) >= 2
    assert sum(map(len, one)) > 100
