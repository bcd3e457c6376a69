"""Hits on which the engine and the unmodified reference library disagree, with their RMSDs:
    python tools/gpu/diff_probe.py [workload] [scale]"""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import oracle                                                                 # noqa: E402
from pyjess_b200 import _capi, synth                                          # noqa: E402
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params    # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3dense"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tpl, mols, thr, kw = synth.workload(name, scale=scale)
ctx = _capi.default_context()
tt = TemplateTable(tpl)
dev = _capi.DeviceTemplates(ctx, tt)
qp = query_params(*thr, **kw)
got, _ = _capi.query_host(ctx, dev, MoleculeBatch(mols), qp)
want, _ = oracle.ref_query(tt, oracle.numpy_batch(mols), qp, threads=len(os.sched_getaffinity(0)))


def keys(t):
    return np.concatenate([t.molecule[:, None].astype(np.int64), t.template[:, None].astype(np.int64), t.atoms.astype(np.int64)], axis=1)


kg, kc = keys(got), keys(want)
vg = kg.view([("", kg.dtype)] * kg.shape[1]).reshape(-1)
vc = kc.view([("", kc.dtype)] * kc.shape[1]).reshape(-1)
only_c = np.nonzero(~np.isin(vc, vg))[0]
only_g = np.nonzero(~np.isin(vg, vc))[0]
print(f"{name} x{scale}: gpu {got.count} hits, reference {want.count}; only reference {len(only_c)}, only gpu {len(only_g)}")
for i in only_c[:10]:
    print("  only reference:", kc[i].tolist(), "rmsd", repr(float(want.rmsd[i])), "threshold", thr[0], "diff", float(want.rmsd[i]) - thr[0])
for i in only_g[:10]:
    print("  only gpu:", kg[i].tolist(), "rmsd", repr(float(got.rmsd[i])), "diff", float(got.rmsd[i]) - thr[0])
if len(only_c):
    # does the engine enumerate the tuple at all?  (every tuple of that pair, no RMSD filter)
    i = int(only_c[0])
    m, t = int(kc[i][0]), int(kc[i][1])
    dev1 = _capi.DeviceTemplates(ctx, TemplateTable([tpl[t]]))
    every, _ = _capi.query_host(ctx, dev1, MoleculeBatch([mols[m]]), query_params(*thr, all_tuples=True, **kw))
    n = len(tpl[t])
    hit = np.nonzero((every.atoms[:, :n] == kc[i][2:2 + n]).all(axis=1))[0]
    print("  engine enumerates the tuple:", len(hit) == 1, "rmsd there:", [repr(float(every.rmsd[j])) for j in hit])
    # which sweep loses it?  the pair alone and the whole structure, under every sweep implementation
    key = kc[i][2:2 + n]
    for label, env in (("depth-first", {"JB_DFS_PAIRS": "1000000"}), ("cell grid", {"JB_DFS_PAIRS": "0", "JB_SWEEP": "1"}),
                       ("x stream", {"JB_DFS_PAIRS": "0", "JB_SWEEP": "0"}), ("legacy", {"JB_DFS_PAIRS": "0", "JB_LEGACY": "1"})):
        os.environ.update(env)
        c2 = _capi.Context(0)
        for what, tl, ml in (("pair alone", [tpl[t]], [mols[m]]), ("all templates x the structure", tpl, [mols[m]]),
                             ("the template x all structures", [tpl[t]], mols)):
            d2 = _capi.DeviceTemplates(c2, TemplateTable(tl))
            tab, st = _capi.query_host(c2, d2, MoleculeBatch(ml), query_params(*thr, **kw))
            tt_i = 0 if len(tl) == 1 else t
            mm_i = 0 if len(ml) == 1 else m
            sel = (tab.template == tt_i) & (tab.molecule == mm_i)
            found = bool(((tab.atoms[sel][:, :n] == key).all(axis=1)).any()) if sel.any() else False
            print(f"  {label:12s} {what:32s}: {'found' if found else 'LOST '}  ({int(sel.sum())} hits of the pair, {tab.count} in all, retries {st['retries']})")
            d2.close()
        c2.close()
        for k in env:
            os.environ.pop(k)
