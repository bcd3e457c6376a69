"""jb_molecules_pack on the host cores of the box: serial against JB_PACK_THREADS, and the
drop-in API call that uses it (`Jess.query_many` on cfg3: time to the first hit)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from pyjess_b200 import Jess, synth                    # noqa: E402
from pyjess_b200._jess import _staging_pool             # noqa: E402
from pyjess_b200._tables import MoleculeBatch           # noqa: E402

tpl, mols, thr, kw = synth.workload("cfg3")
out = {"cores": len(os.sched_getaffinity(0)), "atoms": int(sum(len(m) for m in mols))}
ref = None
for threads in ("1", "2", "4", "8", "16"):
    os.environ["JB_PACK_THREADS"] = threads
    best = 1e9
    for _ in range(5):
        t = time.perf_counter()
        b = MoleculeBatch(mols, serial=True, pool=_staging_pool())
        best = min(best, time.perf_counter() - t)
    cur = [b.atom_offset.copy()] + [np.asarray(getattr(b, k))[:b.n_atoms].copy() for k in ("x", "y", "z", "name", "resname", "resseq", "chain", "serial")]
    ref = ref or cur
    out[f"pack_ms_{threads}"] = round(best * 1e3, 3)
    out[f"same_{threads}"] = all(np.array_equal(u, v) for u, v in zip(ref, cur))
jess = Jess(tpl)
for threads in ("1", "8"):
    os.environ["JB_PACK_THREADS"] = threads
    best, hits = 1e9, 0
    for _ in range(4):
        t = time.perf_counter()
        qs = jess.query_many(mols, *thr)
        first = next(iter(qs[0]), None)
        best = min(best, time.perf_counter() - t)
        hits = sum(len(q.hit_table()[1]) for q in qs)
    out[f"first_hit_ms_{threads}"] = round(best * 1e3, 2)
    out[f"hits_{threads}"] = hits
print(json.dumps(out))
