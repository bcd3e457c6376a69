"""Per-level counts and times of one resident engine run (JB_DEBUG=1 prints them on stderr).
    python tools/gpu/debug_levels.py [workload] [scale]"""
import os
import sys

os.environ["JB_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pyjess_b200 import _capi, synth  # noqa: E402
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tpl, mols, params, kw = synth.workload(name, scale=scale)
ctx = _capi.Context(0)
dt = _capi.DeviceTemplates(ctx, TemplateTable(tpl))
dm = _capi.DeviceMolecules(ctx, MoleculeBatch(mols))
qp = query_params(*params, **kw)
for i in range(3):
    sys.stderr.write(f"--- run {i}\n")
    res = _capi.query_run(ctx, dt, dm, qp)
    st = res.stats()
    res.close()
print({k: st[k] for k in ("kernel_ms", "classify_ms", "pairs_ms", "expand_ms", "superpose_ms", "scan_tests", "box_survivors",
                          "frontier_items", "candidate_tests", "tuples", "hits")})
