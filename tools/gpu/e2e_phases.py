"""Wall-clock split of the end-to-end call (host buffers -> hits on the host) on cfg3."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pyjess_b200 import synth, _capi
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
tpl, mols, thr, kw = synth.workload(name)
params = query_params(*thr, **kw)
ctx = _capi.default_context()
tt = TemplateTable(tpl)
dev = _capi.DeviceTemplates(ctx, tt)
t0 = time.perf_counter(); mb = MoleculeBatch(mols, pinned=True); t1 = time.perf_counter()
print(f"pack MoleculeBatch (python): {1e3*(t1-t0):.1f} ms, {mb.nbytes/1e6:.1f} MB")
for it in range(4):
    t0 = time.perf_counter()
    dm = _capi.DeviceMolecules(ctx, mb); ctx.synchronize()
    t1 = time.perf_counter()
    res = _capi.query_run(ctx, dev, dm, params); ctx.synchronize()
    t2 = time.perf_counter()
    table = res.fetch()
    t3 = time.perf_counter()
    st = res.stats()
    dm.close()
    t4 = time.perf_counter()
    tab2, st2 = _capi.query_host(ctx, dev, mb, params)
    t5 = time.perf_counter()
    print(f"iter {it}: upload {1e3*(t1-t0):.2f}  run {1e3*(t2-t1):.2f} (kernel_ms {st['kernel_ms']:.2f})  fetch {1e3*(t3-t2):.2f}  "
          f"hits {table.count}  | query_host total {1e3*(t5-t4):.2f}")
