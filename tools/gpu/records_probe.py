"""Wall time of the three host entry points on one workload, rep by rep: jb_query_host on a batch packed
into torch-pinned memory, the same on the library's own pinned staging, and jb_query_records.
    JB_DEBUG=2 python tools/gpu/records_probe.py [workload] [scale]"""
import sys
import time

sys.path.insert(0, ".")
from pyjess_b200 import _capi, synth                                          # noqa: E402
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params    # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tpl, mols, thr, kw = synth.workload(name, scale=scale)
ctx = _capi.default_context()
dev = _capi.DeviceTemplates(ctx, TemplateTable(tpl))
qp = query_params(*thr, **kw)
pinned = MoleculeBatch(mols, pinned=True)
for label, fn in (("query_host  torch-pinned batch", lambda: _capi.query_host(ctx, dev, pinned, qp)),
                  ("query_records             ", lambda: _capi.query_records(ctx, dev, mols, qp)),
                  ("query_host  torch-pinned batch", lambda: _capi.query_host(ctx, dev, pinned, qp))):
    times = []
    for rep in range(6):
        sys.stderr.write(f"--- {label} rep {rep}\n")
        t0 = time.perf_counter()
        table, st = fn()
        times.append((time.perf_counter() - t0) * 1e3)
        del table
    print(f"{name} x{scale} {label}: " + " ".join(f"{t:.2f}" for t in times) + f" ms; kernel_ms {st['kernel_ms']:.2f} retries {st['retries']}")
