mkdir -p gpurun_out
cat > /tmp/ref_order_run.py <<'PY'
import sys
sys.path.insert(0, ".")
from pyjess_b200 import _capi, synth
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params
tpl, mols, thr, kw = synth.workload("cfg3")
ctx = _capi.default_context()
dev = _capi.DeviceTemplates(ctx, TemplateTable(tpl))
batch = MoleculeBatch(mols, pinned=True, serial=True)
qp = query_params(*thr, hit_order="reference", **kw)
for _ in range(3):
    table, st = _capi.query_host(ctx, dev, batch, qp)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1am_reforder.csv python /tmp/ref_order_run.py > gpurun_out/r1am.log 2>&1
python tools/summarise_launches.py gpurun_out/r1am_reforder.csv | sort -t'|' -k4 -n -r | head -14
