mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/r1ak_cfg2.json 2>gpurun_out/r1ak_cfg2.err; python -c "
import json; b=json.loads(open('gpurun_out/r1ak_cfg2.json').read().strip().splitlines()[-1]); print('cfg2', b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'], b['hits_per_step'], b['cpu_baseline'], b['parity_on_sample'], b['gpu_launches'])"
python bench.py --workload cfg2 --impl reference --steps 5 --warmup 1 | cut -c1-300
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from pyjess_b200 import Jess, Molecule, Template
tpl = [Template.load(f'tests/data/template_0{i}.qry') for i in (1,2,3)]
mol = Molecule.load('tests/data/pdb1a0p.ent')
jess = Jess(tpl)
list(jess.query(mol, 2.0, 3.0, 3.0))
t=time.perf_counter()
for _ in range(50): n=len(list(jess.query(mol, 2.0, 3.0, 3.0)))
print('README example: Jess.query latency', (time.perf_counter()-t)/50*1e3, 'ms,', n, 'hit')
t=time.perf_counter()
for _ in range(20): Molecule.load('tests/data/pdb1a0p.ent')
print('Molecule.load', (time.perf_counter()-t)/20*1e3, 'ms')
PY
