"""Under torchrun (one process per GPU): BASELINE config 2 (3 vendor templates x the 132 vendor
structures) and a cfg3 sample, sharded by structure over the ranks with
`pyjess_b200.sharding.query_sharded`; rank 0 checks that the merged table equals its own
single-GPU table of the whole batch ("identical merged hit set")."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import cfg2_molecules, load_template                                    # noqa: E402
from pyjess_b200 import _capi, sharding, synth                                     # noqa: E402
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params         # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = _capi.Context(local)
cases = [("cfg2", [load_template(f"template_0{i}.qry") for i in (1, 2, 3)], list(cfg2_molecules()), (2.0, 3.0, 3.0), {})]
tpl, mols, thr, kw = synth.workload("cfg3", scale=0.2)
cases.append(("cfg3 x0.2", tpl, mols, thr, kw))
for name, tpl, mols, thr, kw in cases:
    dev = _capi.DeviceTemplates(ctx, TemplateTable(tpl))
    params = query_params(*thr, **kw)
    run = lambda ms: _capi.query_host(ctx, dev, MoleculeBatch(ms), params)[0]      # noqa: E731
    merged = sharding.query_sharded(run, mols, rank, world)
    if rank == 0:
        merged = merged.merged()
        whole = run(mols)
        same = merged.count == whole.count and all(np.array_equal(a, b) for a, b in zip(merged.arrays(), whole.arrays()))
        print(f"{name}: {world} ranks, merged {merged.count} hits, single GPU {whole.count} hits, identical={same}", flush=True)
        assert same
    dev.close()
dist.barrier()
dist.destroy_process_group()
