"""Host logic of the multi-GPU path on CPU: LPT partition, merge, and a world_size-2
`gloo` run in which every rank scans its shard (with the CPU oracle standing in for the
engine -- tests may use it) and rank 0 must reproduce the single-process hit table."""
import os
import socket
import sys

import numpy as np
import pytest

from _util import cfg2_molecules, load_template
from pyjess_b200 import sharding
from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    counts = rng.integers(1000, 12000, size=257).tolist() + [58870]
    for world in (1, 2, 4, 8):
        shards = sharding.partition(counts, world)
        allidx = np.sort(np.concatenate(shards))
        assert allidx.tolist() == list(range(len(counts)))
        loads = [int(np.sum(np.asarray(counts)[s])) for s in shards]
        assert max(loads) - min(loads) <= max(counts)
        for s in shards:
            assert np.all(np.diff(s) > 0)


def test_partition_edge_cases():
    assert [s.tolist() for s in sharding.partition([], 3)] == [[], [], []]
    assert [s.tolist() for s in sharding.partition([5], 2)] == [[0], []]
    with pytest.raises(ValueError):
        sharding.partition([1], 0)


def _scan(templates, mols):
    import oracle
    table, _ = oracle.query(TemplateTable(templates), MoleculeBatch(mols), query_params(2.0, 3.0, 3.0), threads=2)
    return table


def _same(a, b):
    assert a.keys() == b.keys()
    for x, y in zip(a.arrays(), b.arrays()):
        assert np.array_equal(x, y)


def test_merge_reproduces_single_shard(oracle_lib):
    templates = [load_template(f"template_0{i}.qry") for i in (1, 2, 3)]
    mols = cfg2_molecules()[:24]
    whole = _scan(templates, mols)
    assert whole.count > 0
    for world in (2, 3):
        shards = sharding.partition([len(m) for m in mols], world)
        tables = [_scan(templates, [mols[int(i)] for i in s]) if len(s) else None for s in shards]
        _same(sharding.merge(tables, shards, stride=whole.stride), whole)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here, os.path.dirname(here)]
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        templates = [load_template(f"template_0{i}.qry") for i in (1, 2, 3)]
        mols = cfg2_molecules()[:16]
        merged = sharding.query_sharded(lambda ms: _scan(templates, ms), mols, rank, world)
        if rank == 0:
            whole = _scan(templates, mols)
            assert merged.count == whole.count and merged.checksum() == sharding.ShardedHits([whole], [np.arange(len(mols))]).checksum()
            for m in range(len(mols)):                       # the per-structure index points at the right records
                t, lo, hi = merged.of(m)
                assert hi - lo == int((whole.molecule == m).sum())
                assert np.array_equal(t.rmsd[lo:hi], whole.rmsd[whole.molecule == m])
            _same(merged.merged(), whole)
            out.put(("ok", merged.count))
        else:
            assert merged is None
    except Exception as exc:      # surface the failure in the parent
        out.put(("fail", repr(exc)))
        raise
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_query(oracle_lib):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    status, count = out.get(timeout=5)
    assert status == "ok" and count > 0
