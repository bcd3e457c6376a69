#!/usr/bin/env python3
"""bench.py -- template x structure pairs/s of the query hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg3dense|cfg4shard|cfg5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU path (oracle/_ref)

One "step" = one pass of the hot path (candidate lists -> pair set-up -> frontier
sweep -> superposition/filter/compaction) over one batch of synthetic structures.
Workload at N=1: BASELINE config 3 (1,000 M-CSA-shaped templates x 1,000 synthetic
300-residue proteins = 1e6 pairs/step), the largest single-GPU configuration;
weak scaling: every rank processes its own 1,000 structures (structures are the
shard unit, templates replicated, no data-path collective).

Prints ONE JSON line (rank 0).  `value` = pairs/s with inputs resident in HBM,
`e2e` = the same through the C-ABI call with pinned HOST buffers (upload + kernels
+ hit download inside the timed region), `roofline` = the frontier sweep against the
measured HBM peak, `cpu_baseline` = the reference C library on this box's cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "template-structure pairs/sec"
UNIT = "pairs/s"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_traffic(workload):
    """DRAM bytes per step of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return int(t["k_expand_dram_bytes_per_step"]) if t.get("workload") == workload else None
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a
    thread every few milliseconds (the engine calls release the GIL); `nvidia-smi -lms` as a
    fallback when the NVML binding is unavailable (its first sample can arrive after a short
    timed region has ended, which is why NVML is preferred)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []
        self.nvml, self.handle, self.samples, self.stop_flag, self.thread = None, None, [], False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(device).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(device), "pci_bus_id") else None
            self.handle = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self.handle = h
                        break
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                reasons = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.samples.append((sm, int(reasons)))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.nvml is not None:
            n = self.nvml
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            seen = 0
            for _, r in self.samples:
                seen |= r
            reasons = sorted(k for k, b in bits.items() if seen & b)
            sm = [s for s, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                    "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def load_workload(name, scale=1.0, rank=0):
    """Synthetic BASELINE configs 3-5 (pyjess_b200/synth.py) or, for `cfg2`, the reference's own
    example set: 3 vendor templates x the 132 vendor structures at (2, 3, 3) -- 396 pairs, a
    LATENCY datapoint (the structures come from tests/golden/cfg2_structures.npz, written by the
    reference's parser; the same 132 structures on every rank)."""
    if name == "cfg2":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _util import cfg2_molecules, load_template
        tpl = [load_template(f"template_0{i}.qry") for i in (1, 2, 3)]
        return tpl, list(cfg2_molecules()), (2.0, 3.0, 3.0), {}
    from pyjess_b200 import synth
    return synth.workload(name, scale=scale, rank=rank)


def run_reference(args, rank, world):
    """--impl reference: the reference's own C implementation (oracle/_ref, built from
    the unmodified sources) on all host cores, on a bounded sample of the workload."""
    if rank != 0:
        return
    import oracle
    from pyjess_b200 import synth
    from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params
    tpl, mols, params, kw = load_workload(args.workload)
    sample = mols[:args.cpu_structures]
    tt, mb = TemplateTable(tpl), MoleculeBatch(sample)
    qp = query_params(*params, **kw)
    cores = host_cores()
    kind = "reference" if oracle.ref_available() else "port"
    fn = oracle.ref_query if kind == "reference" else oracle.query
    for _ in range(args.warmup):
        fn(tt, mb, qp, threads=cores)
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        hits, st = fn(tt, mb, qp, threads=cores)
        wall = time.perf_counter() - t0
        times.append(st["kernel_ms"] / 1e3 if kind == "reference" else wall)
    pairs = len(tpl) * len(sample)
    sec = float(np.mean(times))
    value = pairs / sec
    desc = f"{len(tpl)} templates x first {len(sample)} structures of {args.workload} per step, query loops only"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "templates": len(tpl), "structures": len(sample),
                   "thresholds": list(params)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hits_per_step": int(hits.count),
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; invalid as a result)")
    ap.add_argument("--cpu-structures", type=int, default=0,
                    help="structures in the bounded CPU-baseline sample (0: per-workload default, ~10-20 s of "
                         "host time at N=1; a third of it per step of --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.cpu_structures <= 0:
        full = {"cfg2": 132, "cfg3": 1000, "cfg3dense": 1000, "cfg4shard": 300, "cfg5": 60}.get(args.workload, 100)
        args.cpu_structures = full if args.impl == "b200" else max(1, full // 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pyjess_b200 import _build
    _build.build()                         # no-op unless the (git-ignored) library is missing or stale; lock-protected
    from pyjess_b200 import _capi, synth
    from pyjess_b200._tables import MoleculeBatch, TemplateTable, query_params

    # ---- inputs (synthetic, seeded; structures differ per rank, templates replicated);
    # generated before CUDA is initialised: the cfg 4 generator forks worker processes
    tpl, mols, params, kw = load_workload(args.workload, scale=args.scale, rank=rank)
    qp = query_params(*params, **kw)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()           # the engine's stream; torch events are recorded on it
    torch.cuda.set_stream(stream)
    ctx = _capi.Context(local, stream=stream.cuda_stream)
    tt = TemplateTable(tpl)
    dev_tpl = _capi.DeviceTemplates(ctx, tt)
    batch = MoleculeBatch(mols, pinned=True)
    dev_mol = _capi.DeviceMolecules(ctx, batch)
    pairs_rank = len(tpl) * len(mols)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input arm (kernels only) -------------------------------------------
    stats = None
    for _ in range(args.warmup):
        res = _capi.query_run(ctx, dev_tpl, dev_mol, qp)
        stats = res.stats()
        res.close()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    expand_ms, launches = [], 0
    t_wall = time.perf_counter()
    for e0, e1 in evs:
        flush.zero_()                      # L2 flush between timed iterations
        e0.record(stream)
        res = _capi.query_run(ctx, dev_tpl, dev_mol, qp)
        e1.record(stream)
        st = res.stats()
        res.close()
        expand_ms.append(st["expand_ms"])
        launches += st["launches"]
        stats = st
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    step_ms = float(np.sum([a.elapsed_time(b) for a, b in evs]))

    # ---- end-to-end arm: pinned host buffers -> hits on the host ----------------------
    for _ in range(2):
        table, _ = _capi.query_host(ctx, dev_tpl, batch, qp)
    barrier()
    ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    d2h = 0
    for e0, e1 in ee:
        flush.zero_()
        e0.record(stream)
        table, _ = _capi.query_host(ctx, dev_tpl, batch, qp)
        e1.record(stream)
        d2h = int(table.count) * (8 + 4 * table.stride + 8 * 16)
    barrier()
    e2e_ms = float(np.sum([a.elapsed_time(b) for a, b in ee]))

    # ---- reduce over ranks: MAX time, SUM pairs -----------------------------------------
    red = torch.tensor([step_ms, e2e_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(pairs_rank), float(stats["hits"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    step_ms, e2e_ms = float(red[0]), float(red[1])
    pairs_all, hits_all = float(tot[0]), int(tot[1])

    if rank == 0:
        value = pairs_all * args.steps / (step_ms / 1e3)
        e2e = pairs_all * args.steps / (e2e_ms / 1e3)
        # roofline of the dominant stage (frontier sweep): algorithmic bytes of SURVEY 8d
        peak, peak_src = measured_peak()
        alg_bytes = 32.0 * stats["candidate_atoms"] + 8.0 * stats["frontier_ints"]
        exp_ms = float(np.mean(expand_ms))
        achieved = alg_bytes / (exp_ms / 1e3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "templates": len(tpl), "structures_per_gpu": len(mols),
                       "atoms_per_gpu": int(batch.n_atoms), "pairs_per_step": int(pairs_all),
                       "thresholds": list(params), "sharding": f"structures x{world}, templates replicated",
                       "cache": "L2 flushed (512 MiB write) before every timed step", "scale": args.scale},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": int(batch.nbytes), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_expand (frontier sweep: one launch per level, summed over a step)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": measured_traffic(args.workload),
                         "algorithmic_bytes_per_step": alg_bytes,
                         "algorithmic_bytes": "32 B x sum_k|C_k| over live pairs + 8 B x sum_l l*F_l (SURVEY 8d)",
                         "kernel_ms_per_step": exp_ms, "launches_per_step": int(stats["launches"]),
                         "share_of_step": exp_ms / max(stats["kernel_ms"], 1e-9),
                         "note": "DRAM traffic is far below the algorithmic bytes (de-duplicated candidate lists stay "
                                 "in L2); the kernel is issue/L1-gather bound, see profiles/"},
            "stage_ms": {k: stats[k] for k in ("kernel_ms", "classify_ms", "pairs_ms", "expand_ms", "superpose_ms")},
            "work": {k: stats[k] for k in ("pairs", "live_pairs", "candidate_atoms", "candidate_records",
                                            "frontier_items", "frontier_ints", "candidate_tests", "tuples", "hits",
                                            "nan_dropped", "chunks", "retries")},
            "hits_per_step": hits_all,
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline:
            import oracle
            cores = host_cores()
            sample = mols[:max(1, int(args.cpu_structures * args.scale))]
            mb = MoleculeBatch(sample)
            kind = "reference" if oracle.ref_available() else "port"
            fn = oracle.ref_query if kind == "reference" else oracle.query
            t0 = time.perf_counter()
            cpu_hits, cst = fn(tt, mb, qp, threads=cores)
            wall = time.perf_counter() - t0
            sec = cst["kernel_ms"] / 1e3 if kind == "reference" else wall
            out["cpu_baseline"] = {
                "value": len(tpl) * len(sample) / sec, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": f"{len(tpl)} templates x first {len(sample)} structures of rank 0 "
                          f"({len(tpl) * len(sample)} pairs), query loops only, {sec:.2f} s"}
            # parity of the timed GPU result on that same sample (not timed)
            keep = table.molecule < len(sample)
            gpu_keys = sorted(zip(table.molecule[keep].tolist(), table.template[keep].tolist(),
                                  map(tuple, table.atoms[keep].tolist())))
            cpu_keys = sorted(zip(cpu_hits.molecule.tolist(), cpu_hits.template.tolist(),
                                  map(tuple, cpu_hits.atoms.tolist())))
            out["parity_on_sample"] = "identical hit set" if gpu_keys == cpu_keys else "MISMATCH"
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
