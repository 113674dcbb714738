#!/usr/bin/env python
"""bench.py -- stories/sec of the 3-hop quantized MemN2N inference forward (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C1|C3|C4|C5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic bAbI-shaped stories.  At N=1 the
workload is BASELINE.json configs[1]: all-tasks-joint shape, V=256 (192 words + 64 time columns),
d=50, 50 memory slots, 3 hops, fixed-point dot-product attention, 20 000 stories (1.04 GB of dense
fp32 bag-of-words, resident in HBM and larger than the 126 MB L2, so every step streams it from
HBM).  With N ranks every rank processes its own 20 000-story shard with no data-path collective
(stories are independent): weak scaling, value = stories of all ranks / max-over-ranks time.

The JSON line carries: value (device-resident), e2e (host arenas in, predictions out through
qmann_infer_host, H2D/D2H inside the timed region), roofline of the dominant kernel (algorithmic
bytes per launch / its CUDA-event duration, against the measured HBM copy bandwidth), cpu_baseline
(the CPU restatement of the reference on the box's host cores), clocks sampled during the run and
the number of kernels this library launched in the timed region.

--workload C5 is the other multi-GPU shape (BASELINE.json configs[4]): ONE pre-embedded memory of 2^20 slots,
d=256, 3 hops, slot-sharded across the ranks (strong scaling: the memory is fixed, every rank streams its
S/N slots), Q queries per step, two integer all-reduces per hop over NCCL.

--impl reference: the reference has no runnable CPU forward, so this arm times
the CPU restatement of its CUDA arithmetic (oracle/, kind "port") on all host threads (the thread count is passed
explicitly: torchrun exports OMP_NUM_THREADS=1), on a bounded sample of the same workload.
--impl reference_gpu: the reference's OWN CUDA path (lib/layer.c + lib/layer_cuda.cu compiled unmodified for
sm_100a into oracle/_ref/ref_harness_refcuda, 31 launches per story) timed on this box's GPU on a bounded sample.

The default (C1-C4) line also carries: `roofline` for the WHOLE step (algorithmic bytes of the step / step time) with the
kernels broken out, `path_share` (which kernel tier finished how many stories), `sensitivity` (the same step with
N(0,1) and N(0,2) weights, whose codes saturate more and leave the packed tier), `e2e_ids` (host word-id lists in,
predictions out), `reference_gpu` (N=1) and `c5`: the slot-sharded large-memory workload (the one with a collective)
measured in the same process group at Q = 64 and 1024, sharded == unsharded checked before timing.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

METRIC = "stories/sec for 3-hop quantized MemN2N at 1/2/4/8 B200; % of HBM roofline"
UNIT = "stories/s"
SIGMA = 0.5          # weight scale: N(0, 0.5) stands in for trained weights (N(0,0.1) quantises to all-zero codes)
NOMINAL_HBM_GBS = 8000.0


def workload(name: str, synth):
    cfg = synth.preset_config(name)
    n = {"C1": 1000, "C2": 20000, "C3": 20000, "C4": 65536}[name]
    S = 50
    desc = {
        "C1": "bAbI task-1 shape: V=70 (20 words + 50 time cols), d=20, 50 slots, 3 hops, fixed-point dot attention",
        "C2": "all-20-tasks joint shape: V=256 (192 words + 64 time cols), d=50, 50 slots, 3 hops, fixed-point dot attention",
        "C3": "joint shape, Hamming/approximate attention on 8-bit keys: V=256, d=50, 50 slots, 3 hops",
        "C4": "batch-sweep shape: V=114 (64 words + 50 time cols), d=64, 50 slots, 3 hops, fixed-point dot attention",
    }[name]
    return cfg, n, S, desc


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed ncu --set full capture
    (profiles/r01_traffic.json; the launch it was taken on is named there).  None when there is no capture."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                v = json.load(fh).get(kernel)
            if v is not None:
                return v
        except OSError:
            continue
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nme, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
    return rank, world, local


def host_threads():
    """Host threads this process may use.  Passed explicitly to the oracle: torch.distributed.run exports
    OMP_NUM_THREADS=1, which would silently make the CPU arm single-threaded (round-1 SCALE bug)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_oracle_rate(synth, cfg, w, S, sample, threads=0):
    """stories/s of the CPU restatement (oracle/) on `sample` stories of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qmo
    st = synth.make_stories(cfg, sample, 0x5EED1000 + 99, S=S)
    qmo.lib()
    if threads <= 0:
        threads = host_threads()
    t0 = time.perf_counter()
    out = qmo.forward(cfg, w, st, dump=False, n_threads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, dt, threads, out


def reference_gpu_rate(synth, cfg, w, S, sample=2000, reps=2):
    """stories/s of the reference's own CUDA path (oracle/_ref/ref_harness_refcuda: the unmodified lib/layer.c +
    lib/layer_cuda.cu, one story per 31 launches) on `sample` stories of the workload.  None when it was not built."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness_refcuda")
    if not os.path.exists(exe):
        return None
    st = synth.make_stories(cfg, sample, 0x5EED1000 + 98, S=S)
    with tempfile.TemporaryDirectory() as td:
        case, dump = os.path.join(td, "c.bin"), os.path.join(td, "d.bin")
        synth.write_case(case, cfg, w, st)
        try:
            r = subprocess.run([exe, case, dump, str(reps)], check=True, capture_output=True, timeout=600, text=True)
        except Exception as e:          # noqa: BLE001
            return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    for line in r.stdout.splitlines():
        if line.startswith("TIME_S"):
            sec = float(line.split()[1])
            return {"value": sample / sec, "unit": UNIT, "sample": f"{sample} stories, {reps} timed passes after one dumping pass",
                    "s_per_pass": sec, "kind": "reference CUDA path rebuilt for sm_100a (oracle/Makefile), unmodified"}
    return {"unavailable": "no TIME_S line"}


# ---------------------------------------------------------------------------------------------------
# C5: one very large memory, slot-sharded
# ---------------------------------------------------------------------------------------------------
C5_S, C5_D, C5_V = 1 << 20, 256, 64


C5_MODE = 2          # --c5-attention hamming: 3 (the reference's approximate Hamming scorer; per-element kernel k_big_scores<3>)


def c5_config(synth, mode=None):
    mode = C5_MODE if mode is None else mode
    if mode == 3:
        return synth.ModelConfig(V=C5_V, d=C5_D, S_max=64, V_dict=32, mode=3, iwl=3)
    return synth.ModelConfig(V=C5_V, d=C5_D, S_max=64, V_dict=32, mode=mode)


def c5_cpu_rate(synth, cfg, weights, slots=1 << 16, queries=4):
    """queries/s of the CPU restatement on the FULL memory, extrapolated from the scorer + read over `slots`
    slots (the cost is linear in the slot count; the per-query tail is included in the timed region)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qmo_bigmem as qb
    rng = np.random.default_rng(0x5EED1005)
    f = cfg.formats()
    M8 = np.stack([np.clip(np.rint(rng.standard_normal((slots, cfg.d)) * 0.1 * (1 << f["frac_w"][h])), -127, 127).astype(np.int8) for h in range(cfg.H)])
    C8 = np.stack([np.clip(np.rint(rng.standard_normal((slots, cfg.d)) * 0.5 * (1 << f["frac_w"][h])), -127, 127).astype(np.int8) for h in range(cfg.H)])
    u0 = np.clip(np.rint(rng.standard_normal((queries, cfg.d)) * 3), -127, 127).astype(np.int8)
    t0 = time.perf_counter()
    qb.forward(cfg, weights, M8, C8, u0)
    dt = time.perf_counter() - t0
    per_query_full = dt / queries * (C5_S / slots)
    return 1.0 / per_query_full, dt, slots, queries


def c5_build(torch, synth, qlib, world, rank, dev, Q, group):
    """This rank's slot shard of the synthetic 2^20 x 256 memory (generated on the device, shard by shard, from seeds that
    depend only on (hop, slot block), so the global memory is the same for every world size), the queries and the object."""
    cfg = c5_config(synth)
    weights = synth.make_weights(cfg, 0x5EED0000 + 5, sigma=0.3)
    f = cfg.formats()
    lo, n_loc = qlib.slot_shard(C5_S, world, rank)
    BLK = 1 << 16

    def gen(kind, h):
        out = torch.empty((n_loc, cfg.d), dtype=torch.int8, device=dev)
        sc = (0.1 if kind == 0 else 0.5) * (1 << f["frac_w"][h])
        for b0 in range(lo // BLK * BLK, lo + n_loc, BLK):
            g = torch.Generator(device=dev).manual_seed(0x5EED1000 + 7919 * (b0 // BLK) + 131 * h + kind)
            blk = (torch.randn((BLK, cfg.d), device=dev, generator=g) * sc).round().clamp(-127, 127).to(torch.int8)
            a, b = max(b0, lo), min(b0 + BLK, lo + n_loc)
            out[a - lo:b - lo] = blk[a - b0:b - b0]
        return out
    M8 = [gen(0, h) for h in range(cfg.H)]
    C8 = [gen(1, h) for h in range(cfg.H)]
    gq = torch.Generator(device=dev).manual_seed(0x5EED2000)
    u0 = (torch.randn((Q, cfg.d), device=dev, generator=gq) * 3).round().clamp(-127, 127).to(torch.int8)
    # plant one strongly matching slot per query in hop 0 so that the weighted read is not empty
    for q in range(Q):
        r = (q * 104729 + 17) % C5_S
        if lo <= r < lo + n_loc:
            M8[0][r - lo] = (u0[q].to(torch.int32) * 2).clamp(-127, 127).to(torch.int8)
    mem = qlib.BigMemory(cfg, weights, M8, C8, C5_S, lo, Q_max=Q, device=dev, group=group, world=world)
    return cfg, weights, mem, u0, n_loc


def run_bigmem(args):
    import torch
    rank, world, local = dist_setup(args.gpus)
    if world == 1 and args.gpus > 1:
        print("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    pkg = ge.import_package()
    synth, qlib = pkg.synth, pkg.lib
    Qs = [int(x) for x in str(args.queries).split(",") if x]
    Q = max(Qs)
    W, K = max(3, args.warmup), max(1, args.steps)
    group = None
    if world > 1:
        import torch.distributed as dist
        group = dist.group.WORLD
    cfg, weights, mem, u0_all, n_loc = c5_build(torch, synth, qlib, world, rank, dev, Q, group)
    rc = 0
    for qi, Qn in enumerate(Qs):
        rc |= _bigmem_measure(args, torch, qlib, synth, cfg, weights, mem, u0_all[:Qn].contiguous(), Qn, W, K, rank, world, local, dev, n_loc,
                              cpu_wanted=(qi == 0))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return rc


def c5_record(torch, synth, qlib, rank, world, local, Qs=(64, 1024), K=10, W=3):
    """The slot-sharded large-memory workload (BASELINE configs[4], the one with a collective) measured inside the default
    run so that the driver's per-N lines carry it: strong scaling (one 2^20 x 256 memory split over the ranks), queries/s at
    Q = 64 and 1024, the scorer's share and the cost of the two per-hop all-reduces; at N > 1 the sharded predictions are
    checked against an unsharded run of the same memory on rank 0 before anything is timed."""
    dev = f"cuda:{local}"
    group = None
    if world > 1:
        import torch.distributed as dist
        group = dist.group.WORLD
    Qmax = max(Qs)
    cfg, weights, mem, u0, n_loc = c5_build(torch, synth, qlib, world, rank, dev, Qmax, group)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    comm_holder = [None]
    rec = {"workload": "C5: one pre-embedded memory of 2^20 slots, d=256, 3 hops, fixed-point dot attention, slot-sharded x%d "
                       "(strong scaling); per hop all_reduce(SUM) of u32 score histograms [Q][255] and i32 partial reads [Q][256]" % world,
           "slots_per_rank": int(n_loc), "scaling": "strong"}
    if world > 1:
        import torch.distributed as dist
        pred_sh = mem.forward(u0)["pred"].cpu().numpy().copy()
        ok = torch.ones(1, dtype=torch.int32, device=dev)
        if rank == 0:
            _, _, full, u0f, _ = c5_build(torch, synth, qlib, 1, 0, dev, Qmax, None)
            pred_full = full.forward(u0f)["pred"].cpu().numpy().copy()
            full.close()
            del full
            ok[0] = int(np.array_equal(pred_sh, pred_full))
        dist.broadcast(ok, src=0)
        rec["sharded_equals_unsharded"] = bool(int(ok[0]))
        assert rec["sharded_equals_unsharded"], "slot-sharded predictions differ from the unsharded memory"
        torch.cuda.empty_cache()
    for Q in Qs:
        uq = u0[:Q].contiguous()
        for _ in range(W):
            mem.forward(uq)
        barrier()
        mem.profile(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(K):
            mem.forward(uq)
            mem.profile_read()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1) / K
        ms_scores, n_scores = mem.profile_read(reset=True)
        mem.profile(False)
        # the two per-hop exchanges alone, same payloads, same stream
        ar_ms = 0.0
        if world > 1:
            import torch.distributed as dist
            hist, part = mem.hist[:Q], mem.partial[:Q]
            for _ in range(3):
                dist.all_reduce(hist); dist.all_reduce(part)
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(K * cfg.H):
                dist.all_reduce(hist); dist.all_reduce(part)
            a1.record(stream)
            barrier()
            ar_ms = a0.elapsed_time(a1) / K
            hist.zero_(); part.zero_()
        # the same forward through the one-call C entry (qmann_bigmem_forward_sharded): library-side ncclAllReduce, one CUDA graph
        graph_ms = 0.0
        try:
            if comm_holder[0] is None and world > 1:
                comm_holder[0] = qlib.NcclComm()
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                for _ in range(W):
                    pg = mem.forward_sharded(uq, comm_holder[0])
                side.synchronize()
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record(side)
                for _ in range(K):
                    pg = mem.forward_sharded(uq, comm_holder[0])
                g1.record(side)
                side.synchronize()
            barrier()
            graph_ms = g0.elapsed_time(g1) / K
            pg = pg.clone()
            assert torch.equal(pg, mem.forward(uq)["pred"]), "one-call sharded forward differs from the phase API"
        except AssertionError:
            raise
        except Exception as e:          # noqa: BLE001
            graph_ms = 0.0
            rec.setdefault("graph_entry_error", f"{type(e).__name__}: {e}"[:200])
        t = torch.tensor([ms, ms_scores / max(1, n_scores), ar_ms, graph_ms], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, sc_ms, ar_ms, graph_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        peak, _ = measured_peak()
        rec[f"q{Q}"] = {"queries_per_s": Q / (ms / 1e3), "ms_per_step": ms,
                        "one_call_graph": ({"queries_per_s": Q / (graph_ms / 1e3), "ms_per_step": graph_ms,
                                            "api": "qmann_bigmem_forward_sharded (ncclAllReduce inside the library, one CUDA graph per forward)"} if graph_ms > 0 else None),
                        "scorer_ms_per_launch": sc_ms,
                        "scorer_launches_per_step": n_scores / K, "allreduce_ms_per_step": ar_ms,
                        "scorer_GB/s": n_loc * cfg.d / (sc_ms / 1e3) / 1e9 if sc_ms > 0 else None,
                        "scorer_frac_of_hbm": (n_loc * cfg.d / (sc_ms / 1e3) / 1e9 / peak) if sc_ms > 0 else None}
    if comm_holder[0] is not None:
        comm_holder[0].close()
    mem.close()
    del mem
    torch.cuda.empty_cache()
    return rec


def _bigmem_measure(args, torch, qlib, synth, cfg, weights, mem, u0, Q, W, K, rank, world, local, dev, n_loc, cpu_wanted):
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        mem.forward(u0)
    barrier()
    mem.profile(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = qlib.lib().qmann_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        out = mem.forward(u0)
        mem.profile_read()              # folds this step's event pairs (waits for its score kernels only)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = qlib.lib().qmann_launch_count() - launches0
    ms_scores, n_scores = mem.profile_read(reset=True)
    mem.profile(False)
    # keep the GPUs busy a little longer for the 100 ms clock sampler.  The forward contains collectives, so every rank
    # must run the SAME number of extra passes: derive it from the max-over-ranks step time, never from a local clock.
    ms_sync = ms_total
    if world > 1:
        import torch.distributed as dist
        t_ = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        ms_sync = float(t_[0])
    n_extra = int(min(2000, max(0.0, 600.0 - ms_sync) / max(ms_sync / K, 1e-3)))
    for _ in range(n_extra):
        mem.forward(u0)
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    pred_dev = out["pred"].cpu().numpy().copy()

    # end to end: pinned host queries in, host predictions out
    u0_pin = u0.cpu().pin_memory()
    pred_pin = torch.empty(Q, dtype=torch.int32).pin_memory()
    u0_stage = torch.empty_like(u0)
    Ke = args.e2e_steps or min(K, 8)

    def e2e_step():
        u0_stage.copy_(u0_pin, non_blocking=True)
        o = mem.forward(u0_stage)
        pred_pin.copy_(o["pred"], non_blocking=True)
        torch.cuda.synchronize()
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / Ke
    assert np.array_equal(pred_pin.numpy(), pred_dev)

    ms_step = ms_total / K
    k_scores_ms = ms_scores / max(1, n_scores)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_step, e2e_ms, k_scores_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, k_scores_ms = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        value = Q / (ms_step / 1e3)
        peak, peak_src = measured_peak()
        bytes_launch = n_loc * cfg.d                     # one hop's M shard, int8, read once per launch
        achieved = bytes_launch / (k_scores_ms / 1e3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline and cpu_wanted:
            rate, dt, slots, qs = c5_cpu_rate(synth, cfg, weights)
            cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{qs} queries x {slots} of the 2^20 slots in {dt:.1f} s (oracle/qmo_bigmem.py over qmann_oracle.c), "
                             "scaled linearly in the slot count to the full memory"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": f"C5: one pre-embedded memory of 2^20 slots, d=256, 3 hops, {'Hamming (approximate)' if C5_MODE == 3 else 'fixed-point dot'} attention, {Q} queries per step "
                                   "(a story = one query against the whole memory)",
                       "memory": "int8 codes M_h, C_h per hop (1.61 GB), N(0,0.1)/N(0,0.5) quantised in the EN_MQ formats, resident in HBM",
                       "l2": f"{n_loc * cfg.d / 1e6:.0f} MB of M per hop and rank; three different hops' memories are streamed per step "
                             f"({3 * n_loc * cfg.d / 1e6:.0f} MB" + (", larger than the 126 MB L2)" if 3 * n_loc * cfg.d > 126e6 else ", smaller than the 126 MB L2: query blocks after the first hit L2)"),
                       "parallelism": f"slot-sharded x{world}; per hop all_reduce(SUM) of u32 score histograms [Q][255] and i32 partial reads [Q][256] over NCCL"},
            "e2e": {"value": Q / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(Q * cfg.d), "d2h_bytes_per_step": int(4 * Q),
                    "ms_per_step": e2e_ms, "steps": Ke, "api": "qmann_bigmem_* phases (pinned host queries in, host predictions out)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ((ncu_traffic("k_big_scores_fast") or {}).get("dram_bytes_per_launch") if (world == 1 and Q == 1) else
                                     (ncu_traffic("k_big_scores_tq") or {}).get("dram_bytes_per_launch") if (world == 1 and Q == 64) else None),
                         "kernel": "k_big_scores_ham (packed bytes)" if C5_MODE == 3 else "k_big_scores_fast" if Q < 4 else ("k_big_scores_mma" if os.environ.get("QMANN_BIGMEM_TC") == "0" else
                                                                      ("k_big_scores_tc" if os.environ.get("QMANN_BIGMEM_TQ") == "0" else "k_big_scores_tq")),
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / NOMINAL_HBM_GBS,
                         "algorithmic_bytes_per_launch": bytes_launch, "launch_ms": k_scores_ms,
                         "note": "one launch streams one hop's M shard (S_local*d int8) once per block of up to 128 queries (Q > 128: one pass "
                                 "per block); the C rows are a sparse gather of the <= 2^frac slots whose quantised attention weight is non-zero, so "
                                 "they are not streamed.  Q < 4: k_big_scores_fast, HBM-bound.  Q >= 4: k_big_scores_tq, the truncated products "
                                 "as four int8 contractions on tcgen05.mma kind::i8 (query planes as the A operand in tensor memory, TMA-staged "
                                 "memory tiles as B, accumulators in tensor memory; 4*S*d*Q multiply-adds per hop), bound by shared-memory "
                                 "bandwidth (plane construction + operand reads), not by HBM"},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    return 0


def run_reference(args):
    """--impl reference: CPU restatement of the reference (kind "port"), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload == "C5":
        pkg = ge.import_package()
        cfg = c5_config(pkg.synth)
        weights = pkg.synth.make_weights(cfg, 0x5EED0000 + 5, sigma=0.3)
        rates = []
        for i in range(args.warmup + args.steps):
            rate, dt, slots, qs = c5_cpu_rate(pkg.synth, cfg, weights)
            if i >= args.warmup:
                rates.append((rate, dt))
        value = float(np.mean([r for r, _ in rates]))
        ms = 1e3 * float(np.mean([d for _, d in rates]))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": f"C5: one pre-embedded memory of 2^20 slots, d=256, 3 hops; {qs} queries x {slots} slots per step, scaled to the full memory"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{qs} queries x {slots} of 2^20 slots per step (oracle/qmo_bigmem.py), scaled linearly in the slot count"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return 0
    pkg = ge.import_package()
    synth = pkg.synth
    cfg, n, S, desc = workload(args.workload, synth)
    w = synth.make_weights(cfg, 0x5EED0000 + 1, sigma=SIGMA)
    # size the per-step sample so that warmup+steps stays within a few minutes
    probe_rate, _, threads, _ = cpu_oracle_rate(synth, cfg, w, S, 64)
    budget_s = 120.0 / max(1, args.steps + args.warmup)
    sample = int(max(64, min(n, probe_rate * min(budget_s, 15.0))))
    times = []
    for i in range(args.warmup + args.steps):
        r, dt, threads, _ = cpu_oracle_rate(synth, cfg, w, S, sample)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}; {sample}-story sample per step (full workload: {n} stories per GPU)",
                   "weights": f"N(0,{SIGMA}), layer-wise tied, EN_MQ formats (6,1)/(5,2)/(4,3), base (5,2)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} stories of {args.workload} per step, {args.steps} steps, OpenMP over stories; the reference's own "
                                   "CPU forward is dead code (lib/layer.c:254-437), so this is the C restatement of its CUDA arithmetic (oracle/)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_reference_gpu(args):
    """--impl reference_gpu: the reference's OWN CUDA forward (unmodified lib/layer.c + lib/layer_cuda.cu rebuilt for sm_100a by
    oracle/Makefile, one story per 31 launches) on this box's GPU 0, on a bounded sample of the workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    pkg = ge.import_package()
    synth = pkg.synth
    if args.workload not in ("C1", "C2", "C3"):
        print(json.dumps({"impl": "reference_gpu", "unavailable": "the reference cannot launch this shape (dimensions above 1024 / no such preset)"}), flush=True)
        return 0
    cfg, n, S, desc = workload(args.workload, synth)
    w = synth.make_weights(cfg, 0x5EED0000 + 1, sigma=SIGMA)
    sample = 2000
    r = reference_gpu_rate(synth, cfg, w, S, sample=sample, reps=max(1, min(args.steps, 3)))
    if not r or "unavailable" in r:
        print(json.dumps({"impl": "reference_gpu", "unavailable": (r or {}).get("unavailable", "oracle/_ref/ref_harness_refcuda was not built")}), flush=True)
        return 0
    print(json.dumps({
        "impl": "reference_gpu", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["s_per_pass"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}; {sample}-story sample per step", "kind": r["kind"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 31 * sample}), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference_gpu"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--queries", type=str, default="64",
                    help="C5: queries per step; a comma list (e.g. 1,64,1024) runs them one after the other on the same resident memory "
                         "and prints one JSON line each")
    ap.add_argument("--input", default="dense", choices=["dense", "ids"],
                    help="C1-C4: dense fp32 arenas (the reference boundary format, the headline) or the word-id lists they are "
                         "built from (SURVEY 8f-1, reported separately, never mixed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 8)")
    ap.add_argument("--stories", type=int, default=0, help="C1-C4: stories per GPU and step (default: the workload's own count); BASELINE configs[3] sweeps 2^10 .. 2^20")
    ap.add_argument("--c5-attention", default="dot", choices=["dot", "hamming"], help="C5: attention mode 2 (default) or 3")
    ap.add_argument("--quick", action="store_true", help="headline only: no sensitivity / e2e_ids / c5 / reference_gpu sections (profiling runs)")
    args = ap.parse_args()
    global C5_MODE
    C5_MODE = 3 if args.c5_attention == "hamming" else 2
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "reference_gpu":
        return run_reference_gpu(args)
    if args.workload == "C5":
        return run_bigmem(args)

    import torch
    rank, world, local = dist_setup(args.gpus)
    if world == 1 and args.gpus > 1:
        print("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)", file=sys.stderr)
        return 2
    torch.cuda.set_device(local)
    pkg = ge.import_package()
    synth, qlib = pkg.synth, pkg.lib
    cfg, n, S, desc = workload(args.workload, synth)
    if args.stories > 0:
        n = args.stories
    W = max(3, args.warmup)
    K = max(1, args.steps)

    weights = synth.make_weights(cfg, 0x5EED0000 + 1, sigma=SIGMA)
    st = synth.make_stories(cfg, n, 0x5EED1000 + 1 + rank, S=S)        # every rank: its own shard, same size
    model = qlib.Model(cfg, weights, device=f"cuda:{local}")
    use_ids = args.input == "ids"
    ist = synth.ids_from_dense(st) if use_ids else None
    db = model.upload_ids(ist) if use_ids else model.upload(st)
    stream = torch.cuda.current_stream()
    # the id lists of a step (~16 MB) fit the 126 MB L2: evict them between timed steps by writing a 512 MB buffer
    # (outside the per-step events); the dense arenas (1 GB) are larger than L2 and need no flush
    small = (4 * cfg.V * (S + 1) * n) < (200 << 20)      # inputs that fit the 126 MB L2: evict them between timed steps
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=f"cuda:{local}") if (use_ids or small) else None

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    for _ in range(W):
        model.forward(db, with_answers=False)
    barrier()
    model.path_counts()                 # reset the tier counters: the timed region alone is counted
    model.profile(True)
    model.profile_read()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = qlib.lib().qmann_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if use_ids or small:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for a_, b_ in evs:
            flush.zero_()
            a_.record(stream)
            model.forward(db, with_answers=False)
            b_.record(stream)
        barrier()
        ms_total = sum(a_.elapsed_time(b_) for a_, b_ in evs)
    else:
        ev0.record(stream)
        for _ in range(K):
            model.forward(db, with_answers=False)
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    launches = qlib.lib().qmann_launch_count() - launches0
    ms_compact, ms_forward, pairs = model.profile_read()
    model.profile(False)
    tiers = model.path_counts()
    # keep the GPU busy a little longer if the timed region was too short for the 100 ms sampler
    t_end = time.time() + max(0.0, 0.6 - ms_total / 1e3)
    while time.time() < t_end:
        model.forward(db, with_answers=False)
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    pred_dev = db.pred[:n].cpu().numpy().astype(np.uint32)

    # ---------------- end to end: pinned host arenas in, predictions out ----------------
    Ke = args.e2e_steps or min(K, 8)
    if use_ids:
        ids_pin = torch.from_numpy(ist.ids.view(np.int16)).pin_memory().numpy().view(np.uint16)
        off_pin = torch.from_numpy(ist.row_off.view(np.int32)).pin_memory().numpy().view(np.uint32)
        ans_pin = torch.from_numpy(ist.ans.view(np.int32)).pin_memory().numpy().view(np.uint32)
        e2e_call = lambda: model.infer_ids_host(ids_pin, off_pin, ans_pin, st.n_sen)
        h2d = int(ist.ids.nbytes + ist.row_off.nbytes + ist.ans.nbytes)
    else:
        m_pin = torch.from_numpy(st.m).pin_memory()
        q_pin = torch.from_numpy(st.q).pin_memory()
        a_pin = torch.from_numpy(st.a).pin_memory()
        e2e_call = lambda: model.infer_host(m_pin, q_pin, a_pin, st.n_sen)
        h2d = int(st.m.nbytes + st.q.nbytes + st.a.nbytes)
    for _ in range(2):
        pred_h, match_h, _ = e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        pred_h, match_h, _ = e2e_call()      # synchronous: returns host predictions
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / Ke
    assert np.array_equal(pred_h, pred_dev), "end-to-end predictions differ from the device-resident pass"
    assert match_h == int((pred_h == st.ans).sum())
    d2h = int(4 * n + 4)
    # the ceiling of that figure: the same bytes as one plain pinned-host -> device copy, all ranks at the same time (the ranks of a box
    # share the host's memory and PCIe root complexes)
    copy_gbs = None
    if not use_ids:
        stage = torch.empty_like(db.m)
        for _ in range(2):
            stage.copy_(m_pin, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            stage.copy_(m_pin, non_blocking=True)
        torch.cuda.synchronize()
        copy_gbs = 3 * st.m.nbytes / (time.perf_counter() - t0) / 1e9
        del stage

    # ---------------- the same step with weights that saturate more (which tier finishes how many stories) ----------------
    def timed_steps(mdl, dbatch, k):
        for _ in range(3):
            mdl.forward(dbatch, with_answers=False)
        torch.cuda.synchronize()
        mdl.path_counts()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            mdl.forward(dbatch, with_answers=False)
        e1.record(stream)
        torch.cuda.synchronize()
        t3 = mdl.path_counts()
        tot = max(1, k * n)
        return e0.elapsed_time(e1) / k, {"packed": t3[0] / tot, "unpacked": t3[1] / tot, "general": t3[2] / tot}
    sensitivity = {}
    if not use_ids and not args.quick:
        for sg in (1.0, 2.0):
            w2 = synth.make_weights(cfg, 0x5EED0000 + 1, sigma=sg)
            m2 = qlib.Model(cfg, w2, device=f"cuda:{local}")
            ms2, share2 = timed_steps(m2, db, max(3, K // 2))
            sensitivity[f"sigma_{sg:g}"] = {"ms_per_step": ms2, "stories_per_s_per_gpu": n / (ms2 / 1e3), "path_share": share2}
            m2.close()

    # ---------------- end to end with the word-id lists (separate figure, never mixed with the headline) ----------------
    e2e_ids = None
    if not use_ids and not args.quick:
        ist2 = synth.ids_from_dense(st)
        ids_pin = torch.from_numpy(ist2.ids.view(np.int16)).pin_memory().numpy().view(np.uint16)
        off_pin = torch.from_numpy(ist2.row_off.view(np.int32)).pin_memory().numpy().view(np.uint32)
        ans_pin = torch.from_numpy(ist2.ans.astype(np.uint32).view(np.int32)).pin_memory().numpy().view(np.uint32)
        for _ in range(2):
            p_i, m_i, _ = model.infer_ids_host(ids_pin, off_pin, ans_pin, st.n_sen)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            p_i, m_i, _ = model.infer_ids_host(ids_pin, off_pin, ans_pin, st.n_sen)
        torch.cuda.synchronize()
        ids_ms = 1e3 * (time.perf_counter() - t0) / Ke
        assert np.array_equal(p_i, pred_dev), "word-id input predicts differently from the dense input"
        e2e_ids = {"ms_per_step": ids_ms, "h2d_bytes_per_step": int(ist2.ids.nbytes + ist2.row_off.nbytes + ist2.ans.nbytes), "d2h_bytes_per_step": d2h}
        del ist2

    # ---------------- the workload with a collective: slot-sharded large memory, same process group ----------------
    c5 = None
    if not use_ids and not args.quick and args.workload == "C2":
        try:
            c5 = c5_record(torch, synth, qlib, rank, world, local, K=max(4, min(K, 10)))
        except AssertionError:
            raise
        except Exception as e:          # noqa: BLE001  (an allocation failure must not lose the headline line)
            c5 = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    # ---------------- max over ranks ----------------
    ms_step = ms_total / K
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_step, e2e_ms, ms_compact / max(1, pairs), ms_forward / max(1, pairs), e2e_ids["ms_per_step"] if e2e_ids else 0.0,
                          -(copy_gbs or 0.0)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_gbs = -float(t[5]) if copy_gbs else None          # the slowest rank's copy rate
        ms_step, e2e_ms = float(t[0]), float(t[1])
        k_compact_ms, k_forward_ms = float(t[2]), float(t[3])
        if e2e_ids:
            e2e_ids["ms_per_step"] = float(t[4])
    else:
        k_compact_ms, k_forward_ms = ms_compact / max(1, pairs), ms_forward / max(1, pairs)

    if rank == 0:
        total_stories = n * world
        value = total_stories / (ms_step / 1e3)
        e2e_value = total_stories / (e2e_ms / 1e3)
        # algorithmic bytes per story (primary model): dense fp32 BoW rows + question in, 4-byte answer out
        bytes_story = 4 * cfg.V * (S + 1) + 4
        if use_ids:
            # secondary model: the id lists and their row offsets in, 4-byte answer out
            bytes_story = (ist.ids.nbytes + ist.row_off.nbytes) / n + 4
        peak, peak_src = measured_peak()
        launches_per_step = pairs / K
        stories_per_launch = n / max(1.0, launches_per_step)
        step_bytes = bytes_story * n                      # algorithmic bytes one rank moves per step
        launch_bytes = bytes_story * stories_per_launch   # ... and per launch (batches above the chunk size run several launches per step)
        tc_tier = os.environ.get("QMANN_TC") == "1"
        # Whole-step fraction: the number north_star's ">= 60 % of the HBM roofline" is held against.  The two kernels of a step
        # bind on different resources: k_compact streams the dense arenas once (HBM-bound), the forward tiers read only the
        # compact records (L2) and are bound by instruction issue.
        whole = step_bytes / (ms_step / 1e3) / 1e9
        cap = {k: ncu_traffic(k) for k in ("k_compact", "k_story", "k_story_tc", "k_ids_compact")} if (args.workload == "C2" and stories_per_launch == 20000) else {}
        tr = lambda k: (cap.get(k) or {}).get("dram_bytes_per_launch")
        issue = (cap.get("k_story_tc" if tc_tier else "k_story") or {})
        issue_slots = 148 * 4 * 1.965e9                   # warp instructions per second at one per scheduler and clock
        fwd_inst = issue.get("warp_instructions_per_launch")
        kernels = {
            "k_compact": {"ms": k_compact_ms, "bound": "hbm", "GB/s": launch_bytes / (k_compact_ms / 1e3) / 1e9 if k_compact_ms > 0.01 else None,
                          "frac_of_hbm": (launch_bytes / (k_compact_ms / 1e3) / 1e9 / peak) if k_compact_ms > 0.01 else None,
                          "dram_bytes_per_launch_ncu": tr("k_ids_compact" if use_ids else "k_compact")},
            "forward_tiers": {"ms": k_forward_ms, "bound": "instruction issue",
                              "kernel": ("k_story_tc (tcgen05 embedding; it also streams the dense rows, there is no k_compact pass)" if tc_tier else
                                         "k_story<packed> -> k_story<unpacked> -> k_forward on what each declines"),
                              "warp_instructions_per_launch_ncu": fwd_inst,
                              "issue_slot_frac": (fwd_inst / (issue_slots * k_forward_ms / 1e3)) if (fwd_inst and k_forward_ms > 0) else None,
                              "dram_bytes_per_launch_ncu": tr("k_story_tc" if tc_tier else "k_story")},
        }
        dom = "forward_tiers" if k_forward_ms >= k_compact_ms else "k_compact"
        roofline = {
            "bound": "hbm", "achieved": whole, "peak": peak, "unit": "GB/s", "frac": whole / peak,
            "traffic": (tr("k_story_tc") if tc_tier else ((tr("k_ids_compact" if use_ids else "k_compact") or 0) + (tr("k_story") or 0)) or None),
            "traffic_note": "dram read+write bytes of one step from the committed ncu --set full captures (profiles/r02_traffic.json), all kernels of "
                            f"the step; algorithmic bytes of the step {step_bytes:.0f}",
            "scope": "whole step: algorithmic bytes of the step / step time (CUDA events over the timed region)",
            "dominant_kernel": dom, "kernels": kernels, "peak_source": peak_src, "frac_of_nominal_8TBs": whole / NOMINAL_HBM_GBS,
            "algorithmic_bytes_per_story": bytes_story, "stories_per_launch": stories_per_launch,
        }
        if use_ids:
            roofline["note"] = ("word-id input: ~0.6 KB per story, so the HBM roofline is out of reach by construction; the step is "
                                "bound by the instruction issue rate of k_forward_fast (profiles/: smsp__issue_active)")
        ref_gpu = None
        if world == 1 and not args.quick and not use_ids and args.workload in ("C1", "C2", "C3"):
            ref_gpu = reference_gpu_rate(synth, cfg, weights, S, sample=1000, reps=1)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            probe, _, threads, _ = cpu_oracle_rate(synth, cfg, weights, S, 64)
            sample = int(max(256, min(n, probe * 12.0)))
            rate, dt, threads, ref = cpu_oracle_rate(synth, cfg, weights, S, sample)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{sample} stories of {args.workload} in {dt:.1f} s, OpenMP over stories (oracle/qmann_oracle.c; the reference "
                             "ships no runnable CPU forward)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}; {n} stories per GPU per step",
                       "weights": f"N(0,{SIGMA}), layer-wise tied, EN_MQ formats (6,1)/(5,2)/(4,3), base (5,2)",
                       "input_format": ("word-id lists (uint16 ids + uint32 row offsets, SURVEY 8f-1: the lists sample_vectorization scatters "
                                        "into the dense arenas), resident in HBM") if use_ids else
                                       "dense fp32 bag-of-words arenas (reference boundary format), resident in HBM",
                       "l2": (f"inputs {bytes_story * n / 1e6:.0f} MB per step < 126 MB L2: a 512 MB buffer is written between timed steps "
                              "(outside the per-step CUDA events) to evict them") if (use_ids or small) else
                             f"inputs {bytes_story * n / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                       "parallelism": f"batch-sharded x{world}, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "h2d_GB/s_per_rank": h2d / (e2e_ms / 1e3) / 1e9,
                    "host_copy_ceiling_GB/s_per_rank": copy_gbs,
                    "frac_of_host_copy_ceiling": (h2d / (e2e_ms / 1e3) / 1e9 / copy_gbs) if copy_gbs else None,
                    "steps": Ke, "api": ("qmann_infer_ids_host (pinned host id lists in, host predictions + match count out)" if use_ids else
                            "qmann_infer_host (pinned host arenas in, host predictions + match count out)")},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "path_share": {"packed": tiers[0] / max(1, K * n), "unpacked": tiers[1] / max(1, K * n), "general": tiers[2] / max(1, K * n),
                           "note": "share of the stories of the timed region that ENTERED each forward tier on rank 0 (a tier hands what it declines to the next)"},
            "sensitivity": sensitivity or None,
            "e2e_ids": ({"value": total_stories / (e2e_ids["ms_per_step"] / 1e3), "unit": UNIT, "ms_per_step": e2e_ids["ms_per_step"],
                         "h2d_bytes_per_step": e2e_ids["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e_ids["d2h_bytes_per_step"],
                         "api": "qmann_infer_ids_host (pinned host word-id lists in, host predictions out; SURVEY 8f-1 input format, not the headline)"}
                        if e2e_ids else None),
            "c5": c5,
            "reference_gpu": ref_gpu,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "accuracy_check": {"match": int(match_h), "stories": n},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
