"""Batch sharding (C1-C4) over two processes on the CPU (gloo): each rank takes the contiguous story range
qmann_shard_plan gives it, runs the forward on its shard only (here: the CPU oracle stands in for the device), and the
predictions gathered in rank order are the unsharded result; the match counters add up with one all_reduce.  This is
the whole N > 1 protocol of the batched path: no data-path collective."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import __graft_entry__ as ge
    pkg = ge.import_package()
    import qmo
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = pkg.synth.preset_config("C1")
    w = pkg.synth.make_weights(cfg, 5, sigma=0.5)
    st = pkg.synth.make_stories(cfg, 301, 6, ragged=True)            # every rank generates the same stories
    first, count = pkg.lib.shard_plan(st.n_sen, world, rank)
    off = st.offsets()
    sub = pkg.synth.Stories(m=st.m[off[first]:off[first + count]], q=st.q[first:first + count], a=st.a[first:first + count],
                            n_sen=st.n_sen[first:first + count], ans=st.ans[first:first + count])
    out = qmo.forward(cfg, w, sub)
    pred = torch.full((st.N,), -1, dtype=torch.int64)
    pred[first:first + count] = torch.from_numpy(out["pred"].astype(np.int64))
    dist.all_reduce(pred, op=dist.ReduceOp.MAX)                       # the host "concatenates" the disjoint ranges
    match = torch.tensor([int((out["pred"] == sub.ans).sum())])
    dist.all_reduce(match, op=dist.ReduceOp.SUM)
    if rank == 0:
        full = qmo.forward(cfg, w, st)
        ok = np.array_equal(pred.numpy(), full["pred"].astype(np.int64)) and int(match[0]) == int((full["pred"] == st.ans).sum())
        np.save(os.path.join(tmp, "ok.npy"), np.array([int(ok), first, count]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_batch_shard(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r = np.load(tmp_path / "ok.npy")
    assert r[0] == 1 and r[1] == 0 and 0 < r[2] < 301
