"""torchrun entry (not a pytest module): slot-sharded large-memory forward over NCCL, one rank per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/run_bigmem_nccl.py

Every rank holds S/world contiguous slots; the per-hop exchanges are all_reduce(SUM) of the integer score
histograms and of the integer partial reads.  Rank 0 also runs the unsharded memory on its own GPU and checks
that the sharded result is identical (controller state after every hop, predictions)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from test_bigmem_oracle import _random_memory
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    pkg = ge.import_package()
    ok = True
    dbg = lambda *a: print(f"[rank {rank}]", *a, flush=True) if os.environ.get("QMANN_NCCL_DEBUG") else None
    dbg("creating NcclComm")
    comm = pkg.lib.NcclComm()                          # our own communicator for the one-call C entry
    dbg("NcclComm ready")
    side = torch.cuda.Stream()
    for mode in (2, 3):
        cfg = pkg.synth.ModelConfig(V=40, d=64, S_max=64, V_dict=20, mode=mode, iwl=5 if mode == 2 else 3)
        w = pkg.synth.make_weights(cfg, 6, sigma=0.5)
        M8, C8, u0 = _random_memory(cfg, 40009 if mode == 2 else 3001, 12, 177, sigma=0.6 if mode == 2 else 0.3, plant_scale=3.0 if mode == 2 else 1.0)
        S = M8.shape[1]
        lo, n = pkg.lib.slot_shard(S, world, rank)
        mem = pkg.lib.BigMemory(cfg, w, M8[:, lo:lo + n], C8[:, lo:lo + n], S, lo, Q_max=16, device=f"cuda:{local}", group=dist.group.WORLD, world=world)
        u0d = torch.from_numpy(u0).cuda()
        out = mem.forward(u0d, debug=True)
        torch.cuda.synchronize()
        # qmann_bigmem_forward_sharded: phases + ncclAllReduce inside the library, captured into a CUDA graph on a side stream
        pred_phase = out["pred"].clone()
        for rep in range(3):                               # un-captured first call, capture, replay
            dbg(f"mode {mode} forward_sharded rep {rep}")
            with torch.cuda.stream(side):
                pred_one = mem.forward_sharded(u0d, comm).clone()
            side.synchronize()
            dbg(f"mode {mode} forward_sharded rep {rep} done")
            same = torch.equal(pred_one, pred_phase)
            ok = ok and same
            if not same:
                print(f"mode {mode}: one-call sharded forward (rep {rep}) differs from the phase API on rank {rank}", flush=True)
        out["pred"] = pred_phase
        if rank == 0:
            full = pkg.lib.BigMemory(cfg, w, M8, C8, S, 0, Q_max=16, device=f"cuda:{local}")
            ref = full.forward(u0d, debug=True)
            torch.cuda.synchronize()
            for k in ("u", "o", "g", "hist", "pred"):
                same = torch.equal(out[k], ref[k])
                ok = ok and same
                if not same:
                    print(f"mode {mode}: {k} differs between {world} shards and 1 shard", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("BIGMEM_NCCL_OK" if int(flag.item()) == 1 else "BIGMEM_NCCL_FAIL", flush=True)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
