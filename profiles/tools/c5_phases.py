"""Per-phase CUDA-event times of one large-memory hop (scores + histogram, softmax + read, update) on ONE GPU holding the shard that
rank 0 of `world` ranks would hold:  python profiles/tools/c5_phases.py Q world   (QMANN_BIGMEM_* switches apply; C5_MODE=3: Hamming)."""
import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as ge
import bench
pkg = ge.import_package()
synth, qlib = pkg.synth, pkg.lib
Q = int(sys.argv[1]); world = int(sys.argv[2])
bench.C5_MODE = int(os.environ.get("C5_MODE", "2"))
cfg, weights, mem, u0, n_loc = bench.c5_build(torch, synth, qlib, world, 0, "cuda:0", Q, None)
L = qlib.lib()
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = [0.0, 0.0, 0.0]
for rep in range(6):
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    qlib._bcheck(L.qmann_bigmem_begin(mem._h, u0.data_ptr(), Q, st))
    hist, partial = mem.hist[:Q], mem.partial[:Q]
    for h in range(cfg.H):
        e = [ev() for _ in range(4)]
        e[0].record()
        qlib._bcheck(L.qmann_bigmem_hop_scores(mem._h, h, hist.data_ptr(), st))
        e[1].record()
        qlib._bcheck(L.qmann_bigmem_hop_read(mem._h, h, hist.data_ptr(), partial.data_ptr(), None, st))
        e[2].record()
        qlib._bcheck(L.qmann_bigmem_hop_update(mem._h, h, partial.data_ptr(), None, None, st))
        e[3].record()
        torch.cuda.synchronize()
        if rep >= 3:
            for k in range(3): acc[k] += e[k].elapsed_time(e[k + 1]) / 9
print(f"Q={Q} shard 1/{world} ({n_loc} slots): per hop scores+hist {acc[0]:.3f}  softmax+read {acc[1]:.3f}  update {acc[2]:.3f} ms  tq_wide={os.environ.get('QMANN_BIGMEM_TQ_WIDE','0')}", flush=True)
