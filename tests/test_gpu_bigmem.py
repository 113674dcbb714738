"""Slot-sharded large-memory path (BASELINE config 5) on a real B200, through the C ABI (qmann_bigmem_*).

Bars: score bins, selected slots, reads, linear maps, updates and predicted answers bit-exact against the CPU
restatement (oracle/qmo_bigmem.py, itself pinned to the reference's golden tensors by tests/test_bigmem_oracle.py);
attention weights within 1e-5 relative.  At full size (2^20 slots, d = 256) the checks are the size-independent
ones: the result does not depend on the number of shards, the histogram counts every slot once, and a planted
copy of the query is the slot that gets read."""
import os
import subprocess
import sys

import numpy as np
import pytest

from test_bigmem_oracle import _random_memory

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(qmann, cfg, w, M8, C8, u0, shards=1, debug=True):
    """Forward over `shards` slot shards held by ONE GPU; the two per-hop exchanges are done by adding the
    shards' integer buffers (what all_reduce does across ranks)."""
    import ctypes as C
    import torch
    L = qmann.lib.lib()
    S = M8.shape[1]
    Q = u0.shape[0]
    mems = []
    for r in range(shards):
        lo, n = qmann.lib.slot_shard(S, shards, r)
        mems.append(qmann.lib.BigMemory(cfg, w, M8[:, lo:lo + n], C8[:, lo:lo + n], S, lo, Q_max=Q))
    if shards == 1:
        out = mems[0].forward(torch.from_numpy(u0).cuda(), debug=debug)
        torch.cuda.synchronize()
        return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    u0d = torch.from_numpy(u0).cuda()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    H, d = cfg.H, cfg.d
    res = dict(u=np.zeros((H, Q, d), np.int8), hist=np.zeros((H, Q, mems[0].NB), np.int64))
    for m in mems:
        qmann.lib._bcheck(L.qmann_bigmem_begin(m._h, u0d.data_ptr(), Q, st))
    for h in range(H):
        for m in mems:
            qmann.lib._bcheck(L.qmann_bigmem_hop_scores(m._h, h, m.hist.data_ptr(), st))
        tot = sum(m.hist[:Q] for m in mems)
        res["hist"][h] = tot.cpu().numpy()
        for m in mems:
            m.hist[:Q].copy_(tot)
            qmann.lib._bcheck(L.qmann_bigmem_hop_read(m._h, h, m.hist.data_ptr(), m.partial.data_ptr(), None, st))
        ptot = sum(m.partial[:Q] for m in mems)
        for m in mems:
            m.partial[:Q].copy_(ptot)
            qmann.lib._bcheck(L.qmann_bigmem_hop_update(m._h, h, m.partial.data_ptr(), None, None, st))
        uu = torch.zeros((Q, d), dtype=torch.int8, device="cuda")
        qmann.lib._bcheck(L.qmann_bigmem_state(mems[0]._h, uu.data_ptr(), None, st))
        res["u"][h] = uu.cpu().numpy()
    pred = torch.zeros(Q, dtype=torch.int32, device="cuda")
    qmann.lib._bcheck(L.qmann_bigmem_finish(mems[0]._h, pred.data_ptr(), None, None, st))
    torch.cuda.synchronize()
    res["pred"] = pred.cpu().numpy()
    return res


@pytest.mark.parametrize("mode,d,S,Q,iwl", [(2, 32, 3000, 5, 5), (2, 64, 1031, 1, 5), (2, 48, 2500, 17, 5), (2, 32, 2000, 4, 3),
                                            (3, 32, 600, 5, 2), (3, 64, 900, 1, 3), (3, 64, 1500, 6, 3)])
def test_bigmem_matches_oracle(mode, d, S, Q, iwl, qmann, synth, qmo):
    import qmo_bigmem as qb
    cfg = synth.ModelConfig(V=40, d=d, S_max=64, V_dict=20, mode=mode, iwl=iwl)
    w = synth.make_weights(cfg, 5, sigma=0.5)
    M8, C8, u0 = _random_memory(cfg, S, Q, 100 + d + Q, sigma=0.6 if mode == 2 else 0.3, plant_scale=3.0 if mode == 2 else 1.0)
    ref = qb.forward(cfg, w, M8, C8, u0)
    got = _run(qmann, cfg, w, M8, C8, u0)
    np.testing.assert_array_equal(got["hist"].astype(np.int64), ref["hist"].astype(np.int64), err_msg="score histograms")
    safe = ref["risk"] == 0
    assert safe.any()
    np.testing.assert_allclose(got["pbin"][:, safe], ref["pbin"][:, safe], rtol=1e-5, atol=1e-30)
    for k in ("o", "g", "u"):
        np.testing.assert_array_equal(got[k][:, safe], ref[k][:, safe], err_msg=k)
    np.testing.assert_array_equal(got["z"][safe], ref["z"][safe])
    np.testing.assert_array_equal(got["pred"][safe].astype(np.uint32), ref["pred"][safe])
    assert ref["nsel"].sum() > 0


@pytest.mark.parametrize("d,S,Q,sigma,plant", [(16, 4099, 1, 0.6, 3.0), (32, 3001, 2, 2.0, 3.0), (64, 2050, 5, 0.2, 1.0), (128, 1500, 3, 1.0, 3.0),
                                                 (256, 1111, 4, 0.6, 3.0), (256, 777, 1, 4.0, 3.0), (512, 515, 7, 0.6, 2.0),
                                                 (64, 1999, 64, 0.6, 3.0), (256, 1030, 70, 0.3, 2.0), (128, 523, 9, 3.0, 3.0),
                                                 (256, 5003, 200, 0.6, 3.0), (128, 130, 64, 1.5, 3.0),
                                                 (256, 4224, 130, 0.6, 3.0), (128, 2048, 64, 1.5, 3.0), (256, 70000, 8, 0.6, 3.0)])
def test_bigmem_fast_scorer_equals_per_product(d, S, Q, sigma, plant, qmann, synth, monkeypatch):
    """k_big_scores_fast (packed low-bit / dp4a form) and, for Q >= 4, the tensor-core scorers -- k_big_scores_tq (tcgen05.mma
    kind::i8 with the query planes in tensor memory, 128 queries per pass; d = 128 / 256), k_big_scores_tc (query
    planes in shared memory; d % 128 == 0) and k_big_scores_mma (mma.sync, d % 64 == 0): four int8
    contractions each -- saturating rows recomputed product by product, against the per-product kernel
    k_big_scores on the same memory: identical histograms, controller states and answers.
    sigma = 2..4 makes most rows saturate somewhere (the in-kernel exact path), sigma = 0.2 none."""
    cfg = synth.ModelConfig(V=40, d=d, S_max=64, V_dict=20, mode=2, iwl=5)
    w = synth.make_weights(cfg, 5, sigma=0.5)
    M8, C8, u0 = _random_memory(cfg, S, Q, 900 + d + Q, sigma=sigma, plant_scale=plant)
    M8[0, 7, :] = -128                      # outside the format: Q_att clamps it to -127 in both kernels
    M8[1, 11, :3] = 127
    monkeypatch.setenv("QMANN_BIGMEM_FAST", "0")
    slow = _run(qmann, cfg, w, M8, C8, u0)
    monkeypatch.setenv("QMANN_BIGMEM_FAST", "1")
    fast = _run(qmann, cfg, w, M8, C8, u0)
    for k in ("hist", "o", "g", "u", "z", "pred"):
        np.testing.assert_array_equal(fast[k], slow[k], err_msg=k)
    many = _run(qmann, cfg, w, M8, C8, u0, shards=3)
    np.testing.assert_array_equal(many["hist"], slow["hist"].astype(np.int64))
    np.testing.assert_array_equal(many["u"], slow["u"])
    # the histogram above came from k_big_hist_lanes (lane-private counters); now the shared-memory-atomic kernel
    monkeypatch.setenv("QMANN_BIGMEM_HIST_LANES", "0")
    atom = _run(qmann, cfg, w, M8, C8, u0)
    np.testing.assert_array_equal(atom["hist"], slow["hist"], err_msg="k_big_hist")
    np.testing.assert_array_equal(atom["pred"], slow["pred"])
    monkeypatch.delenv("QMANN_BIGMEM_HIST_LANES")
    if Q > 128:
        # several query blocks: one scorer launch per block, the histogram of a block on a second stream under the next scorer
        monkeypatch.setenv("QMANN_BIGMEM_OVERLAP", "2")
        ovl = _run(qmann, cfg, w, M8, C8, u0)
        for k in ("hist", "o", "u", "pred"):
            np.testing.assert_array_equal(ovl[k], slow[k], err_msg=f"overlapped histogram: {k}")
        monkeypatch.delenv("QMANN_BIGMEM_OVERLAP")
    # d <= 256 took k_big_scores_tq above (queries in tensor memory, 128 per pass); now k_big_scores_tc + k_big_hist
    monkeypatch.setenv("QMANN_BIGMEM_TQ", "0")
    tcs = _run(qmann, cfg, w, M8, C8, u0)
    for k in ("hist", "o", "u", "pred"):
        np.testing.assert_array_equal(tcs[k], slow[k], err_msg=f"tcgen05 scorer (queries in shared memory): {k}")
    # the mma.sync scorer alone (tcgen05 scorer switched off), then the packed CUDA-core kernel alone
    monkeypatch.setenv("QMANN_BIGMEM_TC", "0")
    mma = _run(qmann, cfg, w, M8, C8, u0)
    for k in ("hist", "u", "pred"):
        np.testing.assert_array_equal(mma[k], slow[k], err_msg=f"mma.sync scorer: {k}")
    monkeypatch.setenv("QMANN_BIGMEM_MMA", "0")
    packed = _run(qmann, cfg, w, M8, C8, u0)
    for k in ("hist", "u", "pred"):
        np.testing.assert_array_equal(packed[k], slow[k], err_msg=k)


@pytest.mark.parametrize("iwl,d,S,Q", [(3, 64, 1500, 21), (2, 32, 700, 17), (3, 256, 640, 19), (4, 48, 900, 5), (3, 64, 333, 7)])
def test_bigmem_hamming_nine_bit_equals_literal(iwl, d, S, Q, qmann, synth, qmo, monkeypatch):
    """Mode 3 on the large memory: the packed-byte scorer k_big_scores_ham (four elements per instruction) and the scalar nine-bit form
    (QMANN_BIGMEM_HAM=0) of the approximate Hamming element (w = |A_m - A_u| on sat9(code << shift)) against
    the literal 31-bit sign-magnitude form (QMANN_BIGMEM_FAST=0) and the oracle.  The memory holds the extremes (+-127, the codes that
    saturate the nine-bit operand, the -2^iwl value on the memory side); one query of the LAST block holds the -2^iwl value, which
    sends that block down the literal path while the first block stays on the fast one."""
    import qmo_bigmem as qb
    cfg = synth.ModelConfig(V=40, d=d, S_max=64, V_dict=20, mode=3, iwl=iwl)
    w = synth.make_weights(cfg, 5, sigma=0.5)
    M8, C8, u0 = _random_memory(cfg, S, Q, 4000 + d + Q, sigma=0.3, plant_scale=1.0)
    f = cfg.formats()
    rng = np.random.default_rng(7)
    for h in range(cfg.H):
        sh_m = 8 - f["iwl_att"][h] - f["frac_w"][h]
        M8[h, 3, :] = 127
        M8[h, 4, :] = -127
        M8[h, 5, ::2] = -128
        if 0 <= sh_m <= 7:
            M8[h, 6, :] = -(256 >> sh_m) if (256 >> sh_m) <= 127 else -127      # code << sh_m == -256: the value that encodes to 0
            M8[h, 7, :] = (256 >> sh_m) - 1 if (256 >> sh_m) <= 128 else 127
        M8[h, 8] = rng.integers(-127, 128, d)
    sh_u = 8 - f["iwl_att"][0] - f["frac_w"][0]
    if 0 <= sh_u <= 7 and (256 >> sh_u) <= 127:
        u0[Q - 1, 1] = -(256 >> sh_u)
    u0[0, :4] = [127, -127, 0, 1]
    fast = _run(qmann, cfg, w, M8, C8, u0)
    if d <= 64:                                 # (the oracle recovers the raw sums from the score values: not when a sum saturates, d = 256)
        ref = qb.forward(cfg, w, M8, C8, u0)
        np.testing.assert_array_equal(fast["hist"].astype(np.int64), ref["hist"].astype(np.int64), err_msg="nine-bit form vs oracle")
    monkeypatch.setenv("QMANN_BIGMEM_HAM", "0")
    nine = _run(qmann, cfg, w, M8, C8, u0)
    monkeypatch.setenv("QMANN_BIGMEM_FAST", "0")
    lit = _run(qmann, cfg, w, M8, C8, u0)
    for k in ("hist", "o", "g", "u", "pred"):
        np.testing.assert_array_equal(fast[k], lit[k], err_msg=f"packed-byte scorer vs literal kernel: {k}")
        np.testing.assert_array_equal(nine[k], lit[k], err_msg=f"nine-bit form vs literal kernel: {k}")


@pytest.mark.parametrize("mode", [2, 3])
def test_bigmem_shards_agree(mode, qmann, synth):
    cfg = synth.ModelConfig(V=40, d=64, S_max=64, V_dict=20, mode=mode, iwl=5 if mode == 2 else 3)
    w = synth.make_weights(cfg, 6, sigma=0.5)
    M8, C8, u0 = _random_memory(cfg, 10007 if mode == 2 else 2003, 9, 77, sigma=0.6 if mode == 2 else 0.3, plant_scale=3.0 if mode == 2 else 1.0)
    one = _run(qmann, cfg, w, M8, C8, u0, shards=1)
    for shards in (2, 3, 8):
        many = _run(qmann, cfg, w, M8, C8, u0, shards=shards)
        np.testing.assert_array_equal(many["hist"], one["hist"].astype(np.int64))
        np.testing.assert_array_equal(many["u"], one["u"])
        np.testing.assert_array_equal(many["pred"], one["pred"])


def test_bigmem_full_size_properties(qmann, synth):
    """BASELINE config 5 shape: 2^20 slots, d = 256, 3 hops."""
    import torch
    cfg = synth.ModelConfig(V=64, d=256, S_max=64, V_dict=32, mode=2)
    w = synth.make_weights(cfg, 9, sigma=0.3)
    S, Q = 1 << 20, 4
    g = torch.Generator(device="cuda").manual_seed(1)
    f = cfg.formats()
    M8 = [(torch.randn((S, cfg.d), device="cuda", generator=g) * (0.1 * (1 << f["frac_w"][h]))).round().clamp(-127, 127).to(torch.int8) for h in range(3)]
    C8 = [(torch.randn((S, cfg.d), device="cuda", generator=g) * (0.5 * (1 << f["frac_w"][h]))).round().clamp(-127, 127).to(torch.int8) for h in range(3)]
    u0 = (torch.randn((Q, cfg.d), device="cuda", generator=g) * 3).round().clamp(-127, 127).to(torch.int8)
    planted = [12345, 700001, 1048575, 3]
    for q, r in enumerate(planted):
        M8[0][r] = (u0[q].to(torch.int32) * 2).clamp(-127, 127).to(torch.int8)
    full = qmann.lib.BigMemory(cfg, w, M8, C8, S, 0, Q_max=Q)
    a = full.forward(u0, debug=True)
    torch.cuda.synchronize()
    hist = a["hist"].cpu().numpy().astype(np.int64)
    assert np.all(hist.sum(axis=2) == S), "every slot is counted once per query and hop"
    # hop 0: the planted slot holds the largest score of its query, and it is the row that is read
    u_after0 = a["u"][0].cpu().numpy()
    top_bin = np.array([np.nonzero(hist[0, q])[0].max() for q in range(Q)])
    cnt = hist[0, np.arange(Q), top_bin]
    assert np.all((cnt >= 1) & (cnt <= Q)), "only planted rows can reach the top score bin"
    # sharded 4 ways on the same GPU: identical result
    parts = []
    for r in range(4):
        lo, n = qmann.lib.slot_shard(S, 4, r)
        parts.append(qmann.lib.BigMemory(cfg, w, [m[lo:lo + n] for m in M8], [c[lo:lo + n] for c in C8], S, lo, Q_max=Q))
    import ctypes as C
    L = qmann.lib.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for m in parts:
        qmann.lib._bcheck(L.qmann_bigmem_begin(m._h, u0.data_ptr(), Q, st))
    for h in range(cfg.H):
        for m in parts:
            qmann.lib._bcheck(L.qmann_bigmem_hop_scores(m._h, h, m.hist.data_ptr(), st))
        tot = sum(m.hist[:Q] for m in parts)
        assert torch.equal(tot, a["hist"][h])
        for m in parts:
            m.hist[:Q].copy_(tot)
            qmann.lib._bcheck(L.qmann_bigmem_hop_read(m._h, h, m.hist.data_ptr(), m.partial.data_ptr(), None, st))
        ptot = sum(m.partial[:Q] for m in parts)
        for m in parts:
            m.partial[:Q].copy_(ptot)
            qmann.lib._bcheck(L.qmann_bigmem_hop_update(m._h, h, m.partial.data_ptr(), None, None, st))
        uu = torch.zeros((Q, cfg.d), dtype=torch.int8, device="cuda")
        qmann.lib._bcheck(L.qmann_bigmem_state(parts[0]._h, uu.data_ptr(), None, st))
        assert torch.equal(uu, a["u"][h]), f"hop {h}: sharded controller state differs"
    assert u_after0.shape == (Q, cfg.d)


def test_bigmem_nccl_two_ranks(qmann):
    """Two processes, one GPU each, NCCL all_reduce between the phases (needs >= 2 GPUs on the box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "run_bigmem_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "BIGMEM_NCCL_OK" in r.stdout


@pytest.mark.parametrize("mode,d,Q", [(2, 256, 70), (2, 64, 3), (3, 64, 5), (2, 256, 300)])
def test_bigmem_one_call_forward_equals_phase_api(mode, d, Q, qmann, synth, monkeypatch):
    """qmann_bigmem_forward_sharded on a single shard (no communicator): the whole forward as one C call, captured into a CUDA graph
    on a side stream and replayed, and un-captured on the default stream -- same predictions and controller state as the phase API."""
    import torch
    cfg = synth.ModelConfig(V=40, d=d, S_max=64, V_dict=20, mode=mode, iwl=5 if mode == 2 else 3)
    w = synth.make_weights(cfg, 8, sigma=0.5)
    M8, C8, u0 = _random_memory(cfg, 3001, Q, 1234 + d, sigma=0.6 if mode == 2 else 0.3, plant_scale=3.0 if mode == 2 else 1.0)
    if Q > 128:
        # several query blocks: the per-block launches with the histogram on a second stream (forced: the shard is small), whose
        # fork / join must also survive the graph capture
        monkeypatch.setenv("QMANN_BIGMEM_OVERLAP", "0")
        plain = qmann.lib.BigMemory(cfg, w, M8, C8, M8.shape[1], 0, Q_max=Q).forward(torch.from_numpy(u0).cuda())
        torch.cuda.synchronize()
        pred_plain = plain["pred"].clone()
        monkeypatch.setenv("QMANN_BIGMEM_OVERLAP", "2")
    mem = qmann.lib.BigMemory(cfg, w, M8, C8, M8.shape[1], 0, Q_max=Q)
    u0d = torch.from_numpy(u0).cuda()
    ref = mem.forward(u0d)
    torch.cuda.synchronize()
    pred_ref, u_ref = ref["pred"].clone(), ref["u_final"].clone()
    if Q > 128:
        assert torch.equal(pred_ref, pred_plain), "overlapped histogram differs from the single-stream hop"
    side = torch.cuda.Stream()
    u0b = u0d.clone()
    for rep in range(3):                       # capture + two replays; then another input buffer forces a re-capture
        with torch.cuda.stream(side):
            got = mem.forward_sharded(u0d if rep < 2 else u0b).clone()
        side.synchronize()
        assert torch.equal(got, pred_ref), f"graph path, call {rep}"
    got = mem.forward_sharded(u0d).clone()     # legacy default stream: not captured
    torch.cuda.synchronize()
    assert torch.equal(got, pred_ref)
    uu = torch.zeros_like(u_ref)
    qmann.lib._bcheck(qmann.lib.lib().qmann_bigmem_state(mem._h, uu.data_ptr(), None, None))
    torch.cuda.synchronize()
    assert torch.equal(uu, u_ref)
