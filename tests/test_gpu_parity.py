"""Parity of the CUDA path (called through the C ABI of libqmann_b200.so) on a real B200.

Bars (north_star): bit-exact for every fixed-point / Hamming quantity, selected slots and predicted
answers; fp32 softmax values within 1e-5 relative (they are in fact expected to be bit-identical to
the reference's, because both use __expf + a sequential double total on the same GPU)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu
SOFTMAX_RTOL = 1e-5
FIXED_KEYS = ("u0", "M", "C", "s", "o", "g", "u")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fwd(qmann, cfg, w, st, debug=True, want_h=True):
    import torch
    model = qmann.lib.Model(cfg, w)
    db = model.upload(st)
    out = model.forward(db, with_answers=True, want_h=want_h, debug=debug)
    torch.cuda.synchronize()
    res = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    res["pred"] = res["pred"][:st.N].astype(np.uint32)
    res["h_true"] = res["h_true"][:st.N]
    res["match"] = int(res["match"][0])
    return res


def _assert_fixed_equal(got, ref, cfg, tag):
    for k in FIXED_KEYS:
        if k == "g" and not cfg.lin_map:
            continue
        np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{tag}: {k}")


@pytest.mark.parametrize("name", [n for n in golden_io.case_names() if n != "c1_mode1"])
def test_batched_forward_matches_reference_golden(name, qmann, synth):
    """Batched kernels vs tensors the unmodified reference CUDA code produced on a B200."""
    cfg, w, st, ref = golden_io.load_case(name, synth)
    got = _fwd(qmann, cfg, w, st)
    _assert_fixed_equal(got, ref, cfg, name)
    np.testing.assert_array_equal(got["z"], ref["z"], err_msg="answer logits")
    np.testing.assert_allclose(got["p"], ref["p"], rtol=SOFTMAX_RTOL, atol=0, err_msg="attention weights")
    np.testing.assert_allclose(got["h"], ref["h"], rtol=SOFTMAX_RTOL, atol=0, err_msg="answer probabilities")
    np.testing.assert_array_equal(got["pred"], ref["pred"])
    assert got["match"] == int(ref["match"])
    # same GPU, same __expf, same summation order: expect identical bits, report if not
    assert np.array_equal(got["p"], ref["p"]) and np.array_equal(got["h"], ref["h"]), \
        f"softmax values within {SOFTMAX_RTOL} but not bit-identical to the reference build"


@pytest.mark.parametrize("preset,sigma,seed", [("C1", 0.5, 1), ("C1", 1.0, 2), ("C1", 0.1, 3), ("C2", 0.5, 4), ("C3", 0.5, 5),
                                               ("C3", 1.0, 6), ("C4", 0.5, 7), ("C4", 0.25, 8)])
def test_batched_forward_matches_oracle(preset, sigma, seed, qmann, synth, qmo):
    """Seeded random models vs the CPU oracle.  Stories whose attention weights sit within libm-vs-
    MUFU distance of a truncation boundary (oracle `risk`) are compared up to that hop only."""
    cfg = synth.preset_config(preset)
    w = synth.make_weights(cfg, 100 + seed, sigma=sigma)
    st = synth.make_stories(cfg, 192, 200 + seed, S=min(cfg.S_max, 50), ragged=True)
    ref = qmo.forward(cfg, w, st)
    got = _fwd(qmann, cfg, w, st)
    safe = (ref["risk"] == 0)
    assert safe.mean() > 0.9
    np.testing.assert_array_equal(got["u0"], ref["u0"])
    off = st.offsets()
    rows_safe = np.repeat(safe, st.n_sen)
    for k in ("M", "C", "s"):
        np.testing.assert_array_equal(got[k][:, rows_safe], ref[k][:, rows_safe], err_msg=k)
    np.testing.assert_allclose(got["p"][:, rows_safe], ref["p"][:, rows_safe], rtol=SOFTMAX_RTOL, atol=1e-30)
    for k in ("o", "g", "u"):
        np.testing.assert_array_equal(got[k][:, safe], ref[k][:, safe], err_msg=k)
    np.testing.assert_array_equal(got["z"][safe], ref["z"][safe])
    np.testing.assert_allclose(got["h"][safe], ref["h"][safe], rtol=SOFTMAX_RTOL, atol=1e-30)
    clear = safe & (ref["risk_ans"] > 1e-4)
    np.testing.assert_array_equal(got["pred"][clear], ref["pred"][clear])
    np.testing.assert_allclose(got["h_true"][safe], ref["h_true"][safe], rtol=SOFTMAX_RTOL, atol=1e-30)
    assert off[-1] == st.sum_sen


@pytest.mark.parametrize("preset", ["C1", "C2", "C3", "C4"])
def test_fast_path_equals_debug_path(preset, qmann, synth):
    """The production launch (lazy output-memory rows, no probability pass) must predict exactly what
    the instrumented launch predicts."""
    cfg = synth.preset_config(preset)
    w = synth.make_weights(cfg, 11, sigma=0.5)
    st = synth.make_stories(cfg, 1000, 12, S=min(cfg.S_max, 50), ragged=True)
    a = _fwd(qmann, cfg, w, st, debug=True, want_h=True)
    b = _fwd(qmann, cfg, w, st, debug=False, want_h=False)
    c = _fwd(qmann, cfg, w, st, debug=False, want_h=True)
    np.testing.assert_array_equal(a["pred"], b["pred"])
    np.testing.assert_array_equal(a["pred"], c["pred"])
    np.testing.assert_array_equal(a["h_true"], c["h_true"])
    assert a["match"] == b["match"] == int((a["pred"] == st.ans).sum())


def test_edge_cases(qmann, synth, qmo):
    """Single-sentence stories, empty sentences, an all-zero question, a story dense enough to
    overflow the fixed compaction slot (heap path), repeated words and fractional values."""
    cfg = synth.ModelConfig(V=70, d=20, S_max=50, V_dict=20, mode=2)
    w = synth.make_weights(cfg, 21, sigma=0.6)
    rng = np.random.default_rng(5)
    n_sen = np.array([1, 1, 50, 50, 7, 50, 3, 50], dtype=np.uint32)
    st = synth.make_stories(cfg, len(n_sen), 22, n_sen=n_sen)
    off = st.offsets()
    st.m[off[2]:off[2] + 5] = 0.0                       # empty sentences
    st.q[1] = 0.0                                       # empty question
    st.m[off[3]:off[4]] = rng.integers(1, 4, size=(50, cfg.V)).astype(np.float32)        # dense: 3500 entries > slot
    st.m[off[5]:off[6]] *= rng.choice(np.array([0.5, 1.0, 2.0, 3.0, -1.0], dtype=np.float32), size=(50, cfg.V))
    ref = qmo.forward(cfg, w, st)
    got = _fwd(qmann, cfg, w, st)
    safe = ref["risk"] == 0
    for k in ("u0",):
        np.testing.assert_array_equal(got[k], ref[k])
    rows_safe = np.repeat(safe, st.n_sen)
    for k in ("M", "C", "s"):
        np.testing.assert_array_equal(got[k][:, rows_safe], ref[k][:, rows_safe], err_msg=k)
    for k in ("o", "g", "u"):
        np.testing.assert_array_equal(got[k][:, safe], ref[k][:, safe], err_msg=k)
    np.testing.assert_array_equal(got["pred"][safe & (ref["risk_ans"] > 1e-4)], ref["pred"][safe & (ref["risk_ans"] > 1e-4)])
    fast = _fwd(qmann, cfg, w, st, debug=False, want_h=False)
    np.testing.assert_array_equal(fast["pred"], got["pred"])


def test_infer_host_end_to_end(qmann, synth, qmo):
    """Host arenas in, predictions out (qmann_infer_host) equals the device-resident path and the oracle."""
    cfg = synth.preset_config("C4")
    w = synth.make_weights(cfg, 31, sigma=0.5)
    st = synth.make_stories(cfg, 20000, 32, ragged=True)
    model = qmann.lib.Model(cfg, w)
    pred, match, cost = model.infer_host(st.m, st.q, st.a, st.n_sen, want_cost=True)
    dev = _fwd(qmann, cfg, w, st, debug=False, want_h=True)
    np.testing.assert_array_equal(pred, dev["pred"])
    assert match == dev["match"] == int((pred == st.ans).sum())
    sub = synth.Stories(m=st.m[:st.offsets()[300]], q=st.q[:300], a=st.a[:300], n_sen=st.n_sen[:300], ans=st.ans[:300])
    ref = qmo.forward(cfg, w, sub, dump=False)
    ok = (ref["risk"] == 0) & (ref["risk_ans"] > 1e-4)
    np.testing.assert_array_equal(pred[:300][ok], ref["pred"][ok])
    # cost accumulates -h[y] in story order in fp32, like the reference
    acc = np.float32(0.0)
    for v in dev["h_true"]:
        acc = np.float32(np.float64(acc) + -1.0 * np.float64(v))
    assert cost == pytest.approx(float(acc), rel=1e-6)


def test_full_size_properties(qmann, synth):
    """BASELINE config 2 at full size (20 000 stories): idempotence, permutation equivariance of the
    batch, and shards reassembling to the full result (batch sharding needs no collective)."""
    cfg = synth.preset_config("C2")
    w = synth.make_weights(cfg, 41, sigma=0.5)
    st = synth.make_stories(cfg, 20000, 42, S=50)
    a = _fwd(qmann, cfg, w, st, debug=False, want_h=False)
    b = _fwd(qmann, cfg, w, st, debug=False, want_h=False)
    np.testing.assert_array_equal(a["pred"], b["pred"])
    assert a["match"] == int((a["pred"] == st.ans).sum())
    perm = np.random.default_rng(1).permutation(st.N)
    m3 = st.m.reshape(st.N, 50, cfg.V)[perm].reshape(-1, cfg.V)
    stp = synth.Stories(m=m3, q=st.q[perm], a=st.a[perm], n_sen=st.n_sen[perm], ans=st.ans[perm])
    c = _fwd(qmann, cfg, w, stp, debug=False, want_h=False)
    np.testing.assert_array_equal(c["pred"], a["pred"][perm])
    parts = []
    for rank in range(4):
        first, count = qmann.lib.shard_plan(st.n_sen, 4, rank)
        sub = synth.Stories(m=st.m[first * 50:(first + count) * 50], q=st.q[first:first + count], a=st.a[first:first + count],
                            n_sen=st.n_sen[first:first + count], ans=st.ans[first:first + count])
        parts.append(_fwd(qmann, cfg, w, sub, debug=False, want_h=False)["pred"])
    np.testing.assert_array_equal(np.concatenate(parts), a["pred"])


def test_reference_layer_objects_run_on_our_shim(qmann, synth):
    """The reference's OWN lib/layer.c objects linked against libqmann_b200.so (oracle/_ref/
    ref_harness_b200, built by `make -C oracle ref_b200`) reproduce the golden tensors: the cuda_*
    shim is a drop-in for lib/layer_cuda.cu on the inference path."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_harness_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_harness_b200 not built (needs /root/reference at build time)")
    for name in golden_io.case_names():
        cfg, w, st, ref = golden_io.load_case(name, synth)
        with tempfile.TemporaryDirectory() as td:
            case, dump = os.path.join(td, "c.bin"), os.path.join(td, "d.bin")
            synth.write_case(case, cfg, w, st)
            subprocess.run([exe, case, dump], check=True, capture_output=True)
            got = synth.read_dump(dump)
        keys = FIXED_KEYS if cfg.mode != 1 else ("u0", "M", "C", "s")
        for k in keys:
            if k == "g" and not cfg.lin_map:
                continue
            np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{name}: {k}")
        np.testing.assert_allclose(got["p"], ref["p"], rtol=SOFTMAX_RTOL, atol=0)
        np.testing.assert_allclose(got["h"], ref["h"], rtol=SOFTMAX_RTOL, atol=0)
        np.testing.assert_array_equal(got["pred"], ref["pred"], err_msg=name)
        assert int(got["match"]) == int(ref["match"])
        assert np.array_equal(got["p"], ref["p"]), f"{name}: shim softmax not bit-identical"
        if cfg.mode == 1:
            np.testing.assert_array_equal(got["o"], ref["o"], err_msg="mode-1 fp32 read")
            np.testing.assert_array_equal(got["z"], ref["z"])


def test_shim_layers_match_reference_library_live(qmann):
    """Single cuda_* entry points of our library vs the same symbols of the compiled reference
    (oracle/_ref/libqmann_ref.so) on random fp32 inputs, including off-grid values."""
    import ctypes as C
    import torch
    refp = os.path.join(ROOT, "oracle", "_ref", "libqmann_ref.so")
    if not os.path.exists(refp):
        pytest.skip("oracle/_ref/libqmann_ref.so not built")
    R, O = C.CDLL(refp), qmann.lib.lib()
    u32, b, fp = C.c_uint, C.c_bool, C.c_void_p
    sigs = {
        "cuda_dense_fwd": [fp, fp, fp, fp, fp, u32, u32, C.c_char_p, b, u32, u32, u32, u32, u32, b],
        "cuda_dense_mat_fwd": [fp, fp, fp, fp, fp, u32, u32, u32, b, u32, u32, u32, b],
        "cuda_dot_mat_vec_fwd": [fp, fp, fp, fp, u32, u32, b, b, u32, u32, u32, u32, u32, b],
        "cuda_dot_mat_vec_fwd_appx": [fp, fp, fp, fp, fp, u32, u32, b, u32, u32, u32, u32, b, b],
        "cuda_softmax_fwd": [fp, fp, fp, fp, fp, u32, b, b],
        "cuda_sum_vec_fwd": [fp, fp, fp, u32, b, u32, u32, u32, b],
    }
    for L in (R, O):
        for n, s in sigs.items():
            getattr(L, n).argtypes = s
            getattr(L, n).restype = None
    g = torch.Generator(device="cuda").manual_seed(3)
    rnd = lambda *s, sc=4.0: (torch.randn(*s, device="cuda", generator=g) * sc).contiguous()
    V, d, S = 70, 20, 37

    def both(fn):
        outs = []
        for L in (R, O):
            outs.append(fn(L))
            torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), fn.__name__
    Wm, x, M, u, pv = rnd(d, V, sc=1.0), rnd(V), rnd(S, d), rnd(d), torch.rand(S, device="cuda", generator=g)
    xm = rnd(S, V)

    def dense(L):
        o = torch.zeros(d, device="cuda"); L.cuda_dense_fwd(Wm.data_ptr(), None, x.data_ptr(), o.data_ptr(), None, V, d, b"NULL", True, 5, 2, 6, 1, 3, False); return o
    def dense_f(L):
        o = torch.zeros(d, device="cuda"); L.cuda_dense_fwd(Wm.data_ptr(), None, x.data_ptr(), o.data_ptr(), None, V, d, b"NULL", False, 8, 7, 8, 7, 3, False); return o
    def dense_mat(L):
        o = torch.zeros(S, d, device="cuda"); L.cuda_dense_mat_fwd(Wm.data_ptr(), None, xm.data_ptr(), o.data_ptr(), None, V, d, S, True, 4, 3, 3, False); return o
    def score(L):
        o = torch.zeros(S, device="cuda"); L.cuda_dot_mat_vec_fwd(M.data_ptr(), u.data_ptr(), o.data_ptr(), None, S, d, False, True, 5, 2, 5, 2, 3, False); return o
    def read(L):
        o = torch.zeros(d, device="cuda"); L.cuda_dot_mat_vec_fwd(M.data_ptr(), pv.data_ptr(), o.data_ptr(), None, S, d, True, True, 5, 2, 5, 2, 3, False); return o
    def appx(L):
        o = torch.zeros(S, device="cuda"); L.cuda_dot_mat_vec_fwd_appx(M.data_ptr(), u.data_ptr(), o.data_ptr(), None, None, S, d, True, 5, 2, 3, 8, False, False); return o
    def softmax(L):
        o = torch.zeros(S, device="cuda"); mx = torch.zeros(1, device="cuda"); L.cuda_softmax_fwd(o.data_ptr(), M[:, 0].contiguous().data_ptr(), None, None, mx.data_ptr(), S, False, False); return o
    def sumv(L):
        o = torch.zeros(d, device="cuda"); L.cuda_sum_vec_fwd(u.data_ptr(), M[0].contiguous().data_ptr(), o.data_ptr(), d, True, 5, 2, 3, False); return o
    for fn in (dense, dense_f, dense_mat, score, read, appx, softmax, sumv):
        both(fn)


@pytest.mark.parametrize("preset,sigma", [("C2", 0.5), ("C2", 1.3), ("C4", 0.7), ("C1", 2.5)])
def test_packed_path_equals_unpacked_and_instrumented(preset, sigma, qmann, synth, monkeypatch):
    """Production kernels, three ways: packed embedding + scorer (k_forward_fast<.., 2, true>: stories whose rows are
    narrow, the rest falls through), packed path switched off (QMANN_SWAR=0), and the instrumented general kernel.
    Same predictions and match counts on 3000 ragged stories; sigma 1.3 mixes narrow and non-narrow stories, 2.5 has
    none that is narrow."""
    import torch
    cfg = synth.preset_config(preset)
    w = synth.make_weights(cfg, 41, sigma=sigma)
    st = synth.make_stories(cfg, 3000, 42, S=min(cfg.S_max, 50), ragged=True)
    res = {}
    for tag, env in (("packed", "1"), ("unpacked", "0")):
        monkeypatch.setenv("QMANN_SWAR", env)
        model = qmann.lib.Model(cfg, w)
        db = model.upload(st)
        out = model.forward(db, with_answers=True, want_h=False, debug=False)
        torch.cuda.synchronize()
        res[tag] = (out["pred"].cpu().numpy()[:st.N].copy(), int(out["match"].cpu().numpy()[0]))
        if tag == "packed":
            dbg = model.forward(db, with_answers=True, want_h=True, debug=True)
            torch.cuda.synchronize()
            res["general"] = (dbg["pred"].cpu().numpy()[:st.N].copy(), int(dbg["match"].cpu().numpy()[0]))
    for tag in ("unpacked", "general"):
        np.testing.assert_array_equal(res["packed"][0], res[tag][0], err_msg=tag)
        assert res["packed"][1] == res[tag][1]


def test_optional_layer_variants_match_reference_library_live(qmann):
    """the optional, default-off variants on the same cuda_* surface -- scale layer (EN_SC_ATT,
    lib/layer_cuda.cu:4805), activation layer (EN_NON_LINEARITY, :4548; NULL / SIGMOID / RELU, quantised and fp32),
    shift-based softmax (:2038), attention mode 1 (pure fp32 dot products, lib/layer.c:177-195: scorer and transposed
    read with f_fixed = false) -- entry point by entry point against the compiled reference (oracle/_ref/libqmann_ref.so)
    on random fp32 inputs including off-grid, saturating and negative-zero-producing values.  Bit-identical outputs."""
    import ctypes as C
    import torch
    refp = os.path.join(ROOT, "oracle", "_ref", "libqmann_ref.so")
    if not os.path.exists(refp):
        pytest.skip("oracle/_ref/libqmann_ref.so not built")
    R, O = C.CDLL(refp), qmann.lib.lib()
    u32, b, fp = C.c_uint, C.c_bool, C.c_void_p
    sigs = {
        "cuda_scale_fwd": [fp, fp, fp, u32, b, u32, u32, u32, b],
        "cuda_activation_fwd": [fp, fp, C.c_char_p, u32, b, u32, u32, u32],
        "cuda_softmax_fwd": [fp, fp, fp, fp, fp, u32, b, b],
        "cuda_dot_mat_vec_fwd": [fp, fp, fp, fp, u32, u32, b, b, u32, u32, u32, u32, u32, b],
        "cuda_dense_fwd": [fp, fp, fp, fp, fp, u32, u32, C.c_char_p, b, u32, u32, u32, u32, u32, b],
        "cuda_sum_vec_fwd": [fp, fp, fp, u32, b, u32, u32, u32, b],
    }
    for L in (R, O):
        for n, s in sigs.items():
            getattr(L, n).argtypes = s
            getattr(L, n).restype = None
    g = torch.Generator(device="cuda").manual_seed(11)
    rnd = lambda *s, sc=4.0: (torch.randn(*s, device="cuda", generator=g) * sc).contiguous()
    S, d = 53, 37
    checked = []

    def both(name, fn):
        outs = []
        for L in (R, O):
            outs.append(fn(L))
            torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]), f"{name}: max |diff| {(outs[0] - outs[1]).abs().max().item()}"
        assert torch.isfinite(outs[0]).all() or name.startswith("softmax_shift"), name
        checked.append(name)

    x = rnd(S)
    x[::7] = -0.1                                        # truncates to a negative zero in the quantised variants
    x[3], x[4] = 40.0, -40.0                             # saturate (5,2)
    for wv in (0.37, -2.5, 1.0):
        wt = torch.tensor([wv], device="cuda")
        for ff in (True, False):
            def scale(L, wt=wt, ff=ff):
                o = torch.zeros(S, device="cuda"); L.cuda_scale_fwd(x.data_ptr(), wt.data_ptr(), o.data_ptr(), S, ff, 5, 2, 3, False); return o
            both(f"scale w={wv} fixed={ff}", scale)
    for kind in (b"NULL", b"SIGMOID", b"RELU"):
        for ff, iwl, frac in ((True, 5, 2), (True, 2, 5), (False, 5, 2)):
            def act(L, kind=kind, ff=ff, iwl=iwl, frac=frac):
                o = torch.zeros(S, device="cuda"); L.cuda_activation_fwd(x.data_ptr(), o.data_ptr(), kind, S, ff, iwl, frac, 3); return o
            both(f"activation {kind.decode()} fixed={ff} ({iwl},{frac})", act)
    for sc in (1.0, 6.0):
        s_in = (rnd(S, sc=sc) * 4).round() / 4           # scores on the (5,2) grid, as the scorer produces them
        for shift in (False, True):
            def softmax(L, s_in=s_in, shift=shift):
                o = torch.zeros(S, device="cuda"); mx = torch.zeros(1, device="cuda")
                L.cuda_softmax_fwd(o.data_ptr(), s_in.data_ptr(), None, None, mx.data_ptr(), S, shift, False); return o
            both(f"softmax_shift={shift} scale={sc}", softmax)
    # attention mode 1: fp32 scorer and fp32 transposed read (f_fixed = false), and the fp32 update
    M, u, pv = rnd(S, d), rnd(d), torch.rand(S, device="cuda", generator=g)
    def score_f(L):
        o = torch.zeros(S, device="cuda"); L.cuda_dot_mat_vec_fwd(M.data_ptr(), u.data_ptr(), o.data_ptr(), None, S, d, False, False, 5, 2, 5, 2, 3, False); return o
    def read_f(L):
        o = torch.zeros(d, device="cuda"); L.cuda_dot_mat_vec_fwd(M.data_ptr(), pv.data_ptr(), o.data_ptr(), None, S, d, True, False, 5, 2, 5, 2, 3, False); return o
    def sum_f(L):
        o = torch.zeros(d, device="cuda"); L.cuda_sum_vec_fwd(u.data_ptr(), M[1].contiguous().data_ptr(), o.data_ptr(), d, False, 5, 2, 3, False); return o
    both("mode-1 scorer (fp32)", score_f)
    both("mode-1 read (fp32)", read_f)
    both("fp32 update", sum_f)
    # dense layer with an activation name, as EN_NON_LINEARITY wires it (activation argument of cuda_dense_fwd)
    Wm, xv = rnd(d, S, sc=1.0), rnd(S)
    for kind in (b"NULL", b"SIGMOID", b"RELU"):
        def dense_act(L, kind=kind):
            o = torch.zeros(d, device="cuda"); L.cuda_dense_fwd(Wm.data_ptr(), None, xv.data_ptr(), o.data_ptr(), None, S, d, kind, True, 5, 2, 5, 2, 3, False); return o
        both(f"dense activation={kind.decode()}", dense_act)
    assert len(checked) >= 25


@pytest.mark.parametrize("preset,mode,sc,non_lin", [("C1", 2, [0.37, -1.5, 2.25], False), ("C1", 2, None, True), ("C2", 2, [1.0, 0.5, 3.0], True),
                                                    ("C3", 3, [0.8, 1.7, -0.6], True), ("C4", 2, [0.25, 4.0, 1.0], True)])
def test_batched_optional_layers_match_reference_live(preset, mode, sc, non_lin, qmann, synth):
    """the default-off layers of the reference's graph in the BATCHED forward -- the scale layer between
    scorer and softmax (EN_SC_ATT, MemN2N.c:852, 2446, 2647; s' = s * w in fp32) and the RELU activation after the hop
    update (EN_NON_LINEARITY, MemN2N.c:894, 2423, 2670) -- against the reference's own layer objects run live with those
    layers wired in (oracle/ref_harness.c): every tensor of the graph, predictions and match count."""
    from test_gpu_production import REF_EXE, _ref_live
    if not os.path.exists(REF_EXE):
        pytest.skip("oracle/_ref/ref_harness_refcuda not built (needs /root/reference at build time)")
    cfg = synth.preset_config(preset)
    cfg.mode, cfg.sc_att, cfg.non_lin = mode, sc, non_lin
    w = synth.make_weights(cfg, 300 + mode, sigma=0.5)
    st = synth.make_stories(cfg, 600, 301, S=min(cfg.S_max, 50), ragged=True)
    ref = _ref_live(synth, cfg, w, st)
    got = _fwd(qmann, cfg, w, st)
    tag = f"{preset}/mode{mode}/sc={sc}/relu={non_lin}"
    _assert_fixed_equal(got, ref, cfg, tag)
    np.testing.assert_array_equal(got["z"], ref["z"], err_msg=f"{tag}: answer logits")
    np.testing.assert_allclose(got["p"], ref["p"], rtol=SOFTMAX_RTOL, atol=0, err_msg=f"{tag}: attention weights")
    np.testing.assert_array_equal(got["pred"], ref["pred"], err_msg=tag)
    assert got["match"] == int(ref["match"])
    if non_lin:
        assert (got["u"] >= 0).all() and (ref["u"] >= 0).all(), "RELU output"
    # the production entry (no debug struct) takes the same general kernel when an optional layer is on
    plain = _fwd(qmann, cfg, w, st, debug=False, want_h=False)
    np.testing.assert_array_equal(plain["pred"], ref["pred"], err_msg=f"{tag}: production entry")
    assert plain["match"] == int(ref["match"])
    # and the layers change the result (the test would be vacuous otherwise)
    base_cfg = synth.preset_config(preset)
    base_cfg.mode = mode
    base = _fwd(qmann, base_cfg, w, st)
    assert not np.array_equal(base["u"], got["u"])
