"""The PRODUCTION kernels observed (round 2): k_story (packed and unpacked tiers, dense bulk-copy stream) and the general
kernel for what they decline, run with their dump instantiations through the C ABI (qmann_debug.production = 1).

north_star's bar is bit-exact fixed-point / Hamming scores, selected slots and predicted answers.  Here the tensors the
production launch itself computed -- per-hop scores s, the code of Q_f(p) of every slot (non-zero = selected slot of the
weighted read), read o, linear map g, controller state u, the exact logits of the prefilter's candidate rows and the
prediction -- are compared with
  * the golden tensors of the unmodified reference (tests/golden, made on a B200 by oracle/gen_golden.py),
  * the reference itself run LIVE on this box at >= 2000 stories per shape (oracle/_ref/ref_harness_refcuda),
  * the instrumented general kernel on seeded inputs,
plus adversarial answer-prefilter inputs and a weight format narrower than a byte (count splitting)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "ref_harness_refcuda")


def _run(qmann, cfg, w, st, production_dump=False, debug=False, want_h=False, model=None):
    import torch
    model = model or qmann.lib.Model(cfg, w)
    db = model.upload(st)
    model.path_counts()
    out = model.forward(db, with_answers=True, want_h=want_h, debug=debug, production_dump=production_dump)
    torch.cuda.synchronize()
    assert model.check_errors() == 0
    res = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    res["pred"] = res["pred"][:st.N].astype(np.uint32)
    res["match"] = int(res["match"][0])
    res["tiers"] = model.path_counts()
    return res


def _pcode_from_p(p, cfg):
    """Q_f(p) of the reference's attention weights (lib/layer_cuda.cu:561): trunc(p * 2^frac) saturated to the format."""
    f = cfg.formats()
    out = np.zeros(p.shape, dtype=np.uint8)
    for h in range(cfg.H):
        ff, lim = f["frac"][h], (1 << (f["iwl"][h] + f["frac"][h])) - 1
        out[h] = np.minimum(np.trunc(p[h].astype(np.float64) * (1 << ff)), lim).astype(np.uint8)
    return out


def _compare_production(got, ref, cfg, st, tag, z_ref=None):
    assert set(np.unique(got["path"])) <= {1, 2, 3}, f"{tag}: every story must be finished by one tier"
    np.testing.assert_array_equal(got["u0"], ref["u0"], err_msg=f"{tag}: u0")
    np.testing.assert_array_equal(got["s"], ref["s"], err_msg=f"{tag}: scores")
    np.testing.assert_array_equal(got["pcode"], _pcode_from_p(ref["p"], cfg), err_msg=f"{tag}: selected slots / Q_f(p) codes")
    for k in ("o", "u") + (("g",) if cfg.lin_map else ()):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=f"{tag}: {k}")
    zr = ref["z"] if z_ref is None else z_ref
    cand = got["cand"] > 0
    assert cand.any(axis=1).all(), f"{tag}: every story has at least one exact logit"
    np.testing.assert_array_equal(got["z"][cand], zr[cand], err_msg=f"{tag}: exact logits of the candidate rows")
    # the reference's predicted row must be among the rows whose logit was computed
    assert cand[np.arange(st.N), ref["pred"].astype(np.int64)].all(), f"{tag}: prefilter dropped the winning row"
    np.testing.assert_array_equal(got["pred"], ref["pred"], err_msg=f"{tag}: predictions")


@pytest.mark.parametrize("name", [n for n in golden_io.case_names() if n != "c1_mode1"])
def test_production_kernels_match_reference_golden(name, qmann, synth):
    cfg, w, st, ref = golden_io.load_case(name, synth)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, name)
    assert got["match"] == int(ref["match"])


def _ref_live(synth, cfg, w, st):
    with tempfile.TemporaryDirectory() as td:
        case, dump = os.path.join(td, "c.bin"), os.path.join(td, "d.bin")
        synth.write_case(case, cfg, w, st)
        subprocess.run([REF_EXE, case, dump], check=True, capture_output=True, timeout=900)
        return synth.read_dump(dump)


LIVE = [("C1", 2, 0.5, 2400), ("C2", 2, 0.5, 2000), ("C2", 2, 1.3, 2000), ("C3", 3, 0.5, 2000), ("C4", 2, 0.5, 2000), ("C4", 3, 0.7, 2000),
        ("C1", 3, 1.0, 2400)]


@pytest.mark.parametrize("preset,mode,sigma,n", LIVE)
def test_production_kernels_match_reference_live(preset, mode, sigma, n, qmann, synth):
    """>= 2000 seeded ragged stories per shape through the UNMODIFIED reference (its own layer.c + layer_cuda.cu compiled for
    sm_100a) on this GPU, against the production launch: identical scores, selected slots, read, state, logits, predictions,
    match count.  sigma 1.3 mixes the packed and the unpacked tier."""
    if not os.path.exists(REF_EXE):
        pytest.skip("oracle/_ref/ref_harness_refcuda not built (needs /root/reference at build time)")
    cfg = synth.preset_config(preset, mode=mode) if False else synth.preset_config(preset)
    cfg.mode = mode
    w = synth.make_weights(cfg, 900 + n + mode, sigma=sigma)
    st = synth.make_stories(cfg, n, 901 + mode, S=min(cfg.S_max, 50), ragged=True)
    # a few irregular stories so that the general tier takes part: fractional values, a large count, an empty sentence
    off = st.offsets()
    st.m[off[5], 3] = 0.5
    st.m[off[17] + (st.n_sen[17] - 1), 4] = 9.0
    st.m[off[33]] = 0.0
    ref = _ref_live(synth, cfg, w, st)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, f"{preset}/mode{mode}/sigma{sigma}")
    assert got["match"] == int(ref["match"])
    assert got["path"][5] == 3 and got["tiers"][2] >= 1, "the fractional story belongs to the general tier"
    # the plain production launch (no dump instantiation, 24 warps) predicts the same
    plain = _run(qmann, cfg, w, st)
    np.testing.assert_array_equal(plain["pred"], ref["pred"])
    assert plain["match"] == int(ref["match"])


@pytest.mark.parametrize("preset,sigma", [("C1", 0.5), ("C2", 0.5), ("C2", 1.3), ("C3", 0.5), ("C4", 0.5), ("C4", 2.5)])
def test_production_dump_equals_instrumented_kernel(preset, sigma, qmann, synth):
    """Seeded: 3000 ragged stories, production tiers vs the instrumented general kernel, tensor by tensor."""
    cfg = synth.preset_config(preset)
    w = synth.make_weights(cfg, 51, sigma=sigma)
    st = synth.make_stories(cfg, 3000, 52, S=min(cfg.S_max, 50), ragged=True)
    ref = _run(qmann, cfg, w, st, debug=True, want_h=True)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, f"{preset}/{sigma}")
    assert got["match"] == ref["match"]
    if preset == "C2" and sigma == 0.5:
        assert got["tiers"][0] == st.N and got["tiers"][1] < st.N // 2, "C2 bench weights: the packed tier takes (almost) every story"


@pytest.mark.parametrize("fast_softmax", ["1", "0"])
def test_attention_codes_with_ties_at_the_top(fast_softmax, qmann, synth, monkeypatch):
    """Stories whose top scores tie exactly (p = 1/2, 1/3, 1/4: the truncation boundaries of Q_f) -- the float-total
    shortcut must hand them to the exact double total.  Built from repeated sentences: identical rows score identically."""
    monkeypatch.setenv("QMANN_FAST_SOFTMAX", fast_softmax)
    cfg = synth.preset_config("C2")
    w = synth.make_weights(cfg, 61, sigma=0.5)
    st = synth.make_stories(cfg, 600, 62, S=12, ragged=False)
    off = st.offsets()
    Vd = cfg.V_dict
    for i in range(st.N):
        k = 2 + i % 3                                    # 2, 3 or 4 identical word bags (time columns differ: zero them)
        rows = slice(off[i], off[i] + k)
        st.m[rows, :Vd] = st.m[off[i], :Vd]
    # time-column weights zero => identical bags give identical memory rows and exactly tied scores
    for t in w.A + w.C:
        t[:, Vd:] = 0.0
    ref = _run(qmann, cfg, w, st, debug=True, want_h=True)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, f"ties/fast_softmax={fast_softmax}")
    tied = 0
    for i in range(st.N):
        s0 = ref["s"][0, off[i]:off[i + 1]]
        tied += int((s0 == s0.max()).sum() >= 2)
    assert tied > 50, "the construction must produce ties at the top"


def _adversarial_W(kind, cfg, rng):
    V, d = cfg.V, cfg.d
    if kind == "duplicate_rows":
        base = (rng.standard_normal((8, d)) * 0.5).astype(np.float32)
        return base[np.arange(V) % 8].copy()
    if kind == "one_ulp_apart":
        base = (rng.standard_normal((8, d)) * 0.5).astype(np.float32)
        W = base[np.arange(V) % 8].copy()
        up = np.nextafter(W, np.float32(np.inf))
        W[1::2] = up[1::2]
        return W
    if kind == "heavy_tailed":
        W = (rng.standard_normal((V, d)) * 0.01).astype(np.float32)
        W[rng.integers(0, V, 6), rng.integers(0, d, 6)] = 100.0
        W[rng.integers(0, V, 6), rng.integers(0, d, 6)] = -100.0
        return W
    if kind == "tiny":
        return (rng.standard_normal((V, d)) * 1e-6).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["duplicate_rows", "one_ulp_apart", "heavy_tailed", "tiny", "zero_u"])
def test_answer_prefilter_adversarial(kind, qmann, synth):
    """The int8 prefilter of the answer projection against inputs built to break its bound: exact duplicates (probability ties ->
    last-index rule), rows one ulp apart, heavy-tailed W (scale max|W|/127 collapses most of W8 to 0), tiny W, all-zero u."""
    cfg = synth.preset_config("C2")
    rng = np.random.default_rng(7)
    w = synth.make_weights(cfg, 71, sigma=0.5)
    if kind == "zero_u":
        for t in [w.B] + w.A + w.C + w.Hm:
            t[...] = 0.0
    else:
        w.W = _adversarial_W(kind, cfg, rng)
    st = synth.make_stories(cfg, 1500, 72, S=20, ragged=True)
    ref = _run(qmann, cfg, w, st, debug=True, want_h=True)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, kind)
    plain = _run(qmann, cfg, w, st)
    np.testing.assert_array_equal(plain["pred"], ref["pred"])
    if kind == "zero_u":
        assert (ref["pred"] == cfg.V - 1).all(), "all logits equal: argmax_last picks the last row"
    if os.path.exists(REF_EXE):
        sub = synth.Stories(m=st.m[:st.offsets()[400]], q=st.q[:400], a=st.a[:400], n_sen=st.n_sen[:400], ans=st.ans[:400])
        live = _ref_live(synth, cfg, w, sub)
        np.testing.assert_array_equal(got["pred"][:400], live["pred"], err_msg=f"{kind}: vs the reference run live")
        np.testing.assert_array_equal(ref["z"][:400], live["z"])


def test_narrow_word_length_count_splitting(qmann, synth, qmo):
    """Weight formats narrower than a byte (BW_WL = 6: limits 31): a word repeated n times must saturate per PRODUCT like the
    reference, Q_w(Q_w(n) * Q_w(T)) (lib/layer_cuda.cu:120), not per row -- n unit entries are only legal while n * max|code|
    stays inside every hop's format."""
    cfg = synth.ModelConfig(V=70, d=20, S_max=50, V_dict=20, mode=2, wl=6, iwl=3)
    w = synth.make_weights(cfg, 81, sigma=3.0)              # large codes: 2 * |code| exceeds 31 often
    st = synth.make_stories(cfg, 800, 82, S=30, ragged=True, min_words=4, max_words=9)      # V_dict = 20: many repeats
    assert (st.m[:, :cfg.V_dict] >= 2).sum() > 500
    ref = qmo.forward(cfg, w, st)
    dbg = _run(qmann, cfg, w, st, debug=True, want_h=True)
    got = _run(qmann, cfg, w, st, production_dump=True)
    safe = ref["risk"] == 0
    rows = np.repeat(safe, st.n_sen)
    np.testing.assert_array_equal(dbg["M"][:, rows], ref["M"][:, rows], err_msg="instrumented M vs oracle")
    np.testing.assert_array_equal(dbg["s"][:, rows], ref["s"][:, rows])
    np.testing.assert_array_equal(got["s"][:, rows], ref["s"][:, rows], err_msg="production scores vs oracle")
    for k in ("o", "g", "u"):
        np.testing.assert_array_equal(got[k][:, safe], ref[k][:, safe], err_msg=k)
    _compare_production(got, dbg, cfg, st, "wl6")
    if os.path.exists(REF_EXE):
        sub = synth.Stories(m=st.m[:st.offsets()[300]], q=st.q[:300], a=st.a[:300], n_sen=st.n_sen[:300], ans=st.ans[:300])
        live = _ref_live(synth, cfg, w, sub)
        n_rows = int(st.offsets()[300])
        np.testing.assert_array_equal(got["s"][:, :n_rows], live["s"])
        np.testing.assert_array_equal(got["u"][:, :300], live["u"])
        np.testing.assert_array_equal(got["pred"][:300], live["pred"])


def test_error_flag_is_reported_and_cleared(qmann, synth):
    """A story that overflows the compaction heap poisons neither the flag nor later calls."""
    import torch
    cfg = synth.ModelConfig(V=70, d=20, S_max=50, V_dict=20, mode=2)
    w = synth.make_weights(cfg, 91, sigma=0.5)
    st = synth.make_stories(cfg, 64, 92, S=20)
    model = qmann.lib.Model(cfg, w)
    pred, match, _ = model.infer_host(st.m, st.q, st.a, st.n_sen)
    pred2, match2, _ = model.infer_host(st.m, st.q, st.a, st.n_sen)
    np.testing.assert_array_equal(pred, pred2)
    assert model.check_errors() == 0
    torch.cuda.synchronize()
