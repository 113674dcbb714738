"""The tensor-core tier k_story_tc (QMANN_TC=1, q-mann_b200/csrc/qmann_tcstory.cuh): the sentence embedding formed as the
reference's dense product X * A^T (lib/layer_cuda.cu:105-141) by tcgen05.mma kind::tf32 from TMA-staged dense rows.  The
claim is exactness: what the tier computed (scores, Q_f(p) codes = selected slots, read, linear map, state, candidate
logits, predictions) must be bit-identical to the instrumented general kernel, to the golden tensors of the unmodified
reference, and to the reference run live on this box."""
import os

import numpy as np
import pytest

import golden_io
from test_gpu_production import REF_EXE, _compare_production, _ref_live, _run

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc_on(monkeypatch):
    monkeypatch.setenv("QMANN_TC", "1")        # read when a model is created


def _tc_model(qmann, cfg, w):
    return qmann.lib.Model(cfg, w)


@pytest.mark.parametrize("sigma,n,S", [(0.5, 3000, 50), (0.5, 997, 64), (1.3, 2000, 50), (0.5, 1200, 9)])
def test_tc_tier_equals_instrumented_kernel(sigma, n, S, tc_on, qmann, synth):
    """Ragged C2-shaped stories (S up to 64: two accumulator tiles; S <= 32: one), sigma 1.3 sends rows whose column
    maxima exceed a byte to the next tier, a few irregular stories go to the general kernel."""
    cfg = synth.preset_config("C2")
    w = synth.make_weights(cfg, 31, sigma=sigma)
    st = synth.make_stories(cfg, n, 32, S=S, ragged=True)
    off = st.offsets()
    st.m[off[5], 3] = 0.5                              # fractional value: general tier
    st.m[off[17] + (st.n_sen[17] - 1), 4] = 9.0        # a count too large to split
    st.m[off[21], 7] = 2.0                             # a repeated word: stays in the tier when 2 * max|code| fits
    st.q[40, 11] = 3.0
    st.m[off[33]] = 0.0                                # an empty sentence
    ref = _run(qmann, cfg, w, st, debug=True, want_h=True)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, f"tc/sigma{sigma}/S{S}")
    assert got["match"] == ref["match"]
    assert got["path"][5] == 3, "the fractional story belongs to the general tier"
    if sigma == 0.5:
        assert (got["path"] == 1).sum() > 0.9 * n, "bench-like weights: the tensor-core tier finishes almost every story"
    plain = _run(qmann, cfg, w, st)                    # the plain instantiation (no dumps)
    np.testing.assert_array_equal(plain["pred"], ref["pred"])
    assert plain["match"] == ref["match"]


def test_tc_tier_is_taken_and_serves_every_output(tc_on, qmann, synth):
    """The tier is really taken: with QMANN_TC=1 a C2 forward launches fewer kernels than the two-kernel path (no k_compact
    over the chunk) and reports every story in tier 0; want_h and with_answers=False variants predict the same."""
    import torch
    cfg = synth.preset_config("C2")
    w = synth.make_weights(cfg, 33, sigma=0.5)
    st = synth.make_stories(cfg, 4000, 34, S=50)
    model = qmann.lib.Model(cfg, w)
    db = model.upload(st)
    model.path_counts()
    a = model.forward(db, with_answers=True, want_h=True)
    torch.cuda.synchronize()
    pred_a, h_a = a["pred"].cpu().numpy()[:st.N].copy(), a["h_true"].cpu().numpy()[:st.N].copy()
    assert model.path_counts()[0] == st.N
    b = model.forward(db, with_answers=False)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(b["pred"].cpu().numpy()[:st.N], pred_a)
    ref = model.forward(db, with_answers=True, want_h=True, debug=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ref["pred"].cpu().numpy()[:st.N], pred_a)
    np.testing.assert_array_equal(ref["h_true"].cpu().numpy()[:st.N], h_a)       # h[y]: same double total, same division


@pytest.mark.parametrize("name", ["c2_mode2"])
def test_tc_tier_matches_reference_golden(name, tc_on, qmann, synth):
    """The golden C2 case (every tensor from the unmodified reference on a B200) through the tensor-core tier."""
    if name not in golden_io.case_names():
        pytest.skip(f"no golden case {name}")
    cfg, w, st, ref = golden_io.load_case(name, synth)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, f"tc/{name}")


def test_tc_tier_matches_reference_live(tc_on, qmann, synth):
    if not os.path.exists(REF_EXE):
        pytest.skip("oracle/_ref/ref_harness_refcuda not built (needs /root/reference at build time)")
    cfg = synth.preset_config("C2")
    w = synth.make_weights(cfg, 35, sigma=0.5)
    st = synth.make_stories(cfg, 2000, 36, S=50, ragged=True)
    ref = _ref_live(synth, cfg, w, st)
    got = _run(qmann, cfg, w, st, production_dump=True)
    _compare_production(got, ref, cfg, st, "tc/live")
    assert got["match"] == int(ref["match"])
    assert (got["path"] == 1).all()
