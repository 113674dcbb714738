"""CPU oracle (oracle/qmann_oracle.c) against the golden vectors produced by the UNMODIFIED
reference CUDA implementation on a B200 (oracle/gen_golden.py -> tests/golden/).  This is the pin
that makes the oracle trustworthy: every fixed-point tensor must match bit-for-bit; fp32 softmax
values (libm expf here, MUFU.EX2 there) within 1e-5 relative."""
import numpy as np
import pytest

import golden_io

SOFTMAX_RTOL = 1e-5
FIXED_KEYS = ("u0", "M", "C", "s", "o", "g", "u")


@pytest.mark.parametrize("name", golden_io.case_names())
def test_forward_case_matches_reference(name, synth, qmo):
    cfg, w, st, ref = golden_io.load_case(name, synth)
    out = qmo.forward(cfg, w, st)
    assert int((out["risk"] > 0).sum()) == 0, "fixture sits on a truncation boundary; pick another seed"
    for k in FIXED_KEYS:
        if k == "g" and not cfg.lin_map:
            continue
        if cfg.mode == 1 and k in ("o", "u"):
            # ATTENTION_MODE 1 keeps the read in fp32, so it inherits the softmax tolerance
            np.testing.assert_allclose(out[k], ref[k], rtol=2e-5, atol=1e-6, err_msg=k)
            continue
        np.testing.assert_array_equal(out[k], ref[k], err_msg=f"{name}: {k}")
    np.testing.assert_allclose(out["p"], ref["p"], rtol=SOFTMAX_RTOL, atol=1e-30, err_msg="p")
    if cfg.mode == 1:
        np.testing.assert_allclose(out["z"], ref["z"], rtol=2e-5, atol=1e-5)
    else:
        np.testing.assert_array_equal(out["z"], ref["z"], err_msg="answer logits (sequential fp32)")
    np.testing.assert_allclose(out["h"], ref["h"], rtol=SOFTMAX_RTOL, atol=1e-30, err_msg="h")
    np.testing.assert_array_equal(out["pred"], ref["pred"], err_msg="predicted answers")
    assert int(out["match"]) == int(ref["match"])
    np.testing.assert_allclose(float(out["cost"]), float(ref["cost"]), rtol=1e-5)


def test_kat_fixed_mul(qmo):
    """FIXED_MUL + output quantisation on every pair of 8-bit codes (reference scorer with d=1)."""
    k = golden_io.load_kat("fixed_mul")
    L = qmo.lib()
    vals = k["vals"].astype(np.float64)
    for ci, (im, fm, iv, fv) in enumerate(k["fmt_pairs"].tolist()):
        got = np.empty((255, 255), dtype=np.float32)
        for i, a in enumerate(vals / 2.0 ** fm):
            for j, b in enumerate(vals / 2.0 ** fv):
                got[i, j] = L.qmo_quant(L.qmo_fixed_mul(a, b, im, fm, iv, fv), im, fm)
        np.testing.assert_array_equal(got, k["out"][ci], err_msg=f"fmt {(im, fm, iv, fv)}")
        # and the integer closed form the CUDA kernels use
        ai = np.array([L.qmo_int_quant(a, im, fm) for a in vals / 2.0 ** fm])
        bi = np.array([L.qmo_int_quant(b, iv, fv) for b in vals / 2.0 ** fv])
        if im + fm > 0:
            gi = np.array([[L.qmo_int_mul(int(x), int(y), im, fm, fv) for y in bi] for x in ai], dtype=np.float64)
            np.testing.assert_array_equal((gi / 2.0 ** fm).astype(np.float32), k["out"][ci])


def test_kat_requant(qmo):
    k = golden_io.load_kat("requant")
    L = qmo.lib()
    for ci, (isrc, fsrc, im, fm) in enumerate(k["cases"].tolist()):
        got = np.array([L.qmo_quant(L.qmo_fixed_mul(v / 2.0 ** fsrc, 1.0, im, fm, im, fm), im, fm) for v in k["vals"]],
                       dtype=np.float32)
        np.testing.assert_array_equal(got, k["out"][ci])
        gi = np.array([L.qmo_int_mul(L.qmo_int_requant(int(v), fsrc, im, fm), 1 << fm, im, fm, fm) for v in k["vals"]])
        np.testing.assert_array_equal((gi / 2.0 ** fm).astype(np.float32), k["out"][ci])


def test_kat_fixed_add(qmo):
    k = golden_io.load_kat("fixed_add")
    L = qmo.lib()
    vals = k["vals"].astype(np.float64)
    for ci, (i_, f_, fa, fb) in enumerate(k["cases"].tolist()):
        got = np.array([[L.qmo_fixed_add(a, b, i_, f_, i_, f_) for b in vals / 2.0 ** fb] for a in vals / 2.0 ** fa],
                       dtype=np.float32)
        np.testing.assert_array_equal(got, k["out"][ci], err_msg=f"{(i_, f_, fa, fb)}")


def test_kat_appx_element(qmo):
    """Approximate (Hamming) attention element, incl. the saturating encode and the compiled
    sign|(a+b) overflow behaviour, on every pair of 8-bit codes."""
    k = golden_io.load_kat("appx_element")
    L = qmo.lib()
    for ci, (ia, fM, fu) in enumerate(k["cases"].tolist()):
        M = (k["vals"] / 2.0 ** fM).astype(np.float32)
        got = np.empty((255, 255), dtype=np.float32)
        out = np.empty(255, dtype=np.float32)
        for j, uv in enumerate(k["vals"]):
            u = np.array([uv / 2.0 ** fu], dtype=np.float32)
            L.qmo_approximate_attention(qmo._fp(M), qmo._fp(u), qmo._fp(out), 255, 1, ia, 8, -3)
            got[:, j] = out
        np.testing.assert_array_equal(got, k["out"][ci], err_msg=f"{(ia, fM, fu)}")


def test_kat_appx_rows(qmo):
    k = golden_io.load_kat("appx_rows")
    L = qmo.lib()
    for key in [x for x in k if x.startswith("out_")]:
        ia, fM = (int(t) for t in key.split("_")[1:])
        M = np.ascontiguousarray(k["M"] / 2.0 ** fM, dtype=np.float32)
        u = np.ascontiguousarray(k["u"] / 2.0 ** fM, dtype=np.float32)
        out = np.empty(64, dtype=np.float32)
        L.qmo_approximate_attention(qmo._fp(M), qmo._fp(u), qmo._fp(out), 64, 64, ia, 8, -3)
        np.testing.assert_array_equal(out, k[key], err_msg=key)


def test_kat_softmax(qmo):
    k = golden_io.load_kat("softmax")
    L = qmo.lib()
    for dim, x, ref in zip(k["dims"].tolist(), k["inp"], k["out"]):
        xi = np.ascontiguousarray(x[:dim])
        out = np.empty(dim, dtype=np.float32)
        L.qmo_softmax(qmo._fp(xi), qmo._fp(out), dim)
        np.testing.assert_allclose(out, ref[:dim], rtol=SOFTMAX_RTOL, atol=1e-37)
