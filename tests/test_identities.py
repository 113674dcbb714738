"""The integer identities the packed / tensor-core / nine-bit kernels rest on, checked exhaustively on the CPU.

Each kernel that replaces the reference's per-product arithmetic by something cheaper does so through an identity on
8-bit codes; here every identity is evaluated on ALL operand pairs (or bytes) in numpy and compared with the pinned
oracle (`qmo_fixed_mul`, itself pinned to the reference by tests/golden/kat_fixed_mul.npz) or with the reference's
known-answer table (kat_appx_element.npz).  No GPU."""
import numpy as np

import golden_io

M32 = np.uint64(0xFFFFFFFF)
H = np.uint64(0x80808080)
V = np.arange(-127, 128, dtype=np.int64)


def _trunc0_div(x, sh):
    return np.sign(x) * (np.abs(x) >> sh)


def test_truncated_product_floor_mod_and_sign_magnitude_forms(qmo):
    """k_big_scores_fast / swar_score: 4*trunc0(x/4) = x - (x mod 4) + 4*[x<0 and x mod 4 != 0]  (floor-mod), with
    x mod 4 from the two low bits of y and u; k_big_scores_mma: 4*trunc0(x/4) = y*u - sum_a I_a(y)*R_a(u).  Both against
    the oracle's FIXED_MUL on every pair of codes of format (5,2), where no product saturates."""
    y, u = np.meshgrid(V, V, indexing="ij")
    x = y * u
    t = _trunc0_div(x, 2)
    # the oracle (float interface): Q_(5,2)(Q(y/4) * Q(u/4)) * 4, restricted to |x| < 512 (no saturation of the product)
    L = qmo.lib()
    sub = V[::9]
    for a in sub:
        for b in sub:
            if abs(a * b) < 512:
                assert L.qmo_fixed_mul(a / 4.0, b / 4.0, 5, 2, 5, 2) * 4.0 == float(_trunc0_div(np.int64(a * b), 2))
    # floor-mod form with the bit formulas of the kernels
    y0, y1, u0, u1 = y & 1, (y >> 1) & 1, u & 1, (u >> 1) & 1
    xm = (y0 & u0) + 2 * ((y1 & u0) ^ (y0 & u1))
    assert np.array_equal(xm, x & 3)
    neg = ((y < 0) ^ (u < 0)) & (xm != 0)
    assert np.array_equal(4 * t, x - xm + 4 * neg)
    # byte c = (x mod 4) + 3 with bit 2 cleared when negative-and-nonzero: c = xm + 3 - 4*neg, in 0..6
    c = (xm + 3) & ~np.where((y < 0) ^ (u < 0), 4, 0)
    assert np.array_equal(c, xm + 3 - 4 * neg) and c.min() >= 0 and c.max() <= 6
    # sign-magnitude form: indicator planes of y, rho planes of u
    rho = lambda a, b: (a * b) & 3
    acc = y * u
    for a in (1, 2, 3):
        Ia = np.sign(y) * ((np.abs(y) & 3) == a)
        Ra = np.sign(u) * rho(a, np.abs(u) & 3)
        assert np.abs(Ra).max() <= 3
        acc = acc - Ia * Ra
    assert np.array_equal(acc, 4 * t)
    # saturation of Q_att (la = 127) happens exactly when |y*u| >= 512
    assert np.array_equal(np.abs(t) > 127, np.abs(x) >= 512)


def _pack(b):
    b = np.asarray(b).astype(np.int64) & 0xFF
    return (b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16) | (b[..., 3] << 24)).astype(np.uint64)


def _unpack_signed(w):
    w = np.asarray(w).astype(np.uint64)
    b = np.stack([(w >> np.uint64(8 * i)) & np.uint64(0xFF) for i in range(4)], -1).astype(np.int64)
    return np.where(b >= 128, b - 256, b)


def test_swar_byte_helpers():
    """swar_unbias, the two Q_att shifts and the exact per-byte saturation test of k_forward_fast's packed path, on every
    byte value in every byte lane (neighbouring lanes filled with extreme values to expose carries)."""
    rng = np.random.default_rng(3)
    for lane in range(4):
        for B in range(0, 128):
            a = rng.integers(-B, B + 1, size=(64, 4)) if B else np.zeros((64, 4), np.int64)
            a[:, lane] = rng.integers(-B, B + 1, size=64) if B else 0
            a[0, :] = B; a[1, :] = -B; a[2, lane] = B; a[3, lane] = -B
            s = _pack(a + B)
            Bw = np.uint64(B * 0x01010101)
            r = (((s | H) - Bw) & M32) ^ ((~s) & H & M32)
            assert np.array_equal(_unpack_signed(r), a)
    allb = np.stack(np.meshgrid(V[::2], V[::5], indexing="ij"), -1).reshape(-1, 2)
    a4 = np.concatenate([allb, allb[:, ::-1]], axis=1)                      # every value in lanes 0/3 (and 1/2)
    w = _pack(a4)
    # ka = -1: trunc0(a / 2)
    neg = (w >> np.uint64(7)) & np.uint64(0x01010101)
    t = (((w & np.uint64(0x7F7F7F7F)) + neg) ^ (w & H)) & M32
    y = ((t >> np.uint64(1)) & np.uint64(0x7F7F7F7F)) | (t & H)
    assert np.array_equal(_unpack_signed(y), np.trunc(a4 / 2).astype(np.int64))
    # ka = +1 (|a| <= 63): 2a
    small = a4[(np.abs(a4) <= 63).all(axis=1)]
    ws = _pack(small)
    assert np.array_equal(_unpack_signed(((ws << np.uint64(1)) & np.uint64(0xFEFEFEFE)) & M32), 2 * small)
    # saturation test: byte |y| + (0x80 - tau(|u|)) has bit 7 set iff |y*u| >= 512
    u4 = np.roll(a4, 1, axis=0)
    tau = np.where(np.abs(u4) <= 4, 128, -(-512 // np.maximum(np.abs(u4), 1)))
    fill = _pack(np.where(a4 < 0, -1, 0))
    absy = (((w ^ fill) & M32) + (fill & np.uint64(0x01010101))) & M32
    flag = (absy + _pack(128 - tau)) & M32 & H
    assert np.array_equal(_unpack_signed(flag) != 0, np.abs(a4 * u4) >= 512)


def test_hamming_element_in_nine_bits():
    """Mode 3: the reference's 31-bit sign-magnitude element equals the closed form on A = sat9(code << (k-23)) for every
    pair of 8-bit codes and every format of the known-answer table (produced by the reference's CUDA code)."""
    k = golden_io.load_kat("appx_element")
    vals = k["vals"].astype(np.int64)

    def enc9(n, frac_n, ia):
        sh = 8 - ia - frac_n
        assert sh >= 0
        t = n * (1 << sh)
        return np.where(t == -256, 0, np.clip(t, -255, 255))

    for ci, (ia, fM, fu) in enumerate(k["cases"].tolist()):
        A, U = enc9(vals, fM, ia), enc9(vals, fu, ia)
        w = np.abs(A[:, None] - U[None, :])
        e = (~(w >> 1)) & 0x7F
        neg = ((vals[:, None] ^ vals[None, :]) < 0) & (w < 256)
        got = np.where(neg, -e, e).astype(np.float64) / 1024.0
        np.testing.assert_array_equal(got, k["out"][ci].astype(np.float64), err_msg=str((ia, fM, fu)))


def test_hamming_element_on_packed_bytes():
    """k_big_scores_ham: the nine-bit element on magnitude bytes.  With a = |A_m|, b = |A_u| in 0..255 and the two sign flags,
    e + 127 = 254 - (|a - b| >> 1) when the signs agree, and with m = (a + b) >> 1 (a byte: the halving add) it is m for m < 128 and
    382 - m otherwise when they differ -- every step a per-byte operation without carries between bytes.  Also the per-byte
    encoder a = sat8(|code| << sh) with the -2^iwl value mapped to 0.  All 8-bit code pairs, shifts 0..4."""
    codes = np.arange(-128, 128, dtype=np.int64)
    for sh_m in range(0, 5):
        for sh_u in range(0, 5):
            def enc9(n, sh):
                t = n * (1 << sh)
                return np.where(t == -256, 0, np.clip(t, -255, 255))
            A, U = enc9(codes, sh_m), enc9(codes, sh_u)
            w = np.abs(A[:, None] - U[None, :])
            e = (~(w >> 1)) & 0x7F
            neg = ((codes[:, None] ^ U[None, :]) < 0) & (w < 256)
            want = np.where(neg, -e, e)
            # byte encoder as the kernel does it on four packed bytes
            def enc_bytes(n, sh):
                mag = np.abs(n)                                            # <= 128
                thr = 256 >> sh
                ov = ((mag + (128 - thr)) & 0x80) != 0 if thr <= 128 else np.zeros_like(mag, bool)
                a = np.where(ov, 255, (mag << sh) & 0xFF)
                a = np.where((n < 0) & (mag == thr), 0, a)                 # code << sh == -256 encodes to 0, the sign stays
                return a
            a, b = enc_bytes(codes, sh_m), enc_bytes(codes, sh_u)
            assert np.array_equal(a, np.abs(A)) and np.array_equal(b, np.abs(U))
            sm, su = codes < 0, U < 0                                      # query sign: of the nine-bit value (0 is positive)
            aa, bb = a[:, None], b[None, :]
            xs = 254 - (np.abs(aa - bb) >> 1)
            m = (aa & bb) + (((aa ^ bb) & 0xFE) >> 1)
            assert np.array_equal(m, (aa + bb) >> 1)
            mask = np.where(m & 0x80, 0xFF, 0)
            xd = (m ^ mask) + (mask & 0x7F)
            assert xd.max() <= 254 and xs.min() >= 0
            x = np.where(sm[:, None] ^ su[None, :], xd, xs)
            np.testing.assert_array_equal(x - 127, want, err_msg=str((sh_m, sh_u)))


def test_answer_prefilter_bound():
    """k_forward_fast's answer prefilter: with W8 = rint(W/s), s = max|W|/127 and u = n/2^fu, the fp32 logit computed in
    the reference's order satisfies |z * 2^fu / s - D| <= |n|_1 * (1/2 + 127 * gamma_{d+1}) for D = W8 . n, so the row of
    the largest logit is always inside the candidate window used by the kernel."""
    rng = np.random.default_rng(11)
    for d, V_, sigma, fu in [(50, 256, 0.5, 2), (64, 114, 0.1, 2), (20, 70, 2.0, 3), (256, 64, 0.3, 1)]:
        W = (rng.standard_normal((V_, d)) * sigma).astype(np.float32)
        s = float(np.abs(W).max()) / 127.0
        W8 = np.rint(W.astype(np.float64) / s).astype(np.int64)
        for _ in range(40):
            n = np.clip(np.rint(rng.standard_normal(d) * rng.choice([2, 10, 40])), -127, 127).astype(np.int64)
            u = (n / float(1 << fu)).astype(np.float32)
            z = np.zeros(V_, np.float32)
            for j in range(d):                                   # sequential fp32: fl(z + fl(w*u))
                z = (z + (W[:, j] * u[j]).astype(np.float32)).astype(np.float32)
            D = W8 @ n
            n1 = int(np.abs(n).sum())
            E = n1 * (0.5 + 127 * (d + 2) * 2.0 ** -24)
            assert np.all(np.abs(z.astype(np.float64) * (1 << fu) / s - D) <= E + 1e-9)
            T = n1 + (n1 >> 6) + int(np.ceil(1e-5 * (1 << fu) / s)) + 3
            best = int(np.argmax(z))
            assert D[best] >= D.max() - T
            # every row within 1e-5 of the best logit is a candidate
            near = np.nonzero(z >= z.max() - 1e-5)[0]
            assert np.all(D[near] >= D.max() - T)
