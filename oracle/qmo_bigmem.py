"""CPU restatement of the slot-sharded large-memory path (BASELINE config 5), phase by phase.

TEST INFRASTRUCTURE ONLY (same rule as qmo.py).  Every arithmetic step is delegated to the functions of
qmann_oracle.c that are pinned against the reference's golden tensors (scorer, weighted read, linear map,
update, answer projection: lib/layer_cuda.cu:49-172, 355-635, 1535-1542); this module only adds the
sharding protocol around them:

    bins  = scores(...)                     per shard         (one score bin per slot)
    hist  = histogram(bins)                 per shard   --->  all-reduce SUM
    pbin, pq = softmax_from_hist(hist)      identical on every shard: max and the double total are rebuilt
                                            from the global integer histogram in a FIXED order (bins ascending
                                            in 256 contiguous ranges, range partials added ascending), so the
                                            result does not depend on the number of shards
    part  = partial_read(...)               per shard   --->  all-reduce SUM (int32)
    u'    = update(sum of parts)            identical on every shard

For S <= 1024 the only difference from the reference's softmax (lib/layer_cuda.cu:1969-2060, total summed in
slot order) is that order of the double-precision additions, i.e. <= ~1e-16 relative on `total`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np

import qmo

SOFTMAX_RANGES = 256


def _fmts(cfg, h):
    f = cfg.formats()
    return dict(fw=(f["iwl_w"][h], f["frac_w"][h]), fa=(f["iwl_att"][h], f["frac_att"][h]), ff=(f["iwl"][h], f["frac"][h]),
                fb=(f["iwl_bin"], f["frac_bin"]))


def _lim(fmt):
    return (1 << (fmt[0] + fmt[1])) - 1


def num_bins(cfg) -> int:
    f = cfg.formats()
    if cfg.mode == 3:
        return 2 * 127 * cfg.d + 1
    return 2 * max(_lim((f["iwl_att"][h], f["frac_att"][h])) for h in range(cfg.H)) + 1


def bias(cfg, h) -> int:
    return 127 * cfg.d if cfg.mode == 3 else _lim(_fmts(cfg, h)["fa"])


def scores(cfg, h: int, M8: np.ndarray, u8: np.ndarray, fu: int) -> np.ndarray:
    """Score bin of every slot of this shard for one query.  M8 [S][d] int8 codes (weight format of hop h)."""
    L = qmo.lib()
    F = _fmts(cfg, h)
    S, d = M8.shape
    if S == 0:
        return np.zeros(0, dtype=np.int64)
    Mf = np.ascontiguousarray(M8.astype(np.float32) / np.float32(1 << F["fw"][1]))
    uf = np.ascontiguousarray(u8.astype(np.float32) / np.float32(1 << fu))
    s = np.zeros(S, dtype=np.float32)
    if cfg.mode == 3:
        ia, fa = F["fa"]
        L.qmo_approximate_attention(qmo._fp(Mf), qmo._fp(uf), qmo._fp(s), S, d, ia, 1 + ia + fa, cfg.const_scale)
        sh = 7 - cfg.const_scale
        assert np.all(np.abs(s) < (1 << ia)), "saturating mode-3 scores: the raw sum cannot be recovered from the value"
        n = np.rint(s.astype(np.float64) * (1 << sh)).astype(np.int64)
        assert np.array_equal((n / float(1 << sh)).astype(np.float32), s)
        return n + bias(cfg, h)
    L.qmo_mat_mat_trans_product(qmo._fp(Mf), qmo._fp(uf), qmo._fp(s), S, 1, d, 1, qmo.fmt(*F["fa"]), qmo.fmt(*F["fb"]), qmo.fmt(*F["fa"]))
    n = np.rint(s.astype(np.float64) * (1 << F["fa"][1])).astype(np.int64)
    return n + bias(cfg, h)


def histogram(bins: np.ndarray, NB: int) -> np.ndarray:
    return np.bincount(bins, minlength=NB).astype(np.uint32)


def bin_values(cfg, h: int, NB: int) -> np.ndarray:
    """fp32 score value of every bin (what the softmax sees)."""
    F = _fmts(cfg, h)
    n = np.arange(NB, dtype=np.int64) - bias(cfg, h)
    if cfg.mode == 3:
        v = (n.astype(np.float32) / np.float32(1 << (7 - cfg.const_scale))).astype(np.float32)
        lim = np.float32(1 << F["fa"][0])
        return np.where(v >= lim, lim, np.where(v < -lim, -lim, np.where(v == -lim, np.float32(0), v))).astype(np.float32)
    return (n.astype(np.float32) / np.float32(1 << F["fa"][1])).astype(np.float32)


def softmax_from_hist(cfg, h: int, hist: np.ndarray):
    """-> (pbin fp32 [NB], pq uint8 [NB], risk): attention weight and its Q_f code per score bin, from the GLOBAL
    histogram; risk = number of non-empty bins whose code could flip under the libm-vs-MUFU exp difference."""
    L = qmo.lib()
    F = _fmts(cfg, h)
    NB = hist.shape[0]
    v = bin_values(cfg, h, NB)
    nz = hist > 0
    mx = np.float32(v[nz].max())
    e = np.zeros(NB, dtype=np.float32)
    L.qmo_expf_shifted(qmo._fp(np.ascontiguousarray(v)), C.c_float(float(mx)), qmo._fp(e), NB)
    per = (NB + SOFTMAX_RANGES - 1) // SOFTMAX_RANGES
    total = 0.0
    parts: List[float] = []
    for r in range(SOFTMAX_RANGES):
        acc = 0.0
        for b in range(min(NB, r * per), min(NB, r * per + per)):
            if hist[b]:
                acc += float(hist[b]) * float(e[b])
        parts.append(acc)
    for a in parts:
        total += a
    pbin = np.zeros(NB, dtype=np.float32)
    pq = np.zeros(NB, dtype=np.uint8)
    risk = 0
    iff, ff = F["ff"]
    idx = np.nonzero(nz)[0]
    # worst-case difference between libm expf and the device's ex2.approx path, as in qmann_oracle.c:softmax_flip_risk
    xs = v[idx].astype(np.float64) - float(mx)
    dl = np.where(xs == 0.0, 0.0, 3e-7 + 1.5e-7 * np.abs(xs))
    ev = e[idx].astype(np.float64)
    tl = float(np.sum(hist[idx] * ev * (1.0 - dl)))
    th = float(np.sum(hist[idx] * ev * (1.0 + dl)))
    for k, b in enumerate(idx):
        p = np.float32(float(e[b]) / total)
        pbin[b] = p
        pq[b] = L.qmo_float2fixed(float(p), iff, ff) & 0x7FFFFFFF
        lo = np.float32(ev[k] * (1.0 - dl[k]) / th)
        hi = np.float32(ev[k] * (1.0 + dl[k]) / tl)
        if L.qmo_float2fixed(float(lo), iff, ff) != L.qmo_float2fixed(float(hi), iff, ff):
            risk += 1
    return pbin, pq, risk


def partial_read(cfg, h: int, C8: np.ndarray, bins: np.ndarray, pq: np.ndarray) -> np.ndarray:
    """int32 [d]: sum over this shard's slots of Q_f(Q_f(p) * Q_f(C[r][c])) in units of 2^-frac."""
    L = qmo.lib()
    F = _fmts(cfg, h)
    d = C8.shape[1]
    out = np.zeros(d, dtype=np.int64)
    if bins.shape[0] == 0:
        return out.astype(np.int32)
    code = pq[bins]
    for r in np.nonzero(code)[0]:
        for c in range(d):
            c_f = L.qmo_int_requant(int(C8[r, c]), F["fw"][1], F["ff"][0], F["ff"][1])
            out[c] += L.qmo_int_mul(int(code[r]), c_f, F["ff"][0], F["ff"][1], F["ff"][1])
    return out.astype(np.int32)


def update(cfg, w, h: int, partial_sum: np.ndarray, u8: np.ndarray, fu: int):
    """-> (u' int8 codes with frac[h] fractional bits, o codes, g codes).  Literal layer functions."""
    L = qmo.lib()
    F = _fmts(cfg, h)
    d = cfg.d
    lf = _lim(F["ff"])
    o_code = np.clip(partial_sum.astype(np.int64), -lf, lf)
    o = np.ascontiguousarray(o_code.astype(np.float32) / np.float32(1 << F["ff"][1]))
    uf = np.ascontiguousarray(u8.astype(np.float32) / np.float32(1 << fu))
    g = np.zeros(d, dtype=np.float32)
    if cfg.lin_map:
        Hm = np.ascontiguousarray(w.Hm[h], dtype=np.float32)
        L.qmo_mat_vec_product(qmo._fp(Hm), qmo._fp(uf), qmo._fp(g), d, d, 1, qmo.fmt(*F["fw"]), qmo.fmt(*F["fb"]))
        g_frac = F["fw"][1]
    else:
        g[:] = uf
        g_frac = fu
    un = np.zeros(d, dtype=np.float32)
    L.qmo_vec_vec_sum(qmo._fp(g), qmo._fp(o), qmo._fp(un), d, 1, qmo.fmt(*F["ff"]))
    to_code = lambda x, fr: np.rint(x.astype(np.float64) * (1 << fr)).astype(np.int8)
    return to_code(un, F["ff"][1]), o_code.astype(np.int8), to_code(g, g_frac)


def answer(cfg, w, u8: np.ndarray, fu: int):
    L = qmo.lib()
    V, d = cfg.V, cfg.d
    uf = np.ascontiguousarray(u8.astype(np.float32) / np.float32(1 << fu))
    W = np.ascontiguousarray(w.W, dtype=np.float32)
    z = np.zeros(V, dtype=np.float32)
    hh = np.zeros(V, dtype=np.float32)
    L.qmo_mat_vec_product(qmo._fp(W), qmo._fp(uf), qmo._fp(z), V, d, 0, qmo.fmt(0, 0), qmo.fmt(0, 0))
    L.qmo_softmax(qmo._fp(z), qmo._fp(hh), V)
    return z, hh, int(L.qmo_argmax_last(qmo._fp(hh), V))


def check_read_against_literal(cfg, h, C8, bins, pbin, partial_sum):
    """The integer partial sums, clamped, equal the literal weighted read over ALL slots
    (_cuda_mat_trans_mat_product, lib/layer_cuda.cu:547-579)."""
    L = qmo.lib()
    F = _fmts(cfg, h)
    S, d = C8.shape
    Cf = np.ascontiguousarray(C8.astype(np.float32) / np.float32(1 << F["fw"][1]))
    p = np.ascontiguousarray(pbin[bins].astype(np.float32))
    o = np.zeros(d, dtype=np.float32)
    L.qmo_mat_trans_mat_product(qmo._fp(p), qmo._fp(Cf), qmo._fp(o), S, d, 1, qmo.fmt(*F["ff"]))
    lf = _lim(F["ff"])
    ref = np.clip(partial_sum.astype(np.int64), -lf, lf).astype(np.float32) / np.float32(1 << F["ff"][1])
    assert np.array_equal(o, ref), "integer partial read differs from the literal weighted read"


def forward(cfg, w, M8: np.ndarray, C8: np.ndarray, u0: np.ndarray, shards: int = 1, verify_literal: bool = False,
            allreduce=None) -> Dict[str, np.ndarray]:
    """Whole large-memory forward.  M8, C8 [H][S][d] int8; u0 [Q][d] int8 (hop-0 weight format).
    shards > 1 splits the slots into contiguous ranges and merges exactly like the multi-GPU path
    (histogram sum, partial-read sum); `allreduce(np.ndarray) -> np.ndarray` replaces the in-process sum when
    this process holds only ONE shard (M8/C8 are then the local shard)."""
    H, S, d = M8.shape
    Q = u0.shape[0]
    NB = num_bins(cfg)
    f = cfg.formats()
    out = dict(u=np.zeros((H, Q, d), np.int8), o=np.zeros((H, Q, d), np.int8), g=np.zeros((H, Q, d), np.int8),
               bins=np.zeros((H, Q, S), np.int64), pbin=np.zeros((H, Q, NB), np.float32), risk=np.zeros(Q, np.int64),
               nsel=np.zeros((H, Q), np.int64), pred=np.zeros(Q, np.uint32), z=np.zeros((Q, cfg.V), np.float32),
               h=np.zeros((Q, cfg.V), np.float32), hist=np.zeros((H, Q, NB), np.uint32))
    edges = np.linspace(0, S, shards + 1).astype(np.int64)
    for q in range(Q):
        u8, fu = u0[q].copy(), f["frac_w"][0]
        for h in range(H):
            bins = [scores(cfg, h, M8[h, a:b], u8, fu) for a, b in zip(edges[:-1], edges[1:])]
            hist = sum(histogram(b_, NB).astype(np.int64) for b_ in bins)
            if allreduce is not None:
                hist = allreduce(hist)
            hist = hist.astype(np.uint32)
            pbin, pq, risk = softmax_from_hist(cfg, h, hist)
            parts = [partial_read(cfg, h, C8[h, a:b], b_, pq) for (a, b), b_ in zip(zip(edges[:-1], edges[1:]), bins)]
            psum = sum(p_.astype(np.int64) for p_ in parts)
            if allreduce is not None:
                psum = allreduce(psum)
            allb = np.concatenate(bins)
            if verify_literal and allreduce is None:
                check_read_against_literal(cfg, h, C8[h], allb, pbin, psum)
            u8, o8, g8 = update(cfg, w, h, psum, u8, fu)
            fu = f["frac"][h]
            out["u"][h, q], out["o"][h, q], out["g"][h, q] = u8, o8, g8
            out["bins"][h, q], out["pbin"][h, q], out["hist"][h, q] = allb, pbin, hist
            out["nsel"][h, q] = int((pq[allb] != 0).sum())
            out["risk"][q] += risk
        if cfg.V:
            out["z"][q], out["h"][q], out["pred"][q] = answer(cfg, w, u8, fu)
    return out
