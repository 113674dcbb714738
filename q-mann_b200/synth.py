"""Model configuration and synthetic bAbI-shaped inputs for the quantized MemN2N forward.

Everything here is host-side numpy.  The shapes, value distributions and seeds follow SURVEY.md
section 8(d); the per-hop fixed-point formats follow the reference driver
(MemN2N/MemN2N.c:714-775: iwl/frac from argv, EN_MQ skew on hop 0 and hop 2 weight formats).

Layouts are the reference's own boundary formats (MemN2N/MemN2N.c:2294-2350, lib/layer.c):
  m  [sum n_sen][V]  fp32 dense bag-of-words rows (word counts + one-hot time column), ragged
  q  [N][V]          fp32 question bag-of-words
  a  [N][V]          fp32 one-hot answer
  B  [d][V], A_h/C_h [d][V], Hm_h [d][d], W [V][d]   fp32 weights, row-major [dim_out][dim_in]
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Dict, List, Optional

import numpy as np

BW_WL = 8                      # MemN2N/define.h:21
ATTENTION_CONST_SCALE = -3     # MemN2N/define.h:67
MAX_HOP = 8


@dataclasses.dataclass
class ModelConfig:
    V: int                      # dim_input = dictionary + time columns
    d: int                      # dim_emb
    S_max: int                  # max_line (memory slots)
    H: int = 3                  # NUM_HOP, define.h:254
    mode: int = 2               # ATTENTION_MODE: 1 float dot, 2 fixed dot, 3 approximate (Hamming)
    lin_map: bool = True        # EN_LINEAR_MAPPING, define.h:291
    f_fixed: bool = True        # EN_FIXED_POINT, define.h:31
    const_scale: int = ATTENTION_CONST_SCALE
    iwl: int = 5                # argv[4]; run.sh:18 passes 5 => base format (5,2)
    en_mq: bool = True          # EN_MQ, define.h:79
    V_dict: int = 0             # dictionary size (incl. NULL at 0); time columns are V_dict..V-1
    wl: int = BW_WL             # word length BW_WL (define.h:21); < 8 gives formats narrower than a byte
    sc_att: Optional[List[float]] = None   # EN_SC_ATT (define.h:58): the scale layer's weight per hop, None = layer absent
    non_lin: bool = False       # EN_NON_LINEARITY (define.h:294): RELU activation layer after every hop update

    def formats(self) -> Dict[str, List[int]]:
        """Per-hop (iwl, frac) arrays exactly as MemN2N.c:714-775 computes them."""
        frac = self.wl - 1 - self.iwl
        iwl = [self.iwl] * self.H
        fr = [frac] * self.H
        iwl_w, frac_w = list(iwl), list(fr)
        if self.en_mq and self.H >= 3:
            iwl_w[0] += 1
            frac_w[0] -= 1
            iwl_w[2] -= 1
            frac_w[2] += 1
        return dict(iwl=iwl, frac=fr, iwl_w=iwl_w, frac_w=frac_w, iwl_att=list(iwl), frac_att=list(fr),
                    iwl_bin=self.iwl, frac_bin=frac)


# BASELINE.json configs as concrete shapes (SURVEY.md section 8 header)
PRESETS = {
    "C1": dict(V_dict=20, S=50, d=20, N=1000, mode=2),
    "C2": dict(V_dict=192, S=50, S_max=64, d=50, N=20000, mode=2),
    "C3": dict(V_dict=192, S=50, S_max=64, d=50, N=20000, mode=3),
    "C4": dict(V_dict=64, S=50, d=64, N=1 << 16, mode=2),
}


def preset_config(name: str, **over) -> ModelConfig:
    p = dict(PRESETS[name])
    p.update(over)
    s_max = p.get("S_max", p["S"])
    return ModelConfig(V=p["V_dict"] + s_max, d=p["d"], S_max=s_max, mode=p["mode"], V_dict=p["V_dict"],
                       **{k: v for k, v in p.items() if k in ("H", "lin_map", "iwl", "en_mq", "f_fixed")})


@dataclasses.dataclass
class Weights:
    B: np.ndarray
    A: List[np.ndarray]
    C: List[np.ndarray]
    Hm: List[np.ndarray]
    W: np.ndarray


def make_weights(cfg: ModelConfig, seed: int, sigma: float = 0.1, tied: bool = True) -> Weights:
    """Gaussian weights (the reference initialises with gaussian_random(0, 0.1), lib/layer.c:1738).
    tied=True mirrors TYPE_WEIGHT_TYING 2 (layer-wise: one A, one C, one Hm shared by all hops,
    MemN2N/define.h:287); the NULL word column is zeroed like ZEROING_NULL_WEIGHT does
    (MemN2N.c:1821-1851).  sigma > 0.1 stands in for trained weights so that the int8 images are
    not almost all zero."""
    rng = np.random.default_rng(seed)
    f32 = np.float32

    def g(*shape):
        return (rng.standard_normal(shape) * sigma).astype(f32)

    B = g(cfg.d, cfg.V)
    nA = 1 if tied else cfg.H
    A = [g(cfg.d, cfg.V) for _ in range(nA)]
    C = [g(cfg.d, cfg.V) for _ in range(nA)]
    Hm = [g(cfg.d, cfg.d) for _ in range(nA)]
    for t in [B] + A + C:
        t[:, 0] = 0.0
    if tied:
        A, C, Hm = A * cfg.H, C * cfg.H, Hm * cfg.H
    W = g(cfg.V, cfg.d)
    return Weights(B=B, A=A, C=C, Hm=Hm, W=W)


@dataclasses.dataclass
class Stories:
    m: np.ndarray          # [sum_sen, V] fp32
    q: np.ndarray          # [N, V] fp32
    a: np.ndarray          # [N, V] fp32 one-hot
    n_sen: np.ndarray      # [N] uint32
    ans: np.ndarray        # [N] uint32 answer id

    @property
    def N(self) -> int:
        return int(self.n_sen.shape[0])

    @property
    def sum_sen(self) -> int:
        return int(self.m.shape[0])

    def offsets(self) -> np.ndarray:
        off = np.zeros(self.N + 1, dtype=np.int64)
        np.cumsum(self.n_sen, out=off[1:])
        return off


def make_stories(cfg: ModelConfig, N: int, seed: int, S: Optional[int] = None, ragged: bool = False,
                 min_words: int = 2, max_words: int = 6, q_words: int = 3,
                 n_sen: Optional[np.ndarray] = None) -> Stories:
    """bAbI-shaped stories: each sentence is a bag of 2..6 word ids uniform in [1, V_dict) plus the
    one-hot time column V_dict + n_sen-1-j (MemN2N/sample.c:466-476, 544-548); words are drawn
    with replacement, so counts of 2 occur (SURVEY hard part 5)."""
    rng = np.random.default_rng(seed)
    S = cfg.S_max if S is None else S
    assert S <= cfg.S_max and cfg.V_dict >= 2 and cfg.V_dict + cfg.S_max <= cfg.V
    if n_sen is None:
        n_sen = rng.integers(1, S + 1, size=N) if ragged else np.full(N, S)
    n_sen = np.asarray(n_sen).astype(np.uint32)
    assert n_sen.shape == (N,) and int(n_sen.max(initial=0)) <= cfg.S_max
    tot = int(n_sen.sum())
    m = np.zeros((tot, cfg.V), dtype=np.float32)
    nw = rng.integers(min_words, max_words + 1, size=tot)
    rows = np.repeat(np.arange(tot), nw)
    ids = rng.integers(1, cfg.V_dict, size=int(nw.sum()))
    np.add.at(m, (rows, ids), 1.0)
    off = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(n_sen, out=off[1:])
    story_of_row = np.repeat(np.arange(N), n_sen)
    j = np.arange(tot) - off[story_of_row]
    m[np.arange(tot), cfg.V_dict + n_sen[story_of_row].astype(np.int64) - 1 - j] = 1.0
    q = np.zeros((N, cfg.V), dtype=np.float32)
    qi = rng.integers(1, cfg.V_dict, size=(N, q_words))
    np.add.at(q, (np.repeat(np.arange(N), q_words), qi.ravel()), 1.0)
    ans = rng.integers(1, cfg.V_dict, size=N).astype(np.uint32)
    a = np.zeros((N, cfg.V), dtype=np.float32)
    a[np.arange(N), ans] = 1.0
    return Stories(m=m, q=q, a=a, n_sen=n_sen, ans=ans)


@dataclasses.dataclass
class IdStories:
    """The same stories as word-id lists (what MemN2N/sample.c holds before sample_vectorization builds the dense
    arenas): rows are story-major, question row first, then the story's sentences."""
    ids: np.ndarray        # [n_ids] uint16
    row_off: np.ndarray    # [N + sum_sen + 1] uint32
    ans: np.ndarray        # [N] uint32
    n_sen: np.ndarray      # [N] uint32

    @property
    def N(self) -> int:
        return int(self.n_sen.shape[0])


def ids_from_dense(st: Stories) -> IdStories:
    """Word-id lists whose scatter (every occurrence adds 1.0, MemN2N/sample.c:547-568) reproduces the dense arenas
    exactly; the arenas must hold non-negative integer counts."""
    for t in (st.m, st.q):
        assert np.all(t >= 0) and np.all(t == np.rint(t)), "dense values must be integer counts"
    N, V = st.N, st.q.shape[1]
    off = st.offsets()
    # row order: story i -> question, then sentences off[i] .. off[i+1]
    n_rows = N + st.sum_sen
    src = np.empty(n_rows, dtype=np.int64)          # >= 0: sentence row index; < 0: question of story -1 - v
    first = off[:-1] + np.arange(N)
    src[first] = -1 - np.arange(N)
    mask = np.ones(n_rows, dtype=bool)
    mask[first] = False
    src[mask] = np.arange(st.sum_sen)
    dense = np.empty((n_rows, V), dtype=np.int64)
    dense[first] = st.q.astype(np.int64)
    dense[mask] = st.m.astype(np.int64)
    r, c = np.nonzero(dense)
    rep = dense[r, c]
    ids = np.repeat(c, rep).astype(np.uint16)
    per_row = np.bincount(np.repeat(r, rep), minlength=n_rows)
    row_off = np.zeros(n_rows + 1, dtype=np.uint32)
    np.cumsum(per_row, out=row_off[1:])
    return IdStories(ids=ids, row_off=row_off, ans=st.ans.astype(np.uint32), n_sen=st.n_sen.astype(np.uint32))


# ---------------------------------------------------------------------------------------------
# case / dump files exchanged with oracle/ref_harness.c
# ---------------------------------------------------------------------------------------------
def write_case(path: str, cfg: ModelConfig, w: Weights, st: Stories) -> None:
    f = cfg.formats()
    with open(path, "wb") as fo:
        fo.write(b"QMNCASE1")
        fo.write(struct.pack("<8Ii", cfg.V, cfg.d, cfg.S_max, cfg.H, st.N, cfg.mode, int(cfg.lin_map),
                             int(cfg.f_fixed), cfg.const_scale))
        for key in ("iwl", "frac", "iwl_w", "frac_w", "iwl_att", "frac_att"):
            fo.write(np.asarray(f[key], dtype="<u4").tobytes())
        fo.write(struct.pack("<3I", f["iwl_bin"], f["frac_bin"], st.sum_sen))
        fo.write(np.ascontiguousarray(w.B, dtype="<f4").tobytes())
        for lst in (w.A, w.C, w.Hm):
            for t in lst:
                fo.write(np.ascontiguousarray(t, dtype="<f4").tobytes())
        fo.write(np.ascontiguousarray(w.W, dtype="<f4").tobytes())
        fo.write(np.ascontiguousarray(st.n_sen, dtype="<u4").tobytes())
        for t in (st.m, st.q, st.a):
            fo.write(np.ascontiguousarray(t, dtype="<f4").tobytes())
        if cfg.sc_att is not None or cfg.non_lin:
            # optional-layer extension read by oracle/ref_harness.c when present
            fo.write(b"QMNEXT01")
            fo.write(struct.pack("<I", 1 if cfg.sc_att is not None else 0))
            fo.write(np.asarray(cfg.sc_att if cfg.sc_att is not None else [0.0] * cfg.H, dtype="<f4").tobytes())
            fo.write(struct.pack("<I", 1 if cfg.non_lin else 0))


def read_dump(path: str) -> Dict[str, np.ndarray]:
    buf = open(path, "rb").read()
    assert buf[:8] == b"QMNDUMP1", "bad dump magic"
    N, H, d, V, sum_sen = struct.unpack_from("<5I", buf, 8)
    pos = 28
    out: Dict[str, np.ndarray] = {}

    def take(name, shape, dt="<f4"):
        nonlocal pos
        n = int(np.prod(shape)) if len(shape) else 1
        out[name] = np.frombuffer(buf, dtype=dt, count=n, offset=pos).reshape(shape).copy()
        pos += n * 4

    take("u0", (N, d)); take("M", (H, sum_sen, d)); take("C", (H, sum_sen, d))
    take("s", (H, sum_sen)); take("p", (H, sum_sen))
    take("o", (H, N, d)); take("g", (H, N, d)); take("u", (H, N, d))
    take("z", (N, V)); take("h", (N, V)); take("pred", (N,), "<u4")
    take("cost", ()); take("match", (), "<u4")
    assert pos == len(buf), "dump size mismatch"
    return out
