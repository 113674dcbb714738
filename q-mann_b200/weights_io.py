"""Weight files: the raw fp32 dumps the reference's driver reads and writes in its (commented-out)
weight I/O, MemN2N/MemN2N.c:2553-2618 (load) and :2853-2978 (dump).

Layout of every file: for each hop (files with hops), for each INPUT column j, for each OUTPUT row i: one little-endian
fp32 = w_mat[i][j] -- i.e. the [dim_out][dim_in] matrix of the layer struct stored transposed, hops back to back:
    w_emb_a_float.bin   emb_m[h].w_mat  [d][V]  x H      (A tables)
    w_emb_c_float.bin   emb_c[h].w_mat  [d][V]  x H      (C tables)
    w_emb_q_float.bin   emb_q.w_mat     [d][V]           (B table)
    w_float.bin         ds_ans.w_mat    [V][d]           (answer projection W)
The reference never dumped the linear map (lin_map[h].w_mat [d][d]); it is written here, in the same convention, to
    w_lin_map_float.bin                 [d][d]  x H
when it is absent, loading fails unless `require_lin_map` is False (then Hm is returned as zeros and the caller must run
the model with lin_map = False, the reference's EN_LINEAR_MAPPING off).
"""
from __future__ import annotations

import os
from typing import List

import numpy as np

from .synth import ModelConfig, Weights

FILES = dict(A="w_emb_a_float.bin", C="w_emb_c_float.bin", B="w_emb_q_float.bin", W="w_float.bin", Hm="w_lin_map_float.bin")


def _dump(path: str, mats: List[np.ndarray]) -> None:
    with open(path, "wb") as fh:
        for m in mats:
            fh.write(np.ascontiguousarray(np.asarray(m, dtype="<f4").T).tobytes())      # for j (in) for i (out): w[i][j]


def _load(path: str, n: int, dim_out: int, dim_in: int) -> List[np.ndarray]:
    want = n * dim_out * dim_in * 4
    size = os.path.getsize(path)
    if size != want:
        raise ValueError(f"{path}: {size} bytes, expected {want} ({n} x [{dim_out}][{dim_in}] fp32)")
    raw = np.fromfile(path, dtype="<f4").reshape(n, dim_in, dim_out)
    return [np.ascontiguousarray(raw[k].T) for k in range(n)]


def save_weights(directory: str, cfg: ModelConfig, w: Weights) -> None:
    os.makedirs(directory, exist_ok=True)
    assert len(w.A) == cfg.H and len(w.C) == cfg.H and w.B.shape == (cfg.d, cfg.V) and w.W.shape == (cfg.V, cfg.d)
    _dump(os.path.join(directory, FILES["A"]), w.A)
    _dump(os.path.join(directory, FILES["C"]), w.C)
    _dump(os.path.join(directory, FILES["B"]), [w.B])
    _dump(os.path.join(directory, FILES["W"]), [w.W])
    if cfg.lin_map:
        _dump(os.path.join(directory, FILES["Hm"]), w.Hm)


def load_weights(directory: str, cfg: ModelConfig, require_lin_map: bool = True) -> Weights:
    A = _load(os.path.join(directory, FILES["A"]), cfg.H, cfg.d, cfg.V)
    C = _load(os.path.join(directory, FILES["C"]), cfg.H, cfg.d, cfg.V)
    B = _load(os.path.join(directory, FILES["B"]), 1, cfg.d, cfg.V)[0]
    W = _load(os.path.join(directory, FILES["W"]), 1, cfg.V, cfg.d)[0]
    hp = os.path.join(directory, FILES["Hm"])
    if os.path.exists(hp):
        Hm = _load(hp, cfg.H, cfg.d, cfg.d)
    elif require_lin_map and cfg.lin_map:
        raise FileNotFoundError(f"{hp}: the linear-map weights are missing (the reference never dumped them)")
    else:
        Hm = [np.zeros((cfg.d, cfg.d), np.float32) for _ in range(cfg.H)]
    return Weights(B=B, A=A, C=C, Hm=Hm, W=W)
