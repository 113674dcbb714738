"""ctypes binding of oracle/libqmann_oracle.so (the CPU restatement, qmann_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqmann_oracle.so")
QMO_MAX_HOP = 8


class _Fmt(C.Structure):
    _fields_ = [("iwl", C.c_uint32), ("frac", C.c_uint32)]


_FP = C.POINTER(C.c_float)


class _Model(C.Structure):
    _fields_ = [("V", C.c_uint32), ("d", C.c_uint32), ("H", C.c_uint32), ("mode", C.c_uint32),
                ("lin_map", C.c_uint32), ("f_fixed", C.c_uint32), ("const_scale", C.c_int32),
                ("fmt", _Fmt * QMO_MAX_HOP), ("fmt_w", _Fmt * QMO_MAX_HOP), ("fmt_att", _Fmt * QMO_MAX_HOP),
                ("fmt_bin", _Fmt),
                ("B", _FP), ("A", _FP * QMO_MAX_HOP), ("C", _FP * QMO_MAX_HOP), ("Hm", _FP * QMO_MAX_HOP),
                ("W", _FP)]


class _Dump(C.Structure):
    _fields_ = [(k, _FP) for k in ("u0", "M", "C", "s", "p", "o", "g", "u", "z", "h")] + \
               [("pred", C.POINTER(C.c_uint32)), ("h_true", _FP), ("risk", _FP), ("risk_ans", _FP)]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile) if it is not built yet."""
    src = os.path.join(_HERE, "qmann_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.qmo_float2fixed.restype = C.c_uint32
        L.qmo_float2fixed.argtypes = [C.c_double, C.c_uint32, C.c_uint32]
        L.qmo_fixed2float.restype = C.c_float
        L.qmo_fixed2float.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
        L.qmo_quant.restype = C.c_double
        L.qmo_quant.argtypes = [C.c_double, C.c_uint32, C.c_uint32]
        L.qmo_fixed_mul.restype = C.c_float
        L.qmo_fixed_mul.argtypes = [C.c_float, C.c_float] + [C.c_uint32] * 4
        L.qmo_fixed_add.restype = C.c_float
        L.qmo_fixed_add.argtypes = [C.c_float, C.c_float] + [C.c_uint32] * 4
        L.qmo_int_quant.restype = C.c_int32
        L.qmo_int_quant.argtypes = [C.c_double, C.c_uint32, C.c_uint32]
        L.qmo_int_requant.restype = C.c_int32
        L.qmo_int_requant.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.qmo_int_mul.restype = C.c_int32
        L.qmo_int_mul.argtypes = [C.c_int32, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.qmo_appx_element.restype = C.c_float
        L.qmo_appx_element.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32]
        L.qmo_softmax.restype = None
        L.qmo_softmax.argtypes = [_FP, _FP, C.c_uint32]
        L.qmo_expf_shifted.restype = None
        L.qmo_expf_shifted.argtypes = [_FP, C.c_float, _FP, C.c_uint32]
        L.qmo_argmax_last.restype = C.c_uint32
        L.qmo_argmax_last.argtypes = [_FP, C.c_uint32]
        L.qmo_mat_vec_product.restype = None
        L.qmo_mat_vec_product.argtypes = [_FP, _FP, _FP, C.c_uint32, C.c_uint32, C.c_int, _Fmt, _Fmt]
        L.qmo_mat_mat_trans_product.restype = None
        L.qmo_mat_mat_trans_product.argtypes = [_FP, _FP, _FP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                                _Fmt, _Fmt, _Fmt]
        L.qmo_mat_trans_mat_product.restype = None
        L.qmo_mat_trans_mat_product.argtypes = [_FP, _FP, _FP, C.c_uint32, C.c_uint32, C.c_int, _Fmt]
        L.qmo_approximate_attention.restype = None
        L.qmo_approximate_attention.argtypes = [_FP, _FP, _FP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_int32]
        L.qmo_vec_vec_sum.restype = None
        L.qmo_vec_vec_sum.argtypes = [_FP, _FP, _FP, C.c_uint32, C.c_int, _Fmt]
        L.qmo_forward.restype = C.c_uint32
        L.qmo_forward.argtypes = [C.POINTER(_Model), _FP, _FP, _FP, C.POINTER(C.c_uint32), C.c_uint32,
                                  C.POINTER(_Dump), _FP, C.c_int]
        L.qmo_max_threads.restype = C.c_int
    return _lib


def fmt(iwl: int, frac: int) -> _Fmt:
    return _Fmt(iwl, frac)


def _fp(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_FP)


def make_model(cfg, w) -> _Model:
    """cfg: ModelConfig, w: Weights (q-mann_b200/synth.py).  Keeps numpy arrays alive on the struct."""
    f = cfg.formats()
    m = _Model()
    m.V, m.d, m.H, m.mode = cfg.V, cfg.d, cfg.H, cfg.mode
    m.lin_map, m.f_fixed, m.const_scale = int(cfg.lin_map), int(cfg.f_fixed), cfg.const_scale
    keep = []
    for h in range(cfg.H):
        m.fmt[h] = _Fmt(f["iwl"][h], f["frac"][h])
        m.fmt_w[h] = _Fmt(f["iwl_w"][h], f["frac_w"][h])
        m.fmt_att[h] = _Fmt(f["iwl_att"][h], f["frac_att"][h])
        for name, lst in (("A", w.A), ("C", w.C), ("Hm", w.Hm)):
            arr = np.ascontiguousarray(lst[h], dtype=np.float32)
            keep.append(arr)
            getattr(m, name)[h] = _fp(arr)
    m.fmt_bin = _Fmt(f["iwl_bin"], f["frac_bin"])
    B = np.ascontiguousarray(w.B, dtype=np.float32)
    W = np.ascontiguousarray(w.W, dtype=np.float32)
    keep += [B, W]
    m.B, m.W = _fp(B), _fp(W)
    m._keep = keep
    return m


def forward(cfg, w, st, dump: bool = True, n_threads: int = 0, with_answers: bool = True) -> Dict[str, np.ndarray]:
    """Run the CPU oracle over all stories.  Returns the dump tensors (if dump) plus pred/match/cost."""
    L = lib()
    mdl = make_model(cfg, w)
    N, H, d, V, ss = st.N, cfg.H, cfg.d, cfg.V, st.sum_sen
    out: Dict[str, np.ndarray] = {}
    dp = _Dump()
    shapes = dict(u0=(N, d), M=(H, ss, d), C=(H, ss, d), s=(H, ss), p=(H, ss), o=(H, N, d), g=(H, N, d),
                  u=(H, N, d), z=(N, V), h=(N, V))
    if not dump:
        shapes = {}
    for k, shp in shapes.items():
        out[k] = np.zeros(shp, dtype=np.float32)
        setattr(dp, k, _fp(out[k]))
    out["pred"] = np.zeros(N, dtype=np.uint32)
    dp.pred = out["pred"].ctypes.data_as(C.POINTER(C.c_uint32))
    for k in ("h_true", "risk", "risk_ans"):
        out[k] = np.zeros(N, dtype=np.float32)
        setattr(dp, k, _fp(out[k]))
    cost = C.c_float(0.0)
    m = np.ascontiguousarray(st.m, dtype=np.float32)
    q = np.ascontiguousarray(st.q, dtype=np.float32)
    a = np.ascontiguousarray(st.a, dtype=np.float32) if with_answers else None
    ns = np.ascontiguousarray(st.n_sen, dtype=np.uint32)
    match = L.qmo_forward(C.byref(mdl), _fp(m), _fp(q), _fp(a), ns.ctypes.data_as(C.POINTER(C.c_uint32)),
                          N, C.byref(dp), C.byref(cost), n_threads)
    out["match"] = np.uint32(match)
    out["cost"] = np.float32(cost.value)
    return out
