"""ctypes binding of libqmann_b200.so (C ABI: include/qmann_abi.h).

torch is used only as the owner of device memory and streams: every call below passes raw
device pointers (tensor.data_ptr()) into the C ABI.  There is no CPU fallback: if the CUDA library
is not built this module raises at load time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqmann_b200.so")
MAX_HOP = 8

_FP = C.c_void_p
_U32 = C.c_uint32


class QConfig(C.Structure):
    _fields_ = [("V", _U32), ("d", _U32), ("S_max", _U32), ("H", _U32), ("mode", _U32), ("lin_map", _U32),
                ("const_scale", C.c_int32),
                ("iwl", _U32 * MAX_HOP), ("frac", _U32 * MAX_HOP), ("iwl_w", _U32 * MAX_HOP), ("frac_w", _U32 * MAX_HOP),
                ("iwl_att", _U32 * MAX_HOP), ("frac_att", _U32 * MAX_HOP), ("iwl_bin", _U32), ("frac_bin", _U32),
                ("en_sc_att", _U32), ("sc_att_w", C.c_float * MAX_HOP), ("en_non_lin", _U32)]


class QWeights(C.Structure):
    _fields_ = [("dev_B", _FP), ("dev_A", _FP * MAX_HOP), ("dev_C", _FP * MAX_HOP), ("dev_Hm", _FP * MAX_HOP), ("dev_W", _FP)]


class QDebug(C.Structure):
    _fields_ = [(f"dev_{k}", _FP) for k in ("u0", "M", "C", "s", "p", "o", "g", "u", "z", "h", "pcode", "path", "cand")] + \
               [("production", _U32)]


def build(force: bool = False, extra: str = "") -> str:
    """Compile libqmann_b200.so for sm_100a with the committed recipe (csrc/Makefile)."""
    srcdir = os.path.join(_HERE, "csrc")
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    cmd = ["make", "-j4", "-C", srcdir]
    if extra:
        cmd.append(f"EXTRA={extra}")
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library.  Raises (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc -gencode arch=compute_100a,code=sm_100a); qmann_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.qmann_last_error.restype = C.c_char_p
    L.qmann_version.restype = C.c_char_p
    L.qmann_launch_count.restype = C.c_uint64
    L.qmann_model_create.restype = C.c_int
    L.qmann_model_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(QConfig), C.POINTER(QWeights)]
    L.qmann_model_destroy.restype = None
    L.qmann_model_destroy.argtypes = [C.c_void_p]
    L.qmann_model_load.restype = C.c_int
    L.qmann_model_load.argtypes = [C.POINTER(C.c_void_p), C.POINTER(QConfig), C.c_char_p]
    L.qmann_weights_dump.restype = C.c_int
    L.qmann_weights_dump.argtypes = [C.POINTER(QConfig), C.POINTER(QWeights), C.c_char_p]
    L.qmann_weights_last_error.restype = C.c_char_p
    L.qmann_batch_create.restype = C.c_int
    L.qmann_batch_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_U32), _U32]
    L.qmann_batch_destroy.restype = None
    L.qmann_batch_destroy.argtypes = [C.c_void_p]
    L.qmann_forward_batch.restype = C.c_int
    L.qmann_forward_batch.argtypes = [C.c_void_p, C.c_void_p, _FP, _FP, _FP, _FP, _FP, _FP, C.POINTER(QDebug), C.c_void_p]
    L.qmann_infer_host.restype = C.c_int
    L.qmann_infer_host.argtypes = [C.c_void_p, _FP, _FP, _FP, C.POINTER(_U32), _U32, C.POINTER(_U32), C.POINTER(_U32),
                                   C.POINTER(C.c_float)]
    L.qmann_forward_ids.restype = C.c_int
    L.qmann_forward_ids.argtypes = [C.c_void_p, C.c_void_p, _FP, _FP, _FP, _FP, _FP, _FP, C.POINTER(QDebug), C.c_void_p]
    L.qmann_infer_ids_host.restype = C.c_int
    L.qmann_infer_ids_host.argtypes = [C.c_void_p, _FP, _FP, _FP, C.POINTER(_U32), _U32, C.POINTER(_U32), C.POINTER(_U32),
                                       C.POINTER(C.c_float)]
    L.qmann_check_errors.restype = C.c_int
    L.qmann_check_errors.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_U32)]
    L.qmann_path_counts.restype = C.c_int
    L.qmann_path_counts.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.qmann_profile_enable.restype = C.c_int
    L.qmann_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.qmann_profile_read.restype = C.c_int
    L.qmann_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(_U32)]
    L.qmann_shard_plan.restype = C.c_int
    L.qmann_shard_plan.argtypes = [C.POINTER(_U32), _U32, _U32, _U32, C.POINTER(_U32), C.POINTER(_U32)]
    # part 3: slot-sharded large memory
    L.qmann_bigmem_last_error.restype = C.c_char_p
    L.qmann_bigmem_create.restype = C.c_int
    L.qmann_bigmem_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(QConfig), C.POINTER(QWeights), C.POINTER(_FP), C.POINTER(_FP),
                                      C.c_uint64, C.c_uint64, C.c_uint64, _U32]
    L.qmann_bigmem_destroy.restype = None
    L.qmann_bigmem_destroy.argtypes = [C.c_void_p]
    L.qmann_bigmem_num_bins.restype = _U32
    L.qmann_bigmem_num_bins.argtypes = [C.c_void_p]
    L.qmann_bigmem_forward_sharded.restype = C.c_int
    L.qmann_bigmem_forward_sharded.argtypes = [C.c_void_p, C.c_void_p, _FP, _U32, _FP, C.c_void_p]
    L.qmann_bigmem_begin.restype = C.c_int
    L.qmann_bigmem_begin.argtypes = [C.c_void_p, _FP, _U32, C.c_void_p]
    L.qmann_bigmem_hop_scores.restype = C.c_int
    L.qmann_bigmem_hop_scores.argtypes = [C.c_void_p, _U32, _FP, C.c_void_p]
    L.qmann_bigmem_hop_read.restype = C.c_int
    L.qmann_bigmem_hop_read.argtypes = [C.c_void_p, _U32, _FP, _FP, _FP, C.c_void_p]
    L.qmann_bigmem_hop_update.restype = C.c_int
    L.qmann_bigmem_hop_update.argtypes = [C.c_void_p, _U32, _FP, _FP, _FP, C.c_void_p]
    L.qmann_bigmem_state.restype = C.c_int
    L.qmann_bigmem_state.argtypes = [C.c_void_p, _FP, C.POINTER(C.c_int32), C.c_void_p]
    L.qmann_bigmem_profile_enable.restype = C.c_int
    L.qmann_bigmem_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.qmann_bigmem_profile_read.restype = C.c_int
    L.qmann_bigmem_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(_U32), C.c_int]
    L.qmann_bigmem_finish.restype = C.c_int
    L.qmann_bigmem_finish.argtypes = [C.c_void_p, _FP, _FP, _FP, C.c_void_p]
    _lib = L
    return L


class QmannError(RuntimeError):
    pass


def _check(rc: int):
    if rc != 0:
        raise QmannError(f"qmann error {rc}: {lib().qmann_last_error().decode()}")


def make_config(cfg) -> QConfig:
    f = cfg.formats()
    q = QConfig()
    q.V, q.d, q.S_max, q.H, q.mode, q.lin_map, q.const_scale = cfg.V, cfg.d, cfg.S_max, cfg.H, cfg.mode, int(cfg.lin_map), cfg.const_scale
    for h in range(cfg.H):
        q.iwl[h], q.frac[h] = f["iwl"][h], f["frac"][h]
        q.iwl_w[h], q.frac_w[h] = f["iwl_w"][h], f["frac_w"][h]
        q.iwl_att[h], q.frac_att[h] = f["iwl_att"][h], f["frac_att"][h]
    q.iwl_bin, q.frac_bin = f["iwl_bin"], f["frac_bin"]
    sc = getattr(cfg, "sc_att", None)
    q.en_sc_att = 1 if sc is not None else 0
    for h in range(cfg.H):
        q.sc_att_w[h] = float(sc[h]) if sc is not None else 0.0
    q.en_non_lin = 1 if getattr(cfg, "non_lin", False) else 0
    return q


def shard_plan(n_sen: np.ndarray, world: int, rank: int):
    """Contiguous story range of `rank` (qmann_shard_plan); pure host arithmetic."""
    ns = np.ascontiguousarray(n_sen, dtype=np.uint32)
    first, count = _U32(0), _U32(0)
    _check(lib().qmann_shard_plan(ns.ctypes.data_as(C.POINTER(_U32)), len(ns), world, rank, C.byref(first), C.byref(count)))
    return int(first.value), int(count.value)


class Model:
    """Quantised model resident on the current CUDA device (mirror of the layer structs the reference
    driver builds at MemN2N/MemN2N.c:826-912, collapsed into one object)."""

    def __init__(self, cfg, weights, device: str = "cuda:0"):
        import torch
        self.torch = torch
        self.cfg = cfg
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)
        self.w = dict(B=t(weights.B), W=t(weights.W), A=[t(x) for x in weights.A], C=[t(x) for x in weights.C],
                      Hm=[t(x) for x in weights.Hm])
        qw = QWeights()
        qw.dev_B, qw.dev_W = self.w["B"].data_ptr(), self.w["W"].data_ptr()
        for h in range(cfg.H):
            qw.dev_A[h], qw.dev_C[h], qw.dev_Hm[h] = self.w["A"][h].data_ptr(), self.w["C"][h].data_ptr(), self.w["Hm"][h].data_ptr()
        self._h = C.c_void_p()
        qc = make_config(cfg)
        _check(lib().qmann_model_create(C.byref(self._h), C.byref(qc), C.byref(qw)))

    @classmethod
    def from_weight_dir(cls, cfg, directory: str, device: str = "cuda:0") -> "Model":
        """Model from the reference driver's raw weight files (qmann_model_load, MemN2N/MemN2N.c:2553-2618 layout)."""
        import torch
        self = cls.__new__(cls)
        self.torch, self.cfg, self.device, self.w = torch, cfg, torch.device(device), {}
        torch.cuda.set_device(self.device)
        self._h = C.c_void_p()
        qc = make_config(cfg)
        rc = lib().qmann_model_load(C.byref(self._h), C.byref(qc), directory.encode())
        if rc != 0:
            raise QmannError(f"qmann_model_load error {rc}: {lib().qmann_weights_last_error().decode()}")
        return self

    def dump_weights(self, directory: str):
        """The fp32 device weights this model was created from, written in the reference's dump layout (qmann_weights_dump)."""
        qw = QWeights()
        qw.dev_B, qw.dev_W = self.w["B"].data_ptr(), self.w["W"].data_ptr()
        for h in range(self.cfg.H):
            qw.dev_A[h], qw.dev_C[h], qw.dev_Hm[h] = self.w["A"][h].data_ptr(), self.w["C"][h].data_ptr(), self.w["Hm"][h].data_ptr()
        qc = make_config(self.cfg)
        os.makedirs(directory, exist_ok=True)
        rc = lib().qmann_weights_dump(C.byref(qc), C.byref(qw), directory.encode())
        if rc != 0:
            raise QmannError(f"qmann_weights_dump error {rc}: {lib().qmann_weights_last_error().decode()}")

    def close(self):
        if self._h:
            lib().qmann_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def profile(self, enable: bool):
        _check(lib().qmann_profile_enable(self._h, int(enable)))

    def profile_read(self):
        """(ms in k_compact, ms in k_forward, launch pairs) since the last read; synchronises."""
        a, b, n = C.c_float(0), C.c_float(0), _U32(0)
        _check(lib().qmann_profile_read(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return float(a.value), float(b.value), int(n.value)

    def check_errors(self, stream=None) -> int:
        """Device-side error flag of the asynchronous entries (qmann_check_errors): synchronises, returns and clears it."""
        sptr = stream.cuda_stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        f = _U32(0)
        _check(lib().qmann_check_errors(self._h, C.c_void_p(sptr), C.byref(f)))
        return int(f.value)

    def path_counts(self, stream=None):
        """Stories that entered the (packed, unpacked, general) tier since the last call (qmann_path_counts)."""
        sptr = stream.cuda_stream if stream is not None else self.torch.cuda.current_stream().cuda_stream
        t = (C.c_uint64 * 3)()
        _check(lib().qmann_path_counts(self._h, C.c_void_p(sptr), t))
        return int(t[0]), int(t[1]), int(t[2])

    # ---- device-resident batch ---------------------------------------------------------------
    def upload(self, st) -> "DeviceBatch":
        return DeviceBatch(self, st)

    def forward(self, db: "DeviceBatch", with_answers: bool = True, want_h: bool = False, debug: bool = False,
                stream=None, production_dump: bool = False) -> Dict[str, object]:
        """One pass of the hot path over the whole device-resident batch (asynchronous).
        debug: every story through the instrumented general kernel, all intermediates returned.
        production_dump: the production kernels (k_story tiers + general kernel for what they decline) with their dump
        instantiations: u0, s, pcode (Q_f(p) codes = selected slots), o, g, u, z (exact rows), cand, path."""
        torch = self.torch
        N, H, d, V, ss = db.N, self.cfg.H, self.cfg.d, self.cfg.V, db.sum_sen
        out: Dict[str, object] = {}
        out["pred"] = db.pred
        db.match.zero_()
        dbg = None
        if debug or production_dump:
            dbg = QDebug()
            dbg.production = 1 if production_dump else 0
            shapes = dict(u0=(N, d), M=(H, ss, d), C=(H, ss, d), s=(H, ss), p=(H, ss), o=(H, N, d), g=(H, N, d), u=(H, N, d),
                          z=(N, V), h=(N, V))
            for k, shp in shapes.items():
                out[k] = torch.zeros(shp, dtype=torch.float32, device=self.device)
                setattr(dbg, f"dev_{k}", out[k].data_ptr())
            for k, shp in dict(pcode=(H, ss), path=(N,), cand=(N, V)).items():
                out[k] = torch.zeros(shp, dtype=torch.uint8, device=self.device)
                setattr(dbg, f"dev_{k}", out[k].data_ptr())
        sptr = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        if getattr(db, "ids", None) is not None:
            # word-id input (qmann_forward_ids): same records, same kernels after the compaction step
            _check(lib().qmann_forward_ids(self._h, db._h, db.ids.data_ptr(), db.row_off.data_ptr(),
                                           db.ans.data_ptr() if with_answers else None, db.pred.data_ptr(),
                                           db.h_true.data_ptr() if want_h else None,
                                           db.match.data_ptr() if with_answers else None,
                                           C.byref(dbg) if dbg is not None else None, C.c_void_p(sptr)))
        else:
            _check(lib().qmann_forward_batch(self._h, db._h, db.m.data_ptr(), db.q.data_ptr(),
                                             db.a.data_ptr() if with_answers else None, db.pred.data_ptr(),
                                             db.h_true.data_ptr() if want_h else None,
                                             db.match.data_ptr() if with_answers else None,
                                             C.byref(dbg) if dbg is not None else None, C.c_void_p(sptr)))
        out["match"] = db.match
        out["h_true"] = db.h_true
        return out

    # ---- host arenas in, predictions out (the call a user of the reference would make) -------
    def infer_host(self, m: np.ndarray, q: np.ndarray, a: Optional[np.ndarray], n_sen: np.ndarray, want_cost: bool = False):
        N = len(n_sen)
        pred = np.zeros(N, dtype=np.uint32)
        match, cost = _U32(0), C.c_float(0.0)
        ns = np.ascontiguousarray(n_sen, dtype=np.uint32)
        ptr = lambda x: None if x is None else C.c_void_p(x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr())
        _check(lib().qmann_infer_host(self._h, ptr(m), ptr(q), ptr(a), ns.ctypes.data_as(C.POINTER(_U32)), N,
                                      pred.ctypes.data_as(C.POINTER(_U32)), C.byref(match),
                                      C.byref(cost) if want_cost else None))
        return pred, int(match.value), float(cost.value)


    def upload_ids(self, ist) -> "DeviceIdBatch":
        return DeviceIdBatch(self, ist)

    def infer_ids_host(self, ids: np.ndarray, row_off: np.ndarray, ans: Optional[np.ndarray], n_sen: np.ndarray, want_cost: bool = False):
        """Word-id lists in (host), predictions out: qmann_infer_ids_host."""
        N = len(n_sen)
        pred = np.zeros(N, dtype=np.uint32)
        match, cost = _U32(0), C.c_float(0.0)
        ns = np.ascontiguousarray(n_sen, dtype=np.uint32)
        assert ids.dtype == np.uint16 and row_off.dtype == np.uint32 and (ans is None or ans.dtype == np.uint32)
        ptr = lambda x: None if x is None else C.c_void_p(x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr())
        _check(lib().qmann_infer_ids_host(self._h, ptr(ids), ptr(row_off), ptr(ans), ns.ctypes.data_as(C.POINTER(_U32)), N,
                                          pred.ctypes.data_as(C.POINTER(_U32)), C.byref(match),
                                          C.byref(cost) if want_cost else None))
        return pred, int(match.value), float(cost.value)


class DeviceIdBatch:
    """Stories as word-id lists resident in HBM (synth.IdStories): the input of qmann_forward_ids."""

    def __init__(self, model: Model, ist):
        torch = model.torch
        dev = model.device
        self.N, self.sum_sen = ist.N, int(ist.n_sen.sum())
        self.ids = torch.from_numpy(ist.ids.view(np.int16)).to(dev)
        self.row_off = torch.from_numpy(ist.row_off.view(np.int32)).to(dev)
        self.ans = torch.from_numpy(ist.ans.astype(np.uint32).view(np.int32)).to(dev)
        self.pred = torch.zeros(max(1, ist.N), dtype=torch.int32, device=dev)
        self.h_true = torch.zeros(max(1, ist.N), dtype=torch.float32, device=dev)
        self.match = torch.zeros(1, dtype=torch.int32, device=dev)
        ns = np.ascontiguousarray(ist.n_sen, dtype=np.uint32)
        self._h = C.c_void_p()
        _check(lib().qmann_batch_create(C.byref(self._h), ns.ctypes.data_as(C.POINTER(_U32)), ist.N))

    def close(self):
        if self._h:
            lib().qmann_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceBatch:
    """Packed stories resident in HBM: the arenas cuda_data_in() fills (MemN2N.c:2336-2350)."""

    def __init__(self, model: Model, st):
        torch = model.torch
        dev = model.device
        self.N, self.sum_sen = st.N, st.sum_sen
        self.m = torch.from_numpy(np.ascontiguousarray(st.m, dtype=np.float32)).to(dev)
        self.q = torch.from_numpy(np.ascontiguousarray(st.q, dtype=np.float32)).to(dev)
        self.a = torch.from_numpy(np.ascontiguousarray(st.a, dtype=np.float32)).to(dev)
        self.pred = torch.zeros(max(1, st.N), dtype=torch.int32, device=dev)
        self.h_true = torch.zeros(max(1, st.N), dtype=torch.float32, device=dev)
        self.match = torch.zeros(1, dtype=torch.int32, device=dev)
        ns = np.ascontiguousarray(st.n_sen, dtype=np.uint32)
        self._h = C.c_void_p()
        _check(lib().qmann_batch_create(C.byref(self._h), ns.ctypes.data_as(C.POINTER(_U32)), st.N))

    def close(self):
        if self._h:
            lib().qmann_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _bcheck(rc: int):
    if rc != 0:
        raise QmannError(f"qmann bigmem error {rc}: {lib().qmann_bigmem_last_error().decode()}")


def slot_shard(S_total: int, world: int, rank: int):
    """Contiguous slot range [slot0, slot0 + S_local) of `rank`: slots are uniform work, so equal ranges."""
    lo = S_total * rank // world
    hi = S_total * (rank + 1) // world
    return lo, hi - lo


class NcclComm:
    """A NCCL communicator over the ranks of the default torch.distributed group, for qmann_bigmem_forward_sharded (a C host
    would call ncclGetUniqueId / ncclCommInitRank itself; torch does not expose the ncclComm_t of its process groups).
    Rank 0 makes the unique id, the group broadcasts it, every rank initialises its communicator on the current device."""

    class _Uid(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.nccl = C.CDLL("libnccl.so.2")               # the library the process (torch) already uses
        self.nccl.ncclGetUniqueId.argtypes = [C.POINTER(NcclComm._Uid)]
        self.nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclComm._Uid, C.c_int]
        self.nccl.ncclCommDestroy.argtypes = [C.c_void_p]
        uid = NcclComm._Uid()
        rank, world = dist.get_rank(), dist.get_world_size()
        dbg = (lambda *a: print(f"[NcclComm rank {rank}]", *a, flush=True)) if os.environ.get("QMANN_NCCL_DEBUG") else (lambda *a: None)
        if rank == 0 and self.nccl.ncclGetUniqueId(C.byref(uid)) != 0:
            raise QmannError("ncclGetUniqueId failed")
        dbg("unique id made")
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).cuda()
        dist.broadcast(t, src=0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        dbg("unique id broadcast; ncclCommInitRank")
        self.comm = C.c_void_p()
        if self.nccl.ncclCommInitRank(C.byref(self.comm), world, uid, rank) != 0:
            raise QmannError("ncclCommInitRank failed")

    def close(self, destroy: bool = False):
        """Forget the communicator.  ncclCommDestroy is only called on request: it blocks while CUDA graphs that captured collectives
        of this communicator are alive (destroy the BigMemory objects first); otherwise process exit reclaims it."""
        if self.comm and destroy:
            self.nccl.ncclCommDestroy(self.comm)
        self.comm = C.c_void_p()


class BigMemory:
    """One very large pre-embedded memory, this rank's slot shard resident in HBM (include/qmann_abi.h part 3).

    M8, C8: int8 [H][S_local][d] codes in the hops' weight formats (torch tensors on the device or numpy arrays).
    `group`: a torch.distributed process group (NCCL) spanning the shards, or None for a single shard; the two
    per-hop exchanges are all_reduce(SUM) of the integer score histograms and of the integer partial reads."""

    def __init__(self, cfg, weights, M8, C8, S_total: int, slot0: int, Q_max: int, device: str = "cuda:0", group=None,
                 world: int = 1):
        import torch
        self.torch = torch
        self.cfg = cfg
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        to_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(self.device).contiguous()
        self.M = [to_dev(M8[h]) for h in range(cfg.H)]
        self.C = [to_dev(C8[h]) for h in range(cfg.H)]
        assert all(t.dtype == torch.int8 for t in self.M + self.C)
        self.S_local = int(self.M[0].shape[0])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)
        self.w = dict(W=t(weights.W), Hm=[t(x) for x in weights.Hm])
        qw = QWeights()
        qw.dev_W = self.w["W"].data_ptr()
        for h in range(cfg.H):
            qw.dev_Hm[h] = self.w["Hm"][h].data_ptr()
        Mp, Cp = (_FP * MAX_HOP)(), (_FP * MAX_HOP)()
        for h in range(cfg.H):
            Mp[h], Cp[h] = self.M[h].data_ptr(), self.C[h].data_ptr()
        self._h = C.c_void_p()
        qc = make_config(cfg)
        _bcheck(lib().qmann_bigmem_create(C.byref(self._h), C.byref(qc), C.byref(qw), Mp, Cp, S_total, slot0, self.S_local, Q_max))
        self.NB = int(lib().qmann_bigmem_num_bins(self._h))
        self.Q_max = Q_max
        self.group, self.world = group, world
        self.hist = torch.zeros((Q_max, self.NB), dtype=torch.int32, device=self.device)
        self.partial = torch.zeros((Q_max, cfg.d), dtype=torch.int32, device=self.device)
        self.pred = torch.zeros(Q_max, dtype=torch.int32, device=self.device)
        self.u_out = torch.zeros((Q_max, cfg.d), dtype=torch.int8, device=self.device)

    def close(self):
        if self._h:
            lib().qmann_bigmem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def profile(self, enable: bool):
        _bcheck(lib().qmann_bigmem_profile_enable(self._h, int(enable)))

    def profile_read(self, reset: bool = False):
        """(ms spent in k_big_scores, launches) accumulated since the last reset; call after every forward."""
        ms, n = C.c_float(0), _U32(0)
        _bcheck(lib().qmann_bigmem_profile_read(self._h, C.byref(ms), C.byref(n), int(reset)))
        return float(ms.value), int(n.value)

    def _allreduce(self, t):
        if self.group is not None and self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def forward_sharded(self, u0, comm: Optional["NcclComm"] = None):
        """The whole forward in one C call (qmann_bigmem_forward_sharded): the exchanges run on `comm` (None: a single shard);
        on a non-default current stream the sequence is captured into a CUDA graph once and replayed.  Returns the predictions."""
        torch = self.torch
        Q = int(u0.shape[0])
        assert u0.dtype == torch.int8 and u0.is_cuda and u0.is_contiguous() and Q <= self.Q_max
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _bcheck(lib().qmann_bigmem_forward_sharded(self._h, comm.comm if comm is not None else None, u0.data_ptr(), Q, self.pred.data_ptr(), st))
        return self.pred[:Q]

    def forward(self, u0, debug: bool = False, answer: bool = True):
        """u0: int8 [Q][d] device tensor (codes in the hop-0 weight format).  Asynchronous on the current stream.
        Returns dict(pred, u) (+ per-hop o, g, u, hist, pbin when debug)."""
        torch = self.torch
        L = lib()
        Q, d, H = int(u0.shape[0]), self.cfg.d, self.cfg.H
        assert u0.dtype == torch.int8 and u0.is_cuda and u0.is_contiguous() and Q <= self.Q_max
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _bcheck(L.qmann_bigmem_begin(self._h, u0.data_ptr(), Q, st))
        hist, partial = self.hist[:Q], self.partial[:Q]
        out = {}
        if debug:
            out.update(o=torch.zeros((H, Q, d), dtype=torch.int8, device=self.device), g=torch.zeros((H, Q, d), dtype=torch.int8, device=self.device),
                       u=torch.zeros((H, Q, d), dtype=torch.int8, device=self.device), hist=torch.zeros((H, Q, self.NB), dtype=torch.int32, device=self.device),
                       pbin=torch.zeros((H, Q, self.NB), dtype=torch.float32, device=self.device))
        for h in range(H):
            _bcheck(L.qmann_bigmem_hop_scores(self._h, h, hist.data_ptr(), st))
            self._allreduce(hist)
            _bcheck(L.qmann_bigmem_hop_read(self._h, h, hist.data_ptr(), partial.data_ptr(), out["pbin"][h].data_ptr() if debug else None, st))
            self._allreduce(partial)
            _bcheck(L.qmann_bigmem_hop_update(self._h, h, partial.data_ptr(), out["o"][h].data_ptr() if debug else None,
                                              out["g"][h].data_ptr() if debug else None, st))
            if debug:
                out["hist"][h].copy_(hist)
                _bcheck(L.qmann_bigmem_state(self._h, out["u"][h].data_ptr(), None, st))
        fb = C.c_int32(0)
        _bcheck(L.qmann_bigmem_state(self._h, self.u_out.data_ptr(), C.byref(fb), st))
        out["u_final"], out["frac_bits"] = self.u_out[:Q], int(fb.value)
        if answer and self.cfg.V:
            z = torch.zeros((Q, self.cfg.V), dtype=torch.float32, device=self.device) if debug else None
            _bcheck(L.qmann_bigmem_finish(self._h, self.pred.data_ptr(), z.data_ptr() if debug else None, None, st))
            out["pred"] = self.pred[:Q]
            if debug:
                out["z"] = z
        return out
