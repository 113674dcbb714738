#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference CUDA implementation
(oracle/_ref, built by `make -C oracle ref` from /root/reference) on a GPU.

TEST INFRASTRUCTURE ONLY.  Run on a B200 box:
    python oracle/gen_golden.py --out gpurun_out/golden
then copy the .npz files into tests/golden/ and commit them.  Two kinds of fixture:

  case_<name>.npz  whole-forward cases: seeded model + stories (stored, so the fixture does not
                   depend on numpy's RNG stream) and EVERY intermediate tensor the reference
                   produced through its own layer API (oracle/ref_harness.c -> ref_harness_refcuda)
  kat_<name>.npz   kernel-level known-answer tables obtained by calling single reference
                   cuda_* entry points (oracle/_ref/libqmann_ref.so) on exhaustive 8-bit operands

Device memory for the kat_* tables is plain torch tensors (pointers passed through ctypes).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib.util
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _load_synth():
    spec = importlib.util.spec_from_file_location("qmann_synth", os.path.join(ROOT, "q-mann_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["qmann_synth"] = mod
    spec.loader.exec_module(mod)
    return mod


synth = _load_synth()

# name -> (config kwargs, weights kwargs, stories kwargs, post-processing tag)
CASES = {
    # BASELINE config 1 shape (d=20, S=50, V=70), fixed-point dot attention
    "c1_mode2": (dict(V=70, d=20, S_max=50, V_dict=20, mode=2), dict(seed=0x5EED0001, sigma=0.5),
                 dict(N=16, seed=0x5EED1001, ragged=True), None),
    # BASELINE config 2 shape (d=50, S<=50 of max_line 64, V=256)
    "c2_mode2": (dict(V=256, d=50, S_max=64, V_dict=192, mode=2), dict(seed=0x5EED0002, sigma=0.5),
                 dict(N=8, seed=0x5EED1002, S=50, ragged=True), None),
    # BASELINE config 3: same shape, Hamming / approximate attention
    "c3_mode3": (dict(V=256, d=50, S_max=64, V_dict=192, mode=3), dict(seed=0x5EED0003, sigma=0.5),
                 dict(N=8, seed=0x5EED1003, S=50, ragged=True), None),
    # BASELINE config 4 shape (d=64, S=50, V=114), saturating weights
    "c4_mode2_sat": (dict(V=114, d=64, S_max=64, V_dict=50, mode=2), dict(seed=0x5EED0004, sigma=1.0),
                     dict(N=8, seed=0x5EED1004, S=50), None),
    # odd shape, untied weights, no linear mapping, no EN_MQ, base format (3,4)
    "odd_mode2": (dict(V=37, d=13, S_max=17, V_dict=20, mode=2, lin_map=False, en_mq=False, iwl=3),
                  dict(seed=0x5EED0005, sigma=0.7, tied=False), dict(N=12, seed=0x5EED1005, ragged=True), None),
    # mode 3 with a narrow integer range so that hop-0 (EN_MQ) values saturate the 31-bit encode
    "sat_mode3": (dict(V=70, d=20, S_max=50, V_dict=20, mode=3, iwl=3), dict(seed=0x5EED0006, sigma=1.5),
                  dict(N=8, seed=0x5EED1006, ragged=True), None),
    # pure-float attention (ATTENTION_MODE 1)
    "c1_mode1": (dict(V=70, d=20, S_max=50, V_dict=20, mode=1), dict(seed=0x5EED0007, sigma=0.5),
                 dict(N=6, seed=0x5EED1007, ragged=True), None),
    # general BoW values: repeated words (counts up to 4) and fractional position-encoding-like
    # weights; single-sentence stories
    "frac_bow": (dict(V=70, d=20, S_max=50, V_dict=20, mode=2), dict(seed=0x5EED0008, sigma=0.8),
                 dict(N=10, seed=0x5EED1008, ragged=True, min_words=1, max_words=9), "frac"),
    # frac=6 base format: many slots keep a non-zero quantised attention weight
    "hifrac_mode2": (dict(V=70, d=20, S_max=50, V_dict=20, mode=2, iwl=1), dict(seed=0x5EED0009, sigma=0.3),
                     dict(N=8, seed=0x5EED1009, ragged=True), None),
}


def build_case(name):
    ckw, wkw, skw, tag = CASES[name]
    cfg = synth.ModelConfig(**ckw)
    w = synth.make_weights(cfg, **wkw)
    if tag == "frac":
        rng = np.random.default_rng(7)
        n_sen = rng.integers(1, cfg.S_max + 1, size=skw["N"])
        n_sen[:3] = 1                          # single-sentence stories
        st = synth.make_stories(cfg, n_sen=n_sen, **skw)
        scale = rng.choice(np.array([0.25, 0.5, 1.0, 1.0, 1.0, 1.5, 3.0, -1.0], dtype=np.float32), size=st.m.shape)
        st.m *= scale
        st.q *= rng.choice(np.array([0.5, 1.0, 1.0, 2.0], dtype=np.float32), size=st.q.shape)
    else:
        st = synth.make_stories(cfg, **skw)
    return cfg, w, st


def sparse(a):
    r, c = np.nonzero(a)
    return r.astype(np.uint32), c.astype(np.uint32), a[r, c].astype(np.float32)


def run_harness(cfg, w, st, time_reps=0):
    exe = os.path.join(HERE, "_ref", "ref_harness_refcuda")
    with tempfile.TemporaryDirectory() as td:
        case, dump = os.path.join(td, "case.bin"), os.path.join(td, "dump.bin")
        synth.write_case(case, cfg, w, st)
        res = subprocess.run([exe, case, dump] + ([str(time_reps)] if time_reps else []), check=True,
                             capture_output=True, text=True)
        out = synth.read_dump(dump)
    t = None
    for line in res.stdout.splitlines():
        if line.startswith("TIME_S"):
            t = float(line.split()[1])
    return out, t


def save_case(path, cfg, w, st, ref):
    import dataclasses
    d = {f"cfg_{k}": np.asarray(v) for k, v in dataclasses.asdict(cfg).items()}
    d["w_B"], d["w_W"] = w.B, w.W
    for h in range(cfg.H):
        d[f"w_A{h}"], d[f"w_C{h}"], d[f"w_Hm{h}"] = w.A[h], w.C[h], w.Hm[h]
    d["m_r"], d["m_c"], d["m_v"] = sparse(st.m)
    d["q_r"], d["q_c"], d["q_v"] = sparse(st.q)
    d["ans"], d["n_sen"] = st.ans, st.n_sen
    for k, v in ref.items():
        d[f"ref_{k}"] = v
    np.savez_compressed(path, **d)


# ---------------------------------------------------------------------------------------------
# kernel-level known-answer tables through libqmann_ref.so
# ---------------------------------------------------------------------------------------------
def kat_tables(out_dir):
    import torch
    L = C.CDLL(os.path.join(HERE, "_ref", "libqmann_ref.so"))
    u32, b, fp = C.c_uint, C.c_bool, C.c_void_p
    L.cuda_dot_mat_vec_fwd.argtypes = [fp, fp, fp, fp, u32, u32, b, b, u32, u32, u32, u32, u32, b]
    L.cuda_dot_mat_vec_fwd_appx.argtypes = [fp, fp, fp, fp, fp, u32, u32, b, u32, u32, u32, u32, b, b]
    L.cuda_sum_vec_fwd.argtypes = [fp, fp, fp, u32, b, u32, u32, u32, b]
    L.cuda_softmax_fwd.argtypes = [fp, fp, fp, fp, fp, u32, b, b]
    dev = torch.device("cuda:0")
    vals = np.arange(-127, 128, dtype=np.int32)

    # (1) FIXED_MUL + final quant via the scorer with d=1: s[r] = Q_m(Q_m(Q_m(M[r]) * Q_v(u)))
    #     for every 8-bit code pair; formats cover base, the EN_MQ skews and cross-format pairs
    fmt_pairs = [(5, 2, 5, 2), (6, 1, 5, 2), (4, 3, 5, 2), (5, 2, 6, 1), (5, 2, 4, 3), (6, 1, 6, 1), (4, 3, 4, 3),
                 (3, 4, 3, 4), (1, 6, 1, 6), (7, 0, 7, 0), (0, 7, 0, 7), (2, 5, 4, 3)]
    mul = np.zeros((len(fmt_pairs), 255, 255), dtype=np.float32)
    for k, (im, fm, iv, fv) in enumerate(fmt_pairs):
        M = torch.tensor(vals / 2.0 ** fm, dtype=torch.float32, device=dev).contiguous()
        out = torch.zeros(255, dtype=torch.float32, device=dev)
        for j, uv in enumerate(vals):
            u = torch.tensor([uv / 2.0 ** fv], dtype=torch.float32, device=dev)
            L.cuda_dot_mat_vec_fwd(M.data_ptr(), u.data_ptr(), out.data_ptr(), None, 255, 1, False, True,
                                   im, fm, iv, fv, 3, False)
            torch.cuda.synchronize()
            mul[k, :, j] = out.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "kat_fixed_mul.npz"), fmt_pairs=np.asarray(fmt_pairs, dtype=np.uint32),
                        vals=vals, out=mul)

    # (2) requantisation of values that live on a DIFFERENT grid than the scorer's matrix format
    #     (hop-0 / hop-2 memories are stored in (6,1) / (4,3) but scored in (5,2)); u fixed at 1.0
    rq_cases = [(6, 1, 5, 2), (4, 3, 5, 2), (5, 2, 5, 2), (4, 3, 3, 4), (2, 5, 3, 4)]
    rq = np.zeros((len(rq_cases), 255), dtype=np.float32)
    for k, (isrc, fsrc, im, fm) in enumerate(rq_cases):
        M = torch.tensor(vals / 2.0 ** fsrc, dtype=torch.float32, device=dev).contiguous()
        u = torch.tensor([1.0], dtype=torch.float32, device=dev)
        out = torch.zeros(255, dtype=torch.float32, device=dev)
        L.cuda_dot_mat_vec_fwd(M.data_ptr(), u.data_ptr(), out.data_ptr(), None, 255, 1, False, True, im, fm, im, fm, 3, False)
        torch.cuda.synchronize()
        rq[k] = out.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "kat_requant.npz"), cases=np.asarray(rq_cases, dtype=np.uint32), vals=vals, out=rq)

    # (3) approximate attention element: d=1, every (m,u) code pair, memory values on the grids
    #     they can have per hop (EN_MQ) and attention iwl in {5,3}
    ap_cases = [(5, 2, 2), (5, 1, 1), (5, 3, 2), (5, 1, 2), (3, 4, 4), (3, 3, 3), (3, 5, 4)]   # (iwl_att, frac of M grid, frac of u grid)
    ap = np.zeros((len(ap_cases), 255, 255), dtype=np.float32)
    for k, (ia, fM, fu) in enumerate(ap_cases):
        M = torch.tensor(vals / 2.0 ** fM, dtype=torch.float32, device=dev).contiguous()
        out = torch.zeros(255, dtype=torch.float32, device=dev)
        for j, uv in enumerate(vals):
            u = torch.tensor([uv / 2.0 ** fu], dtype=torch.float32, device=dev)
            L.cuda_dot_mat_vec_fwd_appx(M.data_ptr(), u.data_ptr(), out.data_ptr(), None, None, 255, 1, True,
                                        ia, 7 - ia, 3, 8, False, False)
            torch.cuda.synchronize()
            ap[k, :, j] = out.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "kat_appx_element.npz"), cases=np.asarray(ap_cases, dtype=np.uint32), vals=vals, out=ap)

    # (3b) approximate attention row sums with d=64 (accumulation + final quant/saturation)
    rng = np.random.default_rng(11)
    Mi = rng.integers(-127, 128, size=(64, 64)).astype(np.float32)
    ui = rng.integers(-127, 128, size=(64,)).astype(np.float32)
    rows = {}
    for ia, fM in [(5, 2), (5, 1), (3, 4), (2, 5)]:
        M = torch.tensor(Mi / 2.0 ** fM, device=dev)
        u = torch.tensor(ui / 2.0 ** fM, device=dev)
        out = torch.zeros(64, dtype=torch.float32, device=dev)
        L.cuda_dot_mat_vec_fwd_appx(M.data_ptr(), u.data_ptr(), out.data_ptr(), None, None, 64, 64, True, ia, 7 - ia, 3, 8, False, False)
        torch.cuda.synchronize()
        rows[f"out_{ia}_{fM}"] = out.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "kat_appx_rows.npz"), M=Mi, u=ui, **rows)

    # (4) FIXED_ADD via cuda_sum_vec_fwd over all code pairs; inputs on finer/coarser grids too
    add_cases = [(5, 2, 2, 2), (5, 2, 1, 2), (5, 2, 3, 2), (3, 4, 4, 4), (4, 3, 2, 3), (0, 7, 7, 7)]   # (iwl,frac, frac of a grid, frac of b grid)
    add = np.zeros((len(add_cases), 255, 255), dtype=np.float32)
    for k, (i_, f_, fa, fb) in enumerate(add_cases):
        a = torch.tensor(np.repeat(vals, 255) / 2.0 ** fa, dtype=torch.float32, device=dev)
        bb = torch.tensor(np.tile(vals, 255) / 2.0 ** fb, dtype=torch.float32, device=dev)
        out = torch.zeros(255 * 255, dtype=torch.float32, device=dev)
        for r in range(255):      # the reference launches <<<1,dim>>>, dim <= 1024
            L.cuda_sum_vec_fwd(a[r * 255:].data_ptr(), bb[r * 255:].data_ptr(), out[r * 255:].data_ptr(), 255, True, i_, f_, 3, False)
        torch.cuda.synchronize()
        add[k] = out.cpu().numpy().reshape(255, 255)
    np.savez_compressed(os.path.join(out_dir, "kat_fixed_add.npz"), cases=np.asarray(add_cases, dtype=np.uint32), vals=vals, out=add)

    # (5) softmax (max kernel + __expf + double total) on score-like vectors, incl. ties, large
    #     negative offsets (ex2 denormal path) and single elements
    vecs = []
    for dim in (1, 2, 3, 7, 50, 64, 256, 1000):
        for kind in range(4):
            if kind == 0:
                v = rng.integers(-127, 128, size=dim) / 4.0
            elif kind == 1:
                v = rng.integers(-8, 9, size=dim) / 4.0
            elif kind == 2:
                v = rng.standard_normal(dim) * 3.0
            else:
                v = np.where(rng.random(dim) < 0.3, 31.75, rng.integers(-127, 128, size=dim) / 4.0)
            vecs.append(v.astype(np.float32))
    vecs.append(np.array([0.0, -87.0, -88.0, -89.0, -100.0, -103.0, -104.0, -110.0], dtype=np.float32))
    sm_in = np.zeros((len(vecs), 1000), dtype=np.float32)
    sm_out = np.zeros((len(vecs), 1000), dtype=np.float32)
    dims = np.array([len(v) for v in vecs], dtype=np.uint32)
    mx = torch.zeros(1, dtype=torch.float32, device=dev)
    for k, v in enumerate(vecs):
        x = torch.tensor(v, device=dev)
        o = torch.zeros(len(v), dtype=torch.float32, device=dev)
        L.cuda_softmax_fwd(o.data_ptr(), x.data_ptr(), None, None, mx.data_ptr(), len(v), False, False)
        torch.cuda.synchronize()
        sm_in[k, :len(v)] = v
        sm_out[k, :len(v)] = o.cpu().numpy()
    np.savez_compressed(os.path.join(out_dir, "kat_softmax.npz"), dims=dims, inp=sm_in, out=sm_out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "golden"))
    ap.add_argument("--time", action="store_true", help="also time the reference GPU path on C1-C3 shapes")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    for name in CASES:
        cfg, w, st = build_case(name)
        ref, _ = run_harness(cfg, w, st)
        save_case(os.path.join(args.out, f"case_{name}.npz"), cfg, w, st, ref)
        print(f"case {name}: N={st.N} sum_sen={st.sum_sen} match={int(ref['match'])} cost={float(ref['cost']):.6g}", flush=True)
    kat_tables(args.out)
    print("kat tables done", flush=True)
    if args.time:
        lines = []
        for pname, n in (("C1", 1000), ("C2", 2000), ("C3", 2000)):
            cfg = synth.preset_config(pname)
            w = synth.make_weights(cfg, 0x5EED0000, sigma=0.5)
            st = synth.make_stories(cfg, n, 0x5EED1000, S=50)
            t0 = time.time()
            _, t = run_harness(cfg, w, st, time_reps=3)
            lines.append(f"{pname}: N={n} reference GPU path (31 launches/story, sm_100a rebuild): {t:.4f} s/pass = {n / t:.1f} stories/s")
            print(lines[-1], f"(harness wall {time.time() - t0:.1f}s)", flush=True)
        with open(os.path.join(args.out, "reference_gpu_timing.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
