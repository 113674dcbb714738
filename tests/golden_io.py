"""Load the committed golden fixtures (tests/golden/*.npz, produced by oracle/gen_golden.py from the
unmodified reference CUDA code on a B200) back into ModelConfig / Weights / Stories objects."""
from __future__ import annotations

import glob
import os
from typing import Dict, Tuple

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return sorted(os.path.basename(p)[len("case_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_case(name: str, synth) -> Tuple[object, object, object, Dict[str, np.ndarray]]:
    z = np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz"))
    ckw = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg_")}
    cfg = synth.ModelConfig(**ckw)
    H = cfg.H
    w = synth.Weights(B=z["w_B"], A=[z[f"w_A{h}"] for h in range(H)], C=[z[f"w_C{h}"] for h in range(H)],
                      Hm=[z[f"w_Hm{h}"] for h in range(H)], W=z["w_W"])
    n_sen = z["n_sen"].astype(np.uint32)
    N, ss = len(n_sen), int(n_sen.sum())
    m = np.zeros((ss, cfg.V), dtype=np.float32)
    m[z["m_r"], z["m_c"]] = z["m_v"]
    q = np.zeros((N, cfg.V), dtype=np.float32)
    q[z["q_r"], z["q_c"]] = z["q_v"]
    ans = z["ans"].astype(np.uint32)
    a = np.zeros((N, cfg.V), dtype=np.float32)
    a[np.arange(N), ans] = 1.0
    st = synth.Stories(m=m, q=q, a=a, n_sen=n_sen, ans=ans)
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    return cfg, w, st, ref


def load_kat(name: str) -> Dict[str, np.ndarray]:
    z = np.load(os.path.join(GOLDEN_DIR, f"kat_{name}.npz"))
    return {k: z[k] for k in z.files}
