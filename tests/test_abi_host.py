"""CPU-only checks of the C-ABI library and the host-side logic (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "qmann_abi.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:cuda|qmann)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(qmann):
    L = qmann.lib.lib()
    names = _declared_symbols()
    assert len([n for n in names if n.startswith("cuda_")]) == 71
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.qmann_version()


def test_library_has_no_cpu_fallback(qmann):
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = qmann.synth.preset_config("C1")
    w = qmann.synth.make_weights(cfg, 0)
    with pytest.raises(Exception):
        qmann.lib.Model(cfg, w)


def test_product_never_touches_the_oracle():
    """Nothing under q-mann_b200/ may include, import, link or load the CPU oracle."""
    pkg = os.path.join(ROOT, "q-mann_b200")
    banned = [r'#\s*include\s*[<"][^>"]*oracle', r"\bimport\s+qmo\b", r"\bfrom\s+oracle\b", r"libqmann_oracle", r"qmo_forward",
              r"-lqmann_oracle"]
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for pat in banned:
                    assert not re.search(pat, txt), (f, pat)


def test_shard_plan_partitions_and_balances(qmann):
    rng = np.random.default_rng(0)
    n_sen = rng.integers(1, 51, size=10007).astype(np.uint32)
    for world in (1, 2, 3, 4, 8):
        cover, loads = [], []
        for r in range(world):
            first, count = qmann.lib.shard_plan(n_sen, world, r)
            cover.append((first, count))
            loads.append(int(n_sen[first:first + count].sum()) + count)
        assert cover[0][0] == 0 and sum(c for _, c in cover) == len(n_sen)
        for (f0, c0), (f1, _) in zip(cover, cover[1:]):
            assert f0 + c0 == f1
        assert max(loads) - min(loads) <= 2 * 51 + 2
    first, count = qmann.lib.shard_plan(np.zeros(0, dtype=np.uint32), 4, 2)
    assert (first, count) == (0, 0)


def test_formats_follow_reference_driver(qmann):
    cfg = qmann.synth.preset_config("C2")
    f = cfg.formats()
    assert (f["iwl_w"], f["frac_w"]) == ([6, 5, 4], [1, 2, 3])       # EN_MQ, MemN2N.c:748-754
    assert f["iwl"] == [5, 5, 5] and f["frac"] == [2, 2, 2] and (f["iwl_bin"], f["frac_bin"]) == (5, 2)
    qc = qmann.lib.make_config(cfg)
    assert qc.V == 256 and qc.d == 50 and qc.S_max == 64 and list(qc.iwl_w)[:3] == [6, 5, 4]


def test_unmodified_reference_driver_links_against_our_library():
    """oracle/Makefile `ref_driver_link`: the reference's own MemN2N.o + sample.o + layer.o + common.o link against
    libqmann_b200.so with --no-undefined, so every cuda_* symbol the driver and the layer code reference (66) is
    resolved by the linker.  Needs the reference sources (build container only)."""
    import subprocess
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir("/root/reference/MemN2N"):
        pytest.skip("reference sources not mounted")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref_driver_link"])
    assert os.path.exists(os.path.join(ref_dir, "memn2n_b200"))
    need = set()
    for obj in ("MemN2N.o", "layer.o"):
        out = subprocess.run(["nm", "-u", os.path.join(ref_dir, obj)], capture_output=True, text=True, check=True).stdout
        need |= {ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("cuda_")}
    have = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "q-mann_b200", "libqmann_b200.so")], capture_output=True, text=True, check=True).stdout
    have = {ln.split()[-1] for ln in have.splitlines() if ln.split()}
    assert len(need) >= 60 and need <= have, sorted(need - have)
