"""Weight files in the layout of the reference's own (commented-out) dumps, MemN2N/MemN2N.c:2553-2618: round trip and
the exact byte order (for each hop, for each input column j, for each output row i: w_mat[i][j])."""
import os
import struct

import numpy as np
import pytest


def test_weight_files_round_trip_and_layout(tmp_path, qmann, synth):
    io = qmann.weights_io
    cfg = synth.preset_config("C1")
    w = synth.make_weights(cfg, 3, sigma=0.3, tied=False)
    io.save_weights(str(tmp_path), cfg, w)
    got = io.load_weights(str(tmp_path), cfg)
    for a, b in zip([w.B, w.W] + w.A + w.C + w.Hm, [got.B, got.W] + got.A + got.C + got.Hm):
        np.testing.assert_array_equal(a, b)
    # literal walk of the reference's read loop over the A file
    raw = open(os.path.join(tmp_path, io.FILES["A"]), "rb").read()
    pos = 0
    for h in range(cfg.H):
        for j in range(cfg.V):
            for i in range(cfg.d):
                (v,) = struct.unpack_from("<f", raw, pos)
                pos += 4
                assert v == w.A[h][i][j]
    assert pos == len(raw)
    # W: for j in dim_in (= d) for i in dim_out (= V)
    raw = open(os.path.join(tmp_path, io.FILES["W"]), "rb").read()
    assert struct.unpack_from("<f", raw, 4 * (3 * cfg.V + 5))[0] == w.W[5][3]
    # a truncated file is refused; a missing linear map is an error unless explicitly tolerated
    with open(os.path.join(tmp_path, io.FILES["B"]), "ab") as fh:
        fh.write(b"\0\0\0\0")
    with pytest.raises(ValueError):
        io.load_weights(str(tmp_path), cfg)
    io.save_weights(str(tmp_path), cfg, w)
    os.remove(os.path.join(tmp_path, io.FILES["Hm"]))
    with pytest.raises(FileNotFoundError):
        io.load_weights(str(tmp_path), cfg)
    z = io.load_weights(str(tmp_path), cfg, require_lin_map=False)
    assert all(not h.any() for h in z.Hm)


@pytest.mark.gpu
def test_c_loader_and_dumper_on_the_device(qmann, synth, tmp_path):
    """qmann_model_load (C): a model built from the reference-layout files predicts exactly like the model built from the
    in-memory tensors (every intermediate of the instrumented pass identical); qmann_weights_dump writes byte-identical files."""
    import torch
    for preset, lin_map in (("C1", True), ("C2", True), ("C4", False)):
        cfg = synth.preset_config(preset, lin_map=lin_map)
        w = synth.make_weights(cfg, 5, sigma=0.5)
        st = synth.make_stories(cfg, 300, 6, S=min(cfg.S_max, 50), ragged=True)
        d1, d2 = str(tmp_path / f"{preset}_py"), str(tmp_path / f"{preset}_c")
        qmann.weights_io.save_weights(d1, cfg, w)
        m_mem = qmann.lib.Model(cfg, w)
        m_file = qmann.lib.Model.from_weight_dir(cfg, d1)
        outs = []
        for m in (m_mem, m_file):
            db = m.upload(st)
            o = m.forward(db, with_answers=True, want_h=True, debug=True)
            torch.cuda.synchronize()
            outs.append({k: v.cpu().numpy().copy() for k, v in o.items() if hasattr(v, "cpu")})
            p = m.forward(db, with_answers=True)          # production tiers
            torch.cuda.synchronize()
            outs[-1]["pred_prod"] = p["pred"].cpu().numpy()[:st.N].copy()
        for k in outs[0]:
            np.testing.assert_array_equal(outs[0][k], outs[1][k], err_msg=f"{preset}: {k}")
        m_mem.dump_weights(d2)
        for name in sorted(os.listdir(d1)):
            assert open(os.path.join(d1, name), "rb").read() == open(os.path.join(d2, name), "rb").read(), name
    # a file of the wrong size is refused with a message, not read past its end
    bad = str(tmp_path / "bad")
    qmann.weights_io.save_weights(bad, cfg, w)
    with open(os.path.join(bad, "w_float.bin"), "ab") as fh:
        fh.write(b"\0\0\0\0")
    with pytest.raises(qmann.lib.QmannError):
        qmann.lib.Model.from_weight_dir(cfg, bad)
