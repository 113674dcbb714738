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
