"""CPU checks of the word-id format helpers (no GPU): the scatter of synth.ids_from_dense reproduces the dense arenas
(MemN2N/sample.c:547-568: every occurrence adds 1.0)."""
import numpy as np


def test_ids_roundtrip(synth):
    cfg = synth.preset_config("C1")
    st = synth.make_stories(cfg, 40, 3, ragged=True, max_words=15)
    st.q[5, :] = 0.0
    ist = synth.ids_from_dense(st)
    off = st.offsets()
    assert ist.row_off.shape == (st.N + st.sum_sen + 1,) and ist.row_off[-1] == ist.ids.shape[0]
    assert (st.m.max() >= 2) and ist.ids.dtype == np.uint16
    for i in range(st.N):
        r0 = off[i] + i
        qrow = np.zeros(cfg.V, np.float32)
        np.add.at(qrow, ist.ids[ist.row_off[r0]:ist.row_off[r0 + 1]].astype(np.int64), 1.0)
        np.testing.assert_array_equal(qrow, st.q[i])
        for j in range(int(st.n_sen[i])):
            row = np.zeros(cfg.V, np.float32)
            np.add.at(row, ist.ids[ist.row_off[r0 + 1 + j]:ist.row_off[r0 + 2 + j]].astype(np.int64), 1.0)
            np.testing.assert_array_equal(row, st.m[off[i] + j])
