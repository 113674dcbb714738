"""Word-id input (qmann_forward_ids / qmann_infer_ids_host) on a real B200.

The id lists are what MemN2N/sample.c holds before sample_vectorization scatters them into the dense arenas; the
bar is that the id path gives bit-identical tensors to the dense path (itself pinned to the reference's golden
tensors and the oracle by test_gpu_parity.py) on the same stories: every intermediate of the layer graph, the
predictions and the match count.  Covers repeated words (counts 2..4: split into unit entries or kept as exception
entries depending on the column's code range), ragged and empty stories, long rows (> 32 ids) and bad ids."""
import numpy as np
import pytest

import golden_io

pytestmark = pytest.mark.gpu
KEYS = ("u0", "M", "C", "s", "p", "o", "g", "u", "z", "h")


def _both(qmann, cfg, w, st, synth, debug=True):
    import torch
    model = qmann.lib.Model(cfg, w)
    dense = model.forward(model.upload(st), with_answers=True, want_h=True, debug=debug)
    torch.cuda.synchronize()
    dense = {k: (v.cpu().numpy().copy() if hasattr(v, "cpu") else v) for k, v in dense.items()}
    ist = synth.ids_from_dense(st)
    ids = model.forward(model.upload_ids(ist), with_answers=True, want_h=True, debug=debug)
    torch.cuda.synchronize()
    ids = {k: (v.cpu().numpy().copy() if hasattr(v, "cpu") else v) for k, v in ids.items()}
    return model, dense, ids, ist


def _same(dense, ids, N, keys):
    for k in keys:
        np.testing.assert_array_equal(ids[k], dense[k], err_msg=k)
    np.testing.assert_array_equal(ids["pred"][:N], dense["pred"][:N])
    np.testing.assert_array_equal(ids["h_true"][:N], dense["h_true"][:N])
    assert int(ids["match"][0]) == int(dense["match"][0])


@pytest.mark.parametrize("preset,sigma,seed,max_words", [("C1", 0.5, 1, 6), ("C2", 0.5, 2, 6), ("C3", 0.5, 3, 6), ("C4", 0.25, 4, 6),
                                                         ("C1", 2.0, 5, 40), ("C2", 3.0, 6, 12)])
def test_ids_forward_equals_dense_forward(preset, sigma, seed, max_words, qmann, synth):
    """sigma 2..3: large table codes, so repeated words cannot be split and travel as exception entries;
    max_words 40 with the 20-word dictionary of C1 gives counts up to ~6, rows longer than one warp and stories
    that overflow the fixed record into the heap."""
    cfg = synth.preset_config(preset)
    w = synth.make_weights(cfg, 300 + seed, sigma=sigma)
    st = synth.make_stories(cfg, 160, 400 + seed, S=min(cfg.S_max, 50), ragged=True, max_words=max_words)
    _, dense, ids, _ = _both(qmann, cfg, w, st, synth)
    _same(dense, ids, st.N, KEYS)


def test_ids_forward_golden_cases(qmann, synth):
    """The golden stories (reference outputs) through the id path wherever they hold integer counts."""
    done = 0
    for name in golden_io.case_names():
        if name == "c1_mode1":
            continue
        cfg, w, st, ref = golden_io.load_case(name, synth)
        if not (np.all(st.m == np.rint(st.m)) and np.all(st.q == np.rint(st.q)) and np.all(st.m >= 0) and np.all(st.q >= 0)):
            continue
        _, dense, ids, _ = _both(qmann, cfg, w, st, synth)
        _same(dense, ids, st.N, KEYS)
        np.testing.assert_array_equal(ids["pred"][:st.N].astype(np.uint32), ref["pred"])
        for k in ("u0", "M", "C", "s", "o", "g", "u"):
            if k == "g" and not cfg.lin_map:
                continue
            np.testing.assert_array_equal(ids[k], ref[k], err_msg=f"{name}: {k}")
        done += 1
    assert done >= 4


def test_ids_fast_path_large_batch_and_host_entry(qmann, synth):
    """Production path (no debug dumps, fast kernel): 5000 stories, device-resident and through the host entry."""
    import torch
    cfg = synth.preset_config("C2")
    w = synth.make_weights(cfg, 11, sigma=0.5)
    st = synth.make_stories(cfg, 5000, 12, S=50, ragged=True)
    model, dense, ids, ist = _both(qmann, cfg, w, st, synth, debug=False)
    _same(dense, ids, st.N, ())
    pred, match, cost = model.infer_ids_host(ist.ids, ist.row_off, ist.ans, ist.n_sen, want_cost=True)
    pred_d, match_d, cost_d = model.infer_host(st.m, st.q, st.a, st.n_sen, want_cost=True)
    np.testing.assert_array_equal(pred, pred_d)
    assert match == match_d and cost == cost_d
    np.testing.assert_array_equal(pred, dense["pred"][:st.N].astype(np.uint32))


def test_ids_edge_cases(qmann, synth):
    cfg = synth.preset_config("C1")
    w = synth.make_weights(cfg, 21, sigma=0.5)
    # stories with zero sentences and an empty question
    n_sen = np.array([0, 3, 0, 50, 1], dtype=np.uint32)
    st = synth.make_stories(cfg, 5, 22, n_sen=n_sen)
    st.q[2, :] = 0.0
    model, dense, ids, ist = _both(qmann, cfg, w, st, synth)
    _same(dense, ids, st.N, KEYS)
    # an id >= V is refused
    bad = synth.ids_from_dense(st)
    bad.ids[3] = cfg.V + 5
    with pytest.raises(qmann.lib.QmannError):
        model.infer_ids_host(bad.ids, bad.row_off, bad.ans, bad.n_sen)
    # and the model keeps working afterwards
    pred, _, _ = model.infer_ids_host(ist.ids, ist.row_off, ist.ans, ist.n_sen)
    np.testing.assert_array_equal(pred, dense["pred"][:st.N].astype(np.uint32))


def test_parsed_set_through_the_id_path(qmann, synth, tmp_path):
    """A set in the reference's parsed-file grammar -> babi reader -> qmann_forward_ids, against the dense arenas built
    literally as MemN2N/sample.c:544-572 does (test_babi_reader._literal_dense) through qmann_forward_batch."""
    import torch
    from test_babi_reader import _literal_dense, _write_set
    babi = qmann.babi
    rng = np.random.default_rng(9)
    vocab = [f"w{i}" for i in range(24)]
    samples = []
    for _ in range(200):
        ns = int(rng.integers(1, 21))
        sens = [[vocab[int(k)] for k in rng.integers(0, len(vocab), size=int(rng.integers(1, 7)))] for _ in range(ns)]
        samples.append((sens, [vocab[int(k)] for k in rng.integers(0, len(vocab), size=3)], [vocab[int(rng.integers(0, len(vocab)))]]))
    path = str(tmp_path / "toy_test_set")
    _write_set(path, samples)
    data = babi.read_parsed_set(path, max_len=64)
    d = babi.Dictionary(data)
    dims = babi.dims_from_train(data, d)
    ist = babi.to_id_stories(data, d, dims)
    cfg = synth.ModelConfig(V=dims.dim_input, d=20, S_max=dims.max_line, V_dict=dims.dim_dict, mode=2)
    w = synth.make_weights(cfg, 77, sigma=0.5)
    m, q, a = _literal_dense(babi, data, d, dims)
    st = synth.Stories(m=m, q=q, a=a, n_sen=ist.n_sen.copy(), ans=ist.ans.copy())
    model = qmann.lib.Model(cfg, w)
    dense = model.forward(model.upload(st), with_answers=True, want_h=True, debug=True)
    ids = model.forward(model.upload_ids(ist), with_answers=True, want_h=True, debug=True)
    torch.cuda.synchronize()
    for k in KEYS:
        np.testing.assert_array_equal(ids[k].cpu().numpy(), dense[k].cpu().numpy(), err_msg=k)
    np.testing.assert_array_equal(ids["pred"].cpu().numpy()[:st.N], dense["pred"].cpu().numpy()[:st.N])
    fast = model.forward(model.upload_ids(ist), with_answers=True, want_h=False, debug=False)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(fast["pred"].cpu().numpy()[:st.N], dense["pred"].cpu().numpy()[:st.N])


def test_c_host_demo(qmann):
    """examples/batched_demo.c: C host code (plain pointers, four cudart calls) through qmann_infer_host and
    qmann_infer_ids_host on the same stories; built by __graft_entry__.build()."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "batched_demo")
    if not os.path.exists(exe):
        pytest.skip("examples/batched_demo has not been built (python -c 'import __graft_entry__ as g; g.build()')")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("DEMO_OK"), r.stdout + r.stderr
