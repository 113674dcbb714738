"""Reader of the reference's parsed bAbI sets (q-mann_b200/babi.py) -> word-id lists, no GPU.

The expected arenas are built by a literal restatement of MemN2N/sample.c:544-572 (words add 1.0, the time column is
set to 1.0, question and answer words add 1.0) from an independent walk over the samples; scattering the id lists must
reproduce them.  Where /root/reference is mounted (the build container), the shipped qa1 sets pin the reader to the
facts of SURVEY.md A.8 (dictionary 20 incl. NULL, stories <= 10 sentences, <= 6 words per sentence: V = 30, dim_word = 7)."""
import os

import numpy as np
import pytest

REF_SETS = "/root/reference/MemN2N/dataset/en_1k_parsed"


def _write_set(path, samples):
    with open(path, "w") as fh:
        fh.write("\n+NS+\n%d\n" % len(samples))
        for i, (sens, q, a) in enumerate(samples):
            fh.write("\n+I+\n%d\n+S+\n%d\n" % (i, len(sens)))
            for s in sens:
                fh.write(" ".join(s) + " \n")
            fh.write("+Q+\n" + " ".join(q) + " \n+A+\n" + " ".join(a) + "\n")


def _literal_dense(babi, samples, dictionary, dims):
    """sample_init + sample_vectorization, literally, one story at a time."""
    V = dims.dim_input
    m_rows, q_rows, a_rows = [], [], []
    for s in samples:
        ns = len(s.sentences)
        for j, sen in enumerate(s.sentences):
            words = sen[:dims.dim_word - 1]
            row = np.zeros(V, np.float32)
            for w in words:
                row[dictionary.idx(w)] += 1.0
            row[dims.dim_dict + ns - j - 1] = 1.0
            m_rows.append(row)
        q = np.zeros(V, np.float32)
        for w in s.question[:dims.dim_word - 1]:
            q[dictionary.idx(w)] += 1.0
        a = np.zeros(V, np.float32)
        for w in s.answer[:dims.dim_word - 1]:
            a[dictionary.idx(w)] += 1.0
        q_rows.append(q); a_rows.append(a)
    return np.array(m_rows).reshape(-1, V), np.array(q_rows), np.array(a_rows)


def _scatter(ist, V):
    n_rows = len(ist.row_off) - 1
    dense = np.zeros((n_rows, V), np.float32)
    rows = np.repeat(np.arange(n_rows), np.diff(ist.row_off.astype(np.int64)))
    np.add.at(dense, (rows, ist.ids.astype(np.int64)), 1.0)
    off = np.concatenate([[0], np.cumsum(ist.n_sen.astype(np.int64))])
    first = off[:-1] + np.arange(len(ist.n_sen))
    mask = np.ones(n_rows, bool); mask[first] = False
    return dense[mask], dense[first]


def test_reader_on_a_synthetic_set(tmp_path, qmann):
    babi = qmann.babi
    rng = np.random.default_rng(5)
    vocab = ["Mary", "john", "Went", "to", "the", "kitchen", "garden", "where", "is", "apple", "took", "left", "hallway"]
    samples = []
    for _ in range(30):
        ns = int(rng.integers(0, 15))
        sens = [[vocab[int(k)] for k in rng.integers(0, len(vocab), size=int(rng.integers(1, 9)))] for _ in range(ns)]
        samples.append((sens, [vocab[int(k)] for k in rng.integers(0, len(vocab), size=3)], [vocab[int(rng.integers(0, len(vocab)))]]))
    samples[3] = ([["mary", "MARY", "Mary", "went"]], ["where", "is", "mary"], ["kitchen"])       # case-insensitive, counts of 3
    path = os.path.join(tmp_path, "toy_train_set")
    _write_set(path, samples)
    train = babi.read_parsed_set(path, max_len=64)
    assert len(train) == 30 and [len(s.sentences) for s in train] == [len(x[0]) for x in samples]
    d = babi.Dictionary(train)
    assert d.words[0] == babi.NULL_WORD and d.idx("MARY") == d.idx("mary") > 0 and d.idx("zebra") == -1
    dims = babi.dims_from_train(train, d)
    assert dims.dim_word == max(len(s) for x in samples for s in x[0]) + 1 and dims.max_line == max(len(x[0]) for x in samples)
    # a test set read with max_len = max_line keeps the LAST sentences of a longer story
    short = babi.read_parsed_set(path, max_len=4)
    for a, b in zip(short, train):
        assert a.sentences == b.sentences[-4:] if len(b.sentences) > 4 else a.sentences == b.sentences
    ist = babi.to_id_stories(train, d, dims)
    m_ref, q_ref, a_ref = _literal_dense(babi, train, d, dims)
    m_got, q_got = _scatter(ist, dims.dim_input)
    np.testing.assert_array_equal(m_got, m_ref)
    np.testing.assert_array_equal(q_got, q_ref)
    assert np.array_equal(ist.ans, a_ref.argmax(axis=1)) and m_ref.max() >= 2.0
    # truncation of long sentences to dim_word - 1 words: force a smaller dim_word
    small = babi.Dims(dim_dict=dims.dim_dict, max_line=dims.max_line, dim_word=4)
    ist2 = babi.to_id_stories(train, d, small)
    m_ref2, q_ref2, _ = _literal_dense(babi, train, d, small)
    m_got2, q_got2 = _scatter(ist2, small.dim_input)
    np.testing.assert_array_equal(m_got2, m_ref2)
    np.testing.assert_array_equal(q_got2, q_ref2)
    with pytest.raises(KeyError):
        babi.to_id_stories([babi.Sample([["zebra"]], ["where"], ["kitchen"])], d, dims)


@pytest.mark.skipif(not os.path.isdir(REF_SETS), reason="the reference's parsed sets are only mounted in the build container")
def test_reader_on_the_reference_qa1_sets(qmann):
    babi = qmann.babi
    train = babi.read_parsed_set(os.path.join(REF_SETS, "qa1_single-supporting-fact_train_set"), max_len=64)
    d = babi.Dictionary(train)
    dims = babi.dims_from_train(train, d)
    assert len(train) == 1000 and len(d) == 20 and dims.max_line == 10 and dims.dim_word == 7 and dims.dim_input == 30
    test = babi.read_parsed_set(os.path.join(REF_SETS, "qa1_single-supporting-fact_test_set"), max_len=dims.max_line)
    assert len(test) == 1000
    ist = babi.to_id_stories(test, d, dims)
    assert ist.N == 1000 and int(ist.n_sen.max()) <= 10 and int(ist.ids.max()) < dims.dim_input
    m_ref, q_ref, a_ref = _literal_dense(babi, test, d, dims)
    m_got, q_got = _scatter(ist, dims.dim_input)
    np.testing.assert_array_equal(m_got, m_ref)
    np.testing.assert_array_equal(q_got, q_ref)
    assert np.array_equal(ist.ans, a_ref.argmax(axis=1))
    # first test story of the set (SURVEY A.8 / the file itself): "Where is John" -> "hallway"
    assert d.words[int(ist.ans[0])].lower() == "hallway"
