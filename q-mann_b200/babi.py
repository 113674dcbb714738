"""Host-side reader of the reference's parsed bAbI sets -> word-id lists for qmann_forward_ids.

Mirrors, in the reference's own terms (no tensors, no GPU):
  * sample_constructor      MemN2N/sample.c:87-247   file grammar  "\\n+NS+\\n<N>\\n" then per sample
                                                    "\\n+I+\\n<i>\\n+S+\\n<n_sen>\\n<sentences>\\n+Q+\\n<question>\\n+A+\\n<answer>\\n";
                                                    a story longer than max_len keeps its LAST max_len sentences (:158-166)
  * dictionary_constructor  MemN2N/sample.c:852-921  index 0 is the NULL word, then first occurrence order over the
                                                    dictionary set (sentences, question, answer of each sample),
                                                    case-insensitive (strcasecmp)
  * sample_init             MemN2N/sample.c:337-411  a sentence keeps at most dim_word-1 words and gets one time id;
                                                    question and answer keep at most dim_word-1 words
  * sample_vectorization    MemN2N/sample.c:466-496  word -> id, time id of sentence j = dim_dict + n_sen - j - 1
and the dimension rules of MemN2N/MemN2N.c:544-582 (dim_dict = |dictionary|, dim_input = dim_dict + max_line,
dim_word = max_word + 1 with time encoding).  The dense arenas the reference builds from these ids (sample.c:544-572:
every occurrence adds 1.0, the time column is set to 1.0) are what scattering IdStories gives.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence

import numpy as np

NULL_WORD = "NULL"          # NULL_CHAR of the reference; index 0 (null_ind, MemN2N.c:542)


@dataclasses.dataclass
class Sample:
    sentences: List[List[str]]
    question: List[str]
    answer: List[str]


def read_parsed_set(path: str, max_len: int, num_samples: Optional[int] = None) -> List[Sample]:
    """sample_constructor: max_len = MAX_SEN_LEN for the train/dictionary set, max_line for the test set."""
    with open(path, "r") as fh:
        lines = fh.read().split("\n")
    pos = 0

    def take() -> str:
        nonlocal pos
        if pos >= len(lines):
            raise ValueError(f"{path}: unexpected end of file")
        ln = lines[pos]
        pos += 1
        return ln

    take(); tag = take()
    if tag != "+NS+":
        raise ValueError(f"{path}: expected +NS+, found {tag!r}")
    n_s = int(take())
    if num_samples is not None:
        n_s = min(n_s, num_samples)
    out: List[Sample] = []
    while len(out) < n_s and pos < len(lines):
        take()                                   # blank line
        if pos >= len(lines) or lines[pos] != "+I+":
            break
        take(); take()                           # +I+, index (the reader numbers samples itself, sample.c:143)
        if take() != "+S+":
            raise ValueError(f"{path}: expected +S+ in sample {len(out)}")
        n_ori = int(take())
        sens = [take() for _ in range(n_ori)]
        sens = sens[max(0, n_ori - max_len):]    # the first n_ori - max_len sentences are skipped
        if take() != "+Q+":
            raise ValueError(f"{path}: expected +Q+ in sample {len(out)}")
        q = take()
        if take() != "+A+":
            raise ValueError(f"{path}: expected +A+ in sample {len(out)}")
        a = take()
        out.append(Sample([s.split() for s in sens], q.split(), a.split()))
    return out


class Dictionary:
    """dictionary_constructor + word_idx."""

    def __init__(self, samples: Sequence[Sample]):
        self.words: List[str] = [NULL_WORD]
        self._idx: Dict[str, int] = {NULL_WORD.lower(): 0}
        for s in samples:
            for group in (*s.sentences, s.question, s.answer):
                for w in group:
                    k = w.lower()
                    if k not in self._idx:
                        self._idx[k] = len(self.words)
                        self.words.append(w)

    def __len__(self) -> int:
        return len(self.words)

    def idx(self, w: str) -> int:
        """word_idx: -1 (and the reference prints NO WORD IN DICT) when the word is unknown."""
        return self._idx.get(w.lower(), -1)


@dataclasses.dataclass
class Dims:
    dim_dict: int
    max_line: int
    dim_word: int

    @property
    def dim_input(self) -> int:          # V
        return self.dim_dict + self.max_line


def dims_from_train(train: Sequence[Sample], dictionary: Dictionary) -> Dims:
    """MemN2N.c:544-582 with time encoding and DIM_FORCED false."""
    max_line = max((len(s.sentences) for s in train), default=0)
    max_word = max((len(x) for s in train for x in s.sentences), default=0)
    return Dims(dim_dict=len(dictionary), max_line=max_line, dim_word=max_word + 1)


def to_id_stories(samples: Sequence[Sample], dictionary: Dictionary, dims: Dims):
    """sample_init + the id part of sample_vectorization -> IdStories (rows story-major: question, then sentences).
    Unknown words (word_idx = -1) raise: the reference would index out of bounds."""
    from .synth import IdStories
    ids: List[int] = []
    row_off: List[int] = [0]
    ans: List[int] = []
    n_sen: List[int] = []

    def wid(w: str) -> int:
        k = dictionary.idx(w)
        if k < 0:
            raise KeyError(f"NO WORD IN DICT : {w}")
        return k

    for s in samples:
        ns = len(s.sentences)
        if ns > dims.max_line:
            raise ValueError("a story has more sentences than max_line (read the set with max_len = max_line)")
        n_sen.append(ns)
        ids.extend(wid(w) for w in s.question[:dims.dim_word - 1])
        row_off.append(len(ids))
        for j, sen in enumerate(s.sentences):
            ids.extend(wid(w) for w in sen[:dims.dim_word - 1])
            ids.append(dims.dim_dict + ns - j - 1)                          # time id, sample.c:474
            row_off.append(len(ids))
        ans.append(wid(s.answer[0]) if s.answer else 0)
    assert max(ids, default=0) < 65536
    return IdStories(ids=np.asarray(ids, dtype=np.uint16), row_off=np.asarray(row_off, dtype=np.uint32),
                     ans=np.asarray(ans, dtype=np.uint32), n_sen=np.asarray(n_sen, dtype=np.uint32))
