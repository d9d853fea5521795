"""Host-side feeding glue with the reference's names (/root/reference/utils/embedding.py:103-175): Vocab,
prepare_sentences, chunkize_serial, batch_generator, RepeatCorpusNTimes -- plus `paths_to_rows`, the vectorised
id -> row mapping the batched learners use.  `train_sg` (the fused pass) is re-exported from training_sdg_inner, as
the reference does at embedding.py:10-14."""
import itertools

import numpy as np

from .training_sdg_inner import TOKEN_NONE, train_sg  # noqa: F401


class Vocab(object):
    """A vocabulary item: `count` (node degree), `index` (row in the tables), `sample_probability`
    (embedding.py:164-175)."""

    def __init__(self, **kwargs):
        self.count = 0
        self.__dict__.update(kwargs)

    def __lt__(self, other):
        return self.count < other.count

    def __str__(self):
        vals = ["%s:%r" % (k, self.__dict__[k]) for k in sorted(self.__dict__) if not k.startswith("_")]
        return "<" + ", ".join(vals) + ">"


def chunkize_serial(iterable, chunksize, as_numpy=False):
    """Lists of `chunksize` items, the last one possibly shorter (embedding.py:103-124)."""
    it = iter(iterable)
    while True:
        chunk = list(itertools.islice(it, int(chunksize)))
        if as_numpy:
            chunk = [np.array(doc) for doc in chunk]
        if not chunk:
            return
        yield chunk


def prepare_sentences(model, paths):
    """id -> Vocab, dropping out-of-vocabulary nodes and down-sampled ones (embedding.py:126-136); consumes
    np.random.random_sample() only for nodes whose sample_probability < 1, like the reference."""
    vocab = model.vocab
    for path in paths:
        out = []
        for node in path:
            v = vocab.get(node) if hasattr(vocab, "get") else (vocab[node] if node in vocab else None)
            if v is None:
                continue
            if v.sample_probability >= 1.0 or v.sample_probability >= np.random.random_sample():
                out.append(v)
        yield out


def paths_to_rows(model, paths):
    """Vectorised prepare_sentences for a whole corpus: returns (flat uint32 row tokens, int64 offsets).
    Exactly prepare_sentences' filtering (OOV dropped; down-sampling draws one random_sample() per kept-candidate
    with probability < 1, in corpus order)."""
    flat, off = [], [0]
    ids, rows, probs = model.id_index()
    any_ds = bool((probs < 1.0).any())
    if isinstance(paths, RepeatCorpusNTimes) and isinstance(paths.corpus, np.ndarray) and paths.corpus.ndim == 2:
        paths = np.concatenate([paths.corpus] * paths.n) if paths.n != 1 else paths.corpus
    if isinstance(paths, np.ndarray) and paths.ndim == 2 and paths.dtype.kind in "iu":
        # a rectangular corpus (edge list, fixed-length walks): map every token at once
        a = paths.astype(np.int64, copy=False)
        if ids.size and int(ids[-1]) - int(ids[0]) + 1 == ids.size:  # dense ids (the usual 1..N): no search needed
            pos = a.ravel() - int(ids[0])
            keep = (pos >= 0) & (pos < ids.size)
            pos = np.where(keep, pos, 0)
        else:
            pos = np.searchsorted(ids, a.ravel())
            pos[pos >= ids.size] = 0
            keep = ids[pos] == a.ravel()
        if any_ds:  # one random_sample() per in-vocabulary token with probability < 1, in corpus order
            p = probs[pos]
            cand = np.flatnonzero(keep & (p < 1.0))
            keep[cand] = p[cand] >= np.random.random_sample(cand.size)
        lens = keep.reshape(a.shape).sum(1)
        offs = np.zeros(a.shape[0] + 1, np.int64)
        np.cumsum(lens, out=offs[1:])
        return np.ascontiguousarray(rows[pos[keep]].astype(np.uint32)), offs
    for path in paths:
        a = np.asarray(path, dtype=np.int64).ravel()
        pos = np.searchsorted(ids, a)
        pos[pos >= ids.size] = 0
        ok = ids[pos] == a
        r = rows[pos[ok]]
        if any_ds:
            p = probs[pos[ok]]
            keep = np.ones(r.size, bool)
            for q in np.flatnonzero(p < 1.0):
                keep[q] = p[q] >= np.random.random_sample()
            r = r[keep]
        flat.append(r.astype(np.uint32))
        off.append(off[-1] + r.size)
    flat = np.concatenate(flat) if flat else np.zeros(0, np.uint32)
    return np.ascontiguousarray(flat, np.uint32), np.asarray(off, np.int64)


def batch_generator(iterable, batch_size=1):
    args = [iter(iterable)] * batch_size
    return itertools.zip_longest(*args, fillvalue=None)


class RepeatCorpusNTimes(object):
    """Iterate `corpus` n times (embedding.py:147-160)."""

    def __init__(self, corpus, n):
        self.corpus = corpus
        self.n = n

    def __iter__(self):
        for _ in range(self.n):
            for document in self.corpus:
                yield document
