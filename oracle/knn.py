"""Exact brute-force L2 retrieval (TEST INFRASTRUCTURE; PARITY UNPINNED).

The reference has only `// TODO: index search` (rvc/src/rvc.rs:159); `index_path`/`index_rate`
are stored and never read (obs-rvc/src/lib.rs:78,81,264).  Restated from upstream RVC
(SURVEY.md Appendix C "Retrieval"):
    score, ix = index.search(feats, k)          # squared L2, ascending
    w = (1/score)^2 ; w /= sum(w)
    feats = index_rate * sum_i w_i * big_npy[ix_i] + (1-index_rate) * feats
applied to the 20 ms HuBERT frames before the 2x repeat.  Distances are float64 here; ties
break toward the lowest row index.
"""
import numpy as np


def search(index: np.ndarray, queries: np.ndarray, k: int, chunk: int = 65536):
    """Returns (d2[Q,k] float64 ascending, idx[Q,k] int64)."""
    q = np.asarray(queries, np.float64)
    Q = q.shape[0]
    best_d = np.full((Q, 0), np.inf)
    best_i = np.zeros((Q, 0), np.int64)
    qn = (q * q).sum(1)[:, None]
    for s in range(0, index.shape[0], chunk):
        blk = np.asarray(index[s:s + chunk], np.float64)
        d = qn - 2.0 * (q @ blk.T) + (blk * blk).sum(1)[None, :]
        if blk.shape[0] > 4 * k:
            # refine the candidates with the cancellation-free form
            cand = np.argpartition(d, 4 * k, axis=1)[:, :4 * k]
        else:
            cand = np.tile(np.arange(blk.shape[0]), (Q, 1))
        dd = ((q[:, None, :] - blk[cand]) ** 2).sum(-1)
        best_d = np.concatenate([best_d, dd], axis=1)
        best_i = np.concatenate([best_i, cand + s], axis=1)
        order = np.lexsort((best_i, best_d), axis=1)[:, :k]
        best_d = np.take_along_axis(best_d, order, 1)
        best_i = np.take_along_axis(best_i, order, 1)
    return best_d, best_i


def blend(index: np.ndarray, feats: np.ndarray, k: int, index_rate: float):
    """feats (Q,C) f32 -> blended (Q,C) f32, plus (d2, idx)."""
    d2, ix = search(index, feats, k)
    with np.errstate(divide="ignore"):
        wgt = np.square(1.0 / d2)
    wgt = wgt / wgt.sum(1, keepdims=True)
    mix = (index[ix].astype(np.float64) * wgt[:, :, None]).sum(1)
    out = index_rate * mix + (1.0 - index_rate) * feats.astype(np.float64)
    return out.astype(np.float32), d2, ix
