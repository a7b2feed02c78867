"""Exact brute-force L2 retrieval (TEST INFRASTRUCTURE; PARITY UNPINNED).

The reference has only `// TODO: index search` (rvc/src/rvc.rs:159); `index_path`/`index_rate`
are stored and never read (obs-rvc/src/lib.rs:78,81,264).  Restated from upstream RVC
(SURVEY.md Appendix C "Retrieval"):
    score, ix = index.search(feats, k)          # squared L2, ascending
    w = (1/score)^2 ; w /= sum(w)
    feats = index_rate * sum_i w_i * big_npy[ix_i] + (1-index_rate) * feats
applied to the 20 ms HuBERT frames before the 2x repeat.  `search` works in float64; ties
break toward the lowest row index.

`search_f32_ordered` is the literal-equality oracle of the engine's retrieval: it evaluates the fp32 squared
distance in the DEFINED summation order of the CUDA kernels (kernels_knn.cu knn_scan_kernel,
kernels_knn_umma.cu exact_d2) - lane l of 32 accumulates the float4 chunks l, l+32, ... of the row in ascending
order with one fused multiply-add per element (fmaf(d, d, a), d = x - y rounded to f32), then the 32 partial sums
are added by the xor-butterfly 16, 8, 4, 2, 1 - and selects the k smallest by (distance, row).  The fused
multiply-add is emulated exactly (f64 product is exact; the f64 sum is forced to round-to-odd before the final
rounding to f32, which removes the double rounding).  GPU indices AND distances must equal it bit for bit.
"""
import numpy as np


def _fmaf_sq(d32: np.ndarray, a32: np.ndarray) -> np.ndarray:
    """float32 fmaf(d, d, a) for d, a >= 0-sum accumulators (a >= 0), exactly rounded."""
    d = d32.astype(np.float64)
    p = d * d                                   # exact: 24-bit x 24-bit significands
    a = a32.astype(np.float64)
    s = p + a
    bb = s - p
    err = (p - (s - bb)) + (a - bb)             # TwoSum: s + err == p + a exactly
    bits = s.view(np.int64).copy()
    inexact = err != 0.0
    even = (bits & 1) == 0
    up = err > 0.0                              # all values are >= 0: +1 in bit space moves away from zero
    adj = np.where(inexact & even, np.where(up, 1, -1), 0).astype(np.int64)
    return (bits + adj).view(np.float64).astype(np.float32)   # round-to-odd f64 -> f32 is correctly rounded


def l2_f32_ordered(x: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """x (C,) f32, rows (M, C) f32 -> (M,) f32 squared distances in the kernels' summation order."""
    x = np.asarray(x, np.float32); rows = np.asarray(rows, np.float32)
    M, C = rows.shape
    assert C % 4 == 0
    C4 = C // 4
    nch = (C4 + 31) // 32
    pad = nch * 32 * 4
    xp = np.zeros(pad, np.float32); xp[:C] = x
    rp = np.zeros((M, pad), np.float32); rp[:, :C] = rows
    d = (xp[None, :] - rp).astype(np.float32).reshape(M, nch, 32, 4)   # [row, chunk round i, lane, element]
    acc = np.zeros((M, 32), np.float32)
    for i in range(nch):
        live = (np.arange(32) + 32 * i) < C4                         # lanes whose chunk exists in this round
        for e in range(4):
            nxt = _fmaf_sq(d[:, i, :, e], acc)
            acc = np.where(live[None, :], nxt, acc)
    for o in (16, 8, 4, 2, 1):
        acc = (acc + acc[:, np.arange(32) ^ o]).astype(np.float32)
    return acc[:, 0]


def search_f32_ordered(index: np.ndarray, queries: np.ndarray, k: int, shortlist: int = 64):
    """Literal-equality oracle: (d2[Q,k] float32, idx[Q,k] int64).  A float64 pass shortlists `shortlist` rows per
    query (every row that can reach the fp32 top-k: the f32 evaluation moves a distance by < 1e-5 relative), the
    ordered fp32 evaluation ranks them."""
    wd, wi = search(index, queries, min(shortlist, index.shape[0]))
    Q = queries.shape[0]
    out_d = np.empty((Q, k), np.float32); out_i = np.empty((Q, k), np.int64)
    for q in range(Q):
        cand = np.sort(wi[q])
        d32 = l2_f32_ordered(queries[q], index[cand])
        order = np.lexsort((cand, d32))[:k]
        out_d[q] = d32[order]; out_i[q] = cand[order]
        # the shortlist must be comfortably wider than the top-k in f64 terms
        assert wd[q, -1] > wd[q, k - 1] * (1 + 1e-4) or index.shape[0] <= shortlist
    return out_d, out_i


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
