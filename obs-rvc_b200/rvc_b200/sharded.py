"""Index-sharded retrieval across the GPUs of a node (SURVEY 8e "optional": the FAISS index of a voice model split by
rows; the reference hook is the TODO at rvc/src/rvc.rs:159, index settings obs-rvc/src/lib.rs:78,81,264).

Rank r holds rows [row_offset_r, row_offset_r + n_r) of the N x C index and its own batch of Q queries.  One search is
    1. all-gather the query batches                      (NCCL over NVLink: W x Q x C floats)
    2. every rank scans its shard for ALL W x Q queries  (the engine's exact top-k: rvc_knn_search)
    3. all-gather the candidate lists                    (W x [W x Q x k] (d2, global row) pairs)
    4. the owner of a query merges its W lists           (k smallest by (d2, row): ties -> lowest global row, the
                                                          single-GPU rule, so sharded == unsharded bit for bit)
The exchange steps are the only collectives of the whole path; everything else shards by stream with none.
`search_fn(queries, k) -> (d2, idx)` is the local searcher: `engine.knn_search` in the product, a CPU restatement in
the no-GPU tests (world-size-2 gloo)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_rows(n_rows: int, rank: int, world: int):
    """Contiguous, nearly equal row ranges: rank r owns [lo, hi)."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_candidates(d2: np.ndarray, idx: np.ndarray, k: int):
    """d2 / idx: (Q, W * k) candidates with GLOBAL row ids -> the k smallest per query, ordered by (d2, row)."""
    order = np.lexsort((idx, d2), axis=1)[:, :k]
    return np.take_along_axis(d2, order, axis=1), np.take_along_axis(idx, order, axis=1)


class ShardedIndex:
    def __init__(self, search_fn, row_offset: int, k: int, group=None, device=None):
        self.search_fn, self.row_offset, self.k, self.group = search_fn, int(row_offset), int(k), group
        self.device = device if device is not None else torch.device("cpu")

    def search(self, queries: np.ndarray):
        """queries: this rank's (Q, C) batch -> (d2, global idx) of shape (Q, k) for THIS rank's queries."""
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        q = torch.from_numpy(np.ascontiguousarray(queries, np.float32)).to(self.device)
        Q, k = q.shape[0], self.k
        allq = [torch.empty_like(q) for _ in range(world)]
        dist.all_gather(allq, q, group=self.group)                                   # step 1
        d2, idx = self.search_fn(torch.cat(allq).cpu().numpy(), k)                   # step 2: (W * Q, k), local rows
        cand = torch.empty((world * Q, k, 2), dtype=torch.float32, device=self.device)
        cand[:, :, 0] = torch.from_numpy(np.asarray(d2, np.float32)).to(self.device)
        # global row ids travel as their int32 bit patterns inside the same float32 tensor: one collective, exact
        gidx = (np.asarray(idx, np.int64) + self.row_offset).astype(np.int32)
        cand[:, :, 1] = torch.from_numpy(gidx.view(np.float32)).to(self.device)
        allc = [torch.empty_like(cand) for _ in range(world)]
        dist.all_gather(allc, cand, group=self.group)                                # step 3
        mine = torch.stack([c[rank * Q:(rank + 1) * Q] for c in allc], dim=1).cpu().numpy()   # (Q, W, k, 2)
        md2 = mine[..., 0].reshape(Q, world * k)
        midx = np.ascontiguousarray(mine[..., 1]).view(np.int32).reshape(Q, world * k).astype(np.int64)
        return merge_candidates(md2, midx, k)                                        # step 4
