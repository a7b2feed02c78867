"""kNN retrieval (configs[1] shape and the configs[4] stress shape on one GPU): N x C index, Q queries, top-k.
Host-observed time per rvc_knn_search (includes the H2D of the queries and the D2H of the result), exactness of
sampled queries against a float64 brute force."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "obs-rvc_b200")]
import numpy as np
import rvc_b200
from oracle import weights
p = weights.make_data_dir("/tmp/rvc_b200_data_seed7", seed=7, index_rows=40000)
for (n, c, q, k) in ([(1 << 20, 256, 128, 4)] if os.environ.get('ONLY1M') else [(40000, 768, 11, 8), (1 << 20, 256, 128, 4)]):
    eng = rvc_b200.RvcInfer(p["data"], index_k=k)
    rng = np.random.default_rng(2)
    rows = (rng.standard_normal((n, c), dtype=np.float32) * 0.34)
    eng.set_index(rows, 0.5)
    x = (rng.standard_normal((q, c), dtype=np.float32) * 0.34)
    r = eng.knn_search(x, k); idx = r[0] if np.asarray(r[0]).dtype.kind == "i" else r[1]
    for qi in ([] if os.environ.get('RVC_KU_DBG') else range(0, q, max(1, q // 4))):
        d = ((rows.astype(np.float64) - x[qi].astype(np.float64)) ** 2).sum(1)
        want = np.argsort(d, kind="stable")[:k]
        assert np.array_equal(np.asarray(idx[qi]), want), (qi, idx[qi], want)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); eng.knn_search(x, k); ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    print(f"KNN n={n} c={c} q={q} k={k}: {t*1e6:.1f} us/search (host-observed, incl. copies), index {n*c*4/1e6:.1f} MB -> {n*c*4/t/1e9:.0f} GB/s, "
          f"{2.0*q*n*c/t/1e12:.2f} TFLOP/s (2QNC)")
    for o in eng.profile_ops(10):
        print(f"    op {o['name']}: {o['us']:.1f} us device  ({n*c*4/o['us']/1e3:.0f} GB/s of the fp32 index bytes)")
    print(f"    guard fallbacks: {eng.knn_fallbacks()}")
    eng.close()
