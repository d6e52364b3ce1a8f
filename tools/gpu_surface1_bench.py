"""Surface 1 (ArrayMatcher) vs Surface 1b (IRegionsMatcher) on one 8192 x 8192 pair: per-call wall time, best / median of 20.
What a reference build does per pair with the two adaptors: RegionsMatcher<ArrayMatcher_b200>::Match = SearchNeighbours(NN = 2) + host
ratio filter + de-duplications; RegionsMatcher_b200::Match = upload query + b200m_match_pairs(1 pair, FULL)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alicevision_b200 import ArrayMatcherB200, Regions, RegionsMatcherB200, matching, synth

def stats(ts):
    ts = sorted(ts)
    return f"best {1e3 * ts[0]:.3f} ms, median {1e3 * ts[len(ts) // 2]:.3f} ms"

for m, dtype in ((8192, np.uint8), (8192, np.float32), (2048, np.uint8)):
    ds, xys = synth.sift_images(2, m, dtype, seed=5, pool_factor=1.0)
    am = ArrayMatcherB200(matching.L2_VECTORIZED)
    t0 = time.perf_counter(); assert am.Build(ds[0]); t_build = time.perf_counter() - t0
    ts = []
    for _ in range(22):
        t0 = time.perf_counter(); ok, idx, dist = am.SearchNeighbours(ds[1], NN=2); ts.append(time.perf_counter() - t0)
    assert ok and am.ctx.last_tc_pairs() == 1
    knn = stats(ts[2:])
    # the exact CUDA-core kernel the surface used in round 1 (force_exact routes b200m_knn to it)
    am.ctx.set_force_exact(True)
    te = []
    for _ in range(6):
        t0 = time.perf_counter(); ok2, idx2, dist2 = am.SearchNeighbours(ds[1], NN=2); te.append(time.perf_counter() - t0)
    am.ctx.set_force_exact(False)
    assert np.array_equal(dist, dist2)
    rm = RegionsMatcherB200(Regions(ds[0], xys[0]))
    q = Regions(ds[1], xys[1])
    tr = []
    for _ in range(22):
        t0 = time.perf_counter(); okm, matches = rm.Match(0.8, q); tr.append(time.perf_counter() - t0)
    print(f"{m} x {m} {np.dtype(dtype).name}: Build {1e3 * t_build:.2f} ms | ArrayMatcherB200.SearchNeighbours(NN=2) tensor-core: {knn} | exact CUDA-core kernel (round 1): {stats(te[1:])} | "
          f"RegionsMatcherB200.Match: {stats(tr[2:])} ({len(matches)} matches)", flush=True)
    rm.close()
