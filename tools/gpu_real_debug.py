"""Debug: raw records of the real-valued tensor-core path vs the oracle's top-2 + ratio test, per query (STAGE_RAW, no finishing)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, matching, synth
ora = oracle.best()
n, m, seed = int(os.environ.get("DBG_N", 3)), int(os.environ.get("DBG_M", 700)), int(os.environ.get("DBG_SEED", 302))
descs, xys = synth.sift_images(n, m, np.float32, seed=seed, pool_factor=1.0)
real = [np.ascontiguousarray(d.astype(np.float32)) for d in synth.real_valued(descs, sigma=0.37)]
pairs = [(a, b) for a in range(n) for b in range(n) if a != b] + [(0, 0)]
mm = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_L2_B200)
mm.clear(); mm.upload({i: (real[i], xys[i]) for i in range(n)})
pid, off, raw = mm.match_uploaded(pairs, matching.STAGE_RAW)
print("real pairs", mm.ctx.last_real_tc_pairs(), "fallback rows", mm.ctx.last_fallback_rows(), "errors", mm.ctx.exactness_errors())
r2 = np.float32(0.8) * np.float32(0.8)
tot_bad = 0
for k, (a, b) in enumerate(pid.tolist()):
    rec = raw[off[k]:off[k + 1]]
    ok, ridx, rdist = ora.knn(real[a], real[b], 2, metric="l2_vectorized")
    keep = rdist[:, 0] < r2 * rdist[:, 1]
    want = {int(q): (int(ridx[q, 0]), float(rdist[q, 0]), float(rdist[q, 1])) for q in np.where(keep)[0]}
    got = {}
    dup = 0
    for r in rec:
        if int(r["j"]) in got: dup += 1
        got[int(r["j"])] = (int(r["i"]), float(r["dist"]), float(r["dist"]) / float(r["ratio"]) if r["ratio"] else 0.0)
    missing = sorted(set(want) - set(got)); extra = sorted(set(got) - set(want))
    wrong = [q for q in want if q in got and (got[q][0] != want[q][0] or got[q][1] != want[q][1])]
    if missing or extra or wrong or dup:
        tot_bad += 1
        print(f"pair ({a},{b}): {len(want)} wanted, {len(rec)} records, dup {dup}, missing {missing[:8]}, extra {extra[:8]}, wrong {wrong[:8]}")
        for q in (missing + extra + wrong)[:6]:
            print("   q", q, "want", want.get(q), "got", got.get(q), "true top2 rows", ridx[q].tolist(), "chunks", (ridx[q] // 16).tolist(), "dists", rdist[q].tolist())
print("pairs with differences:", tot_bad, "of", len(pid))
