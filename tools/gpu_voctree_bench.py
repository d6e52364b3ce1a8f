"""Pair-list producer at BASELINE configs[2] shape: N images x 8192 SIFT descriptors through a K^L vocabulary tree,
all-against-all retrieval, top-50 -> pair list; GPU (include/b200voc.h) next to the CPU oracle on a bounded sample."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle
from alicevision_b200 import synth, voctree

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=300); ap.add_argument("--features", type=int, default=8192)
ap.add_argument("--k", type=int, default=10); ap.add_argument("--levels", type=int, default=4); ap.add_argument("--neighbours", type=int, default=50)
ap.add_argument("--cpu-images", type=int, default=3); ap.add_argument("--cpu-queries", type=int, default=24); ap.add_argument("--no-cpu", action="store_true")
a = ap.parse_args()
descs, _ = synth.sift_images(a.images, a.features, np.uint8, seed=synth.SEED_DATA, pool_factor=1.0)
centers, valid = synth.vocabulary_tree(a.k, a.levels, seed=4)
t = voctree.VocabularyTree(a.k, a.levels, centers, valid)
t.quantize(descs[0][:64])                                   # tree upload + warm-up
t0 = time.perf_counter()
db = voctree.Database(t)
for i in range(a.images):
    db.insert(i, descs[i])
t1 = time.perf_counter()
db.computeTfIdfWeights()
t2 = time.perf_counter()
q, ids, sc = db.find_all(a.neighbours)
t3 = time.perf_counter()
pairs = voctree.convertAllMatchesToPairList(q, ids, a.neighbours)
t4 = time.perf_counter()
tw0 = time.perf_counter()
qw, idsw, scw = db.find_all(a.neighbours, "inversedWeightedCommonPoints")
tw1 = time.perf_counter()
weighted = {"query_all_ms": 1e3 * (tw1 - tw0), "scoring_kernels_ms": db.last_gpu_ms()}
out = {"images": a.images, "features": a.features, "tree": f"K={a.k} L={a.levels} ({a.k ** a.levels} words)", "pairs": int(len(pairs)),
       "gpu": {"populate_ms": 1e3 * (t1 - t0), "tfidf_ms": 1e3 * (t2 - t1), "query_all_ms": 1e3 * (t3 - t2), "scoring_kernels_ms": db.last_gpu_ms(),
               "pair_list_ms": 1e3 * (t4 - t3), "total_ms": 1e3 * (t4 - t0), "images_per_s": a.images / (t4 - t0)},
       "gpu_inversedWeightedCommonPoints": weighted}
if a.no_cpu:
    print(json.dumps(out)); sys.exit(0)
# CPU oracle (compiled reference when present) on a bounded sample, extrapolated
kind = "ref" if oracle.VoctreeOracle.available("ref") else "port"
O = oracle.VoctreeOracle(kind)
c0 = time.perf_counter()
for i in range(a.cpu_images):
    w = O.quantize(a.k, a.levels, centers, valid, descs[i])
    assert np.array_equal(w, db.document(i))
c1 = time.perf_counter()
sub = {i: descs[i] for i in range(min(a.cpu_queries, a.images))}
O.image_matching(a.k, a.levels, centers, valid, sub, 0, a.neighbours)
c2 = time.perf_counter()
quant_per_image = (c1 - c0) / a.cpu_images
# image_matching on n images = n quantisations + n^2 sparse distances
n = len(sub)
pair_cost = max(0.0, ((c2 - c1) - n * quant_per_image)) / (n * n)
out["cpu_oracle"] = {"kind": kind, "threads": os.cpu_count(), "quantize_ms_per_image": 1e3 * quant_per_image, "sparse_distance_us_per_pair": 1e6 * pair_cost,
                     "extrapolated_total_ms": 1e3 * (a.images * quant_per_image + a.images * a.images * pair_cost),
                     "sample": f"{a.cpu_images} images quantised (words identical to the GPU's), {n} x {n} retrieval"}
out["speedup_vs_cpu_oracle"] = out["cpu_oracle"]["extrapolated_total_ms"] / out["gpu"]["total_ms"]
print(json.dumps(out))
