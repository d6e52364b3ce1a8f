"""e2e step time vs upload staging parameters (debug): B200M_UP_PARTS / B200M_UP_CHUNK_MB / B200M_UP_LAG are read per upload job."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, matching, synth
n_img = 100
descs, xys = synth.sift_images(n_img, 8192, np.float32, seed=synth.SEED_DATA, pool_factor=1.0)
pairs = synth.exhaustive_pairs(n_img)
views = {i: (descs[i], xys[i]) for i in range(n_img)}
m = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_L2_B200)
m.Match(views, pairs)
for parts, chunk, lag in ((1, 4, 4), (1, 4, 8), (1, 4, 12), (1, 4, 16), (1, 4, 22), (1, 2, 16), (1, 2, 22), (1, 8, 8), (2, 4, 8)):
    os.environ["B200M_UP_PARTS"] = str(parts); os.environ["B200M_UP_CHUNK_MB"] = str(chunk); os.environ["B200M_UP_LAG"] = str(lag)
    ts, up = [], []
    for rep in range(4):
        t0 = time.perf_counter(); m.clear(); out = m.Match(views, pairs); ts.append(time.perf_counter() - t0); del out
        m.clear(); t0 = time.perf_counter(); m.upload(views); m.wait_uploads(); up.append(time.perf_counter() - t0)
    print(f"parts {parts} chunk {chunk} MB lag {lag}: e2e step {1e3*min(ts):.1f} ms (median {1e3*sorted(ts)[2]:.1f}) | upload alone {1e3*min(up):.1f} ms = {0.4194/min(up):.1f} GB/s")
