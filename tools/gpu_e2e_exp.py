"""End-to-end step time (clear + async upload + match FULL) and upload-alone time under engine switches (debug / profiles).
Each variant runs on its own engine context (the switches are read at context creation; B200M_UP_* per upload job)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, matching, synth
n_img = int(os.environ.get("EXP_IMAGES", 100)); M = int(os.environ.get("EXP_FEATURES", 8192))
descs, xys = synth.sift_images(n_img, M, np.float32, seed=synth.SEED_DATA, pool_factor=1.0)
pairs = synth.exhaustive_pairs(n_img)
views = {i: (descs[i], xys[i]) for i in range(n_img)}
VARIANTS = [
    ("defaults", {}),
    ("no u8 staging", {"B200M_U8_STAGING": "0"}),
    ("host finishing", {"B200M_DEVICE_FINISH": "0"}),
    ("r01 behaviour (no u8, host finishing)", {"B200M_U8_STAGING": "0", "B200M_DEVICE_FINISH": "0"}),
    ("lag 6", {"B200M_UP_LAG": "6"}),
    ("lag 20", {"B200M_UP_LAG": "20"}),
    ("host threads 8", {"B200M_HOST_THREADS": "8"}),
    ("host threads 16", {"B200M_HOST_THREADS": "16"}),
    ("numa off", {"B200M_NUMA": "0"}),
]
only = os.environ.get("EXP_ONLY")
for name, env in VARIANTS:
    if only and only not in name:
        continue
    keys = ["B200M_U8_STAGING", "B200M_DEVICE_FINISH", "B200M_UP_LAG", "B200M_HOST_THREADS", "B200M_NUMA"]
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = matching.Context(0)
    m = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_L2_B200, ctx)
    m.Match(views, pairs)
    ts, up, res = [], [], []
    for rep in range(5):
        t0 = time.perf_counter(); m.clear(); out = m.Match(views, pairs); n = out.num_matches(); ts.append(time.perf_counter() - t0); del out
        m.clear(); t0 = time.perf_counter(); m.upload(views); m.wait_uploads(); up.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); m.match_uploaded(pairs, matching.STAGE_FULL); res.append(time.perf_counter() - t0)
    print(f"{name:40s}: e2e step {1e3*min(ts):6.1f} ms (median {1e3*sorted(ts)[2]:6.1f}) = {len(pairs)/min(ts):8.0f} pairs/s | upload alone {1e3*min(up):6.1f} ms | "
          f"match on resident views {1e3*min(res):6.1f} ms (gpu {ctx.last_gpu_ms():.1f}, search kernels {ctx.last_search_kernel_ms():.1f})", flush=True)
    ctx.close()
