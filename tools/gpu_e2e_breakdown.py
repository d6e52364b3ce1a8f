"""Where does the end-to-end time go? (debug)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, matching, synth
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_feat = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
descs, xys = synth.sift_images(n_img, n_feat, np.float32, seed=synth.SEED_DATA, pool_factor=1.0)
pairs = synth.exhaustive_pairs(n_img)
views = {i: (descs[i], xys[i]) for i in range(n_img)}
m = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_L2_B200)
for rep in range(3):
    m.clear()
    t0 = time.perf_counter(); m.upload(views); t1 = time.perf_counter()
    m.match_uploaded(pairs, matching.STAGE_DEVICE); t2 = time.perf_counter()
    m.match_uploaded(pairs, matching.STAGE_RAW); t3 = time.perf_counter()
    pid, off, mat = m.match_uploaded(pairs, matching.STAGE_FULL); t4 = time.perf_counter()
    print(f"rep {rep}: upload {1e3*(t1-t0):.1f} ms | device {1e3*(t2-t1):.1f} | raw {1e3*(t3-t2):.1f} | full {1e3*(t4-t3):.1f} | gpu_ms {m.ctx.last_gpu_ms():.1f} | matches {len(mat)} records {m.ctx.last_records()}")
print("--- e2e steps (clear + Match from host buffers, uploads overlapped)")
for rep in range(3):
    t0 = time.perf_counter(); m.clear(); t1 = time.perf_counter(); m.upload(views); t2 = time.perf_counter()
    pid, off, mat = m.match_uploaded(pairs, matching.STAGE_FULL); t3 = time.perf_counter()
    offs = off.tolist(); out = {}
    for (i, j), a, b in zip(pid.tolist(), offs[:-1], offs[1:]):
        if b > a: out[(i, j)] = mat[a:b]
    t4 = time.perf_counter()
    print(f"e2e rep {rep}: clear {1e3*(t1-t0):.1f} ms | upload call {1e3*(t2-t1):.1f} | match {1e3*(t3-t2):.1f} | dict {1e3*(t4-t3):.1f} | total {1e3*(t4-t0):.1f} | gpu_ms {m.ctx.last_gpu_ms():.1f}")
    del out, pid, off, mat
u8 = {i: (descs[i].astype(np.uint8), xys[i]) for i in range(n_img)}
m.clear(); t0 = time.perf_counter(); m.upload(u8); m.wait_uploads(); print("upload u8:", 1e3 * (time.perf_counter() - t0), "ms")
