import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, synth
rng = np.random.default_rng(0)
descs = [rng.integers(0, 200, (8192, 128)).astype(np.float32) for _ in range(100)]
xys = [synth.positions(8192, rng) for _ in range(100)]
views = {i: (descs[i], xys[i]) for i in range(100)}
m = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_L2_B200)
t0 = time.perf_counter(); big = np.empty((100, 8192, 128), np.float32)
for i in range(100): big[i] = descs[i]
print("numpy single-thread memcpy of 419 MB: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
for rep in range(3):
    m.clear(); t0 = time.perf_counter(); m.upload(views); m.ctx.lib.b200m_clear_views  # noqa
    print("upload %.1f ms" % (1e3 * (time.perf_counter() - t0)))
