"""Guided matching at BASELINE size (8192 x 8192 features, rectified synthetic pair): GPU kernel time next to the CPU oracle."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from alicevision_b200 import Regions, matching
from test_guided import F_RECT, scene
out = []
NO_CPU = "--no-cpu" in sys.argv
for kind in ("u8", "f32", "bin"):
    dl, xl, dr, xr, truth = scene(kind, 8192, seed=11)
    L, R = Regions(dl, xl, binary=kind == "bin"), Regions(dr, xr, binary=kind == "bin")
    ctx = matching.default_context()
    matching.guidedMatching(F_RECT, L, R, 4.0, 0.64)
    t0 = time.perf_counter(); got = matching.guidedMatching(F_RECT, L, R, 4.0, 0.64); t1 = time.perf_counter()
    gpu_ms = ctx.last_gpu_ms()
    c0 = c1 = 0.0
    if not NO_CPU:
        O = oracle.best(); O.set_num_threads(1)
        sub = np.arange(0, 8192, 16)
        c0 = time.perf_counter(); O.guided_match(dl[sub], xl[sub], dr, xr, F_RECT, 4.0, 0.64, binary=kind == "bin"); c1 = time.perf_counter()
    out.append({"descriptors": kind, "left x right": "8192 x 8192", "matches": int(len(got)), "gpu_kernel_ms": gpu_ms, "call_ms_incl_upload": 1e3 * (t1 - t0),
                "cpu_oracle_ms_1_thread_extrapolated": 1e3 * (c1 - c0) * 16, "candidates_per_left_feature": float(np.mean(np.abs(xr[:, 1][None, :2048] - xl[:512, 1][:, None]) < 2.0) * 8192)})
print(json.dumps(out))
