"""Pipeline trace of CTA 0 of the tensor-core kernel (debug): prints per-tile intervals in SM cycles."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, matching, synth

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ablate = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n_img = 12
descs, xys = synth.sift_images(n_img, 8192, np.float32, seed=3, pool_factor=1.0)
m = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_L2_B200)
m.ctx.set_tc_variant(variant)
m.upload({i: (descs[i], xys[i]) for i in range(n_img)})
pairs = synth.exhaustive_pairs(n_img)
m.match_uploaded(pairs, matching.STAGE_DEVICE)          # warm
m.ctx.lib.b200m_debug_trace(m.ctx._h, 1 | (ablate << 8), None, 0)
m.match_uploaded(pairs, matching.STAGE_DEVICE)
tr = np.zeros((4, 512, 4), np.int64); m.ctx.lib.b200m_debug_trace(m.ctx._h, 1 | (ablate << 8), tr.ctypes.data, tr.size)
print(f"variant {variant} ablate {ablate}: gpu_ms {m.ctx.last_gpu_ms():.3f} search_ms {m.ctx.last_search_kernel_ms():.3f} pairs {len(pairs)}")
prod, mma, e0, e1 = tr[0], tr[1], tr[2], tr[3]
lo, hi = 64, 192     # steady-state tiles
print("per-tile period (MMA commit issued -> next):", np.diff(mma[lo:hi, 2]).mean())
print("MMA: barriers ready -> commit issued (issue loop incl. next-tile polls):", (mma[lo:hi, 2] - mma[lo:hi, 1]).mean())
print("MMA: commit issued(t-1) -> barriers ready(t) (gap):", (mma[lo + 1:hi, 1] - mma[lo:hi - 1, 2]).mean())
print("MMA: loop top -> barriers observed:", (mma[lo:hi, 1] - mma[lo:hi, 0]).mean(), " commit issued(t-1) -> loop top(t):", (mma[lo + 1:hi, 0] - mma[lo:hi - 1, 2]).mean())
flags = prod[lo:hi, 2]
print("look-ahead polls: db_full ready %.0f %%, tm_empty ready %.0f %%" % (100 * (flags & 1).mean(), 100 * ((flags >> 1) & 1).mean()))
print("producer: slot free -> issued:", (prod[lo:hi, 1] - prod[lo:hi, 0]).mean(), " period:", np.diff(prod[lo:hi, 0]).mean())
for name, e in (("epi half0", e0), ("epi half1", e1)):
    print(f"{name}: wait tm_full: {(e[lo:hi, 1] - e[lo:hi, 0]).mean():.0f}  process: {(e[lo:hi, 2] - e[lo:hi, 1]).mean():.0f}  period: {np.diff(e[lo:hi, 2]).mean():.0f}")
print("tm_full latency: MMA commit issued(t) -> epilogue sees tm_full(t):", (e0[lo:hi, 1] - mma[lo:hi, 2]).mean())
print("producer lead: TMA issued(t) -> MMA barriers ready(t):", (mma[lo:hi, 1] - prod[lo:hi, 1]).mean())
