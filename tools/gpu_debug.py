"""GPU bring-up diagnostics (not a test): compares exact vs tensor-core raw records on small cases and prints
the first differences. Each experiment runs in its own subprocess so a device trap cannot poison the next one."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EXPERIMENTS = ["tc_small", "tc_medium", "tc_multi", "tc_big"]


def raw(m, pairs, force_exact, stage=1, variant=2):
    from alicevision_b200 import matching
    m.ctx.set_force_exact(force_exact)
    m.ctx.set_tc_variant(variant)
    pid, off, mat = m.match_uploaded(pairs, stage)
    m.ctx.set_force_exact(False)
    return {(int(pid[k, 0]), int(pid[k, 1])): mat[off[k]:off[k + 1]] for k in range(len(pid))}


def cmp_raw(a, b, tag):
    import numpy as np
    ok = True
    for k in a:
        x = np.sort(a[k], order=["j", "i"]); y = np.sort(b[k], order=["j", "i"])
        same = len(x) == len(y) and np.array_equal(x, y)
        print(f"  [{tag}] pair {k}: exact {len(x)} records, tc {len(y)} records, identical={same}")
        if not same:
            ok = False
            sx = {(int(r['i']), int(r['j'])): (float(r['dist']), float(r['ratio'])) for r in x}
            sy = {(int(r['i']), int(r['j'])): (float(r['dist']), float(r['ratio'])) for r in y}
            only_x = sorted(set(sx) - set(sy))[:8]; only_y = sorted(set(sy) - set(sx))[:8]
            print("    only exact:", [(k2, sx[k2]) for k2 in only_x])
            print("    only tc   :", [(k2, sy[k2]) for k2 in only_y])
            diffv = [(k2, sx[k2], sy[k2]) for k2 in sorted(set(sx) & set(sy)) if sx[k2] != sy[k2]][:8]
            print("    value diff:", diffv)
    return ok


def experiment(name):
    import numpy as np
    import oracle
    from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, matching, synth
    m = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_L2_B200)
    if name == "exact_vs_oracle":
        ora = oracle.best()
        descs, xys = synth.sift_images(3, 500, np.uint8, seed=1, pool_factor=1.0)
        m.ctx.set_force_exact(True)
        got = m.Match({i: (descs[i], xys[i]) for i in range(3)}, synth.exhaustive_pairs(3))
        want = ora.collection_match(descs, xys, synth.exhaustive_pairs(3), 0.8)
        print("  exact path == oracle:", got.keys() == want.keys() and all(np.array_equal(got[k], want[k]) for k in want))
        return
    sizes = {"tc_small": (2, 256), "tc_medium": (2, 1000), "tc_multi": (4, 777), "tc_big": (2, 8192)}[name]
    descs, xys = synth.sift_images(sizes[0], sizes[1], np.uint8, seed=2, pool_factor=1.0)
    m.upload({i: (descs[i], xys[i]) for i in range(sizes[0])})
    pairs = synth.exhaustive_pairs(sizes[0])
    a = raw(m, pairs, True)
    print("  exact done; launches", m.ctx.last_launches())
    for variant in (1, 2, 4):
        b = raw(m, pairs, False, variant=variant)
        print(f"  tc variant {variant} done; tc_pairs", m.ctx.last_tc_pairs(), "exactness_errors", m.ctx.exactness_errors(), "gpu_ms", m.ctx.last_gpu_ms(), "search_ms", m.ctx.last_search_kernel_ms(), flush=True)
        cmp_raw(a, b, f"{name}/v{variant}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        experiment(sys.argv[1])
    else:
        for e in EXPERIMENTS:
            print(f"== {e}", flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), e], timeout=240, capture_output=True, text=True)
                print(r.stdout[-3000:], r.stderr[-1500:] if r.returncode else "", f"(exit {r.returncode})", flush=True)
            except subprocess.TimeoutExpired:
                print("  TIMEOUT", flush=True)
