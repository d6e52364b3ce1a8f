"""Post-match filters of aliceVision_featureMatching (matching/matchesFiltering.cpp) restated in alicevision_b200/matches_filtering.py: hand-computed
known answers and properties (the reference's translation unit needs sfmData and cannot be compiled here: unpinned, host-side index arithmetic)."""
import numpy as np

from alicevision_b200 import matches_filtering as mf
from alicevision_b200.matching import MATCH_DTYPE


def _m(pairs, ratios=None):
    m = np.zeros(len(pairs), MATCH_DTYPE)
    m["i"] = [p[0] for p in pairs]; m["j"] = [p[1] for p in pairs]
    m["ratio"] = ratios if ratios is not None else np.linspace(0.1, 0.7, len(pairs))
    return m


def test_min_2d_motion():
    fi = np.array([[10, 10, 0, 0], [100, 100, 1, 0], [50, 50, 3, 0]], np.float32)        # x, y, scale, orientation
    fj = np.array([[10, 13, 0, 0], [100, 100.5, 2, 0], [70, 50, 1, 0]], np.float32)
    pm = {(0, 1): {"sift": _m([(0, 0), (1, 1), (2, 2)])}}
    mf.filterMatchesByMin2DMotion(pm, {0: {"sift": fi}, 1: {"sift": fj}}, 2.0)
    # match 0: |dp| = 3 >= 2 * 2^0 kept; match 1: 0.5 < 2 * 2^2 dropped; match 2: 20 >= 2 * 2^3 = 16 kept
    assert pm[(0, 1)]["sift"]["i"].tolist() == [0, 2]
    pm2 = {(0, 1): {"sift": _m([(0, 0), (1, 1)])}}
    mf.filterMatchesByMin2DMotion(pm2, {0: {"sift": fi}, 1: {"sift": fj}}, -1.0)          # disabled
    assert len(pm2[(0, 1)]["sift"]) == 2
    mf.filterMatchesByMin2DMotion(pm2, {0: {"sift": fi}, 1: {"sift": fj}}, 0.0)           # 0: nothing is "< 0"
    assert len(pm2[(0, 1)]["sift"]) == 2


def test_sorts_and_threshold():
    m = _m([(0, 0), (1, 1), (2, 2), (3, 3)], ratios=[0.5, 0.1, 0.5, 0.3])
    assert mf.sortMatches_byDistanceRatio(m)["i"].tolist() == [1, 3, 0, 2]              # stable on the tie
    fi = np.array([[0, 0, 1, 0], [0, 0, 4, 0], [0, 0, 2, 0], [0, 0, 2, 0]], np.float32)
    fj = np.array([[0, 0, 3, 0], [0, 0, 1, 0], [0, 0, 8, 0], [0, 0, 2, 0]], np.float32)
    # mean scales: 2, 2.5, 5, 2 -> decreasing: match 2, 1, then the tie (0, 3) in input order
    assert mf.sortMatches_byFeaturesScale(m, fi, fj)["i"].tolist() == [2, 1, 0, 3]
    assert len(mf.thresholdMatches(m, 2)) == 2 and len(mf.thresholdMatches(m, 10)) == 4


def test_grid_ordering_known_answer():
    # 300 x 300 images, 3 x 3 grid -> cells of 100 px.  As written in the reference the combined cell index is clamped to [0, gridSize - 1], so only the first
    # row of cells is distinguished; features on lower rows land in cell 2.
    lf = np.array([[10, 10, 1, 0], [20, 20, 1, 0], [150, 10, 1, 0], [250, 250, 1, 0], [30, 30, 1, 0]], np.float32)
    rf = np.array([[10, 10, 1, 0], [150, 20, 1, 0], [150, 10, 1, 0], [250, 250, 1, 0], [250, 30, 1, 0]], np.float32)
    m = _m([(0, 0), (1, 1), (2, 2), (3, 3), (4, 4)])
    # left cells: 0, 0, 1, 2(clamped from 8), 0     right cells (+9): 0, 1, 1, 2, 2
    # k=0: L0 (0) <= R0 (0) -> L0=[0]; k=1: L0 (1) > R1 (0) -> R1=[1]; k=2: L1 (0) <= R1 (1) -> L1=[2]; k=3: L2 (0) <= R2 (0) -> L2=[3]; k=4: L0 (1) > R2 (0) -> R2=[4]
    # interleave round-robin over cells L0, L1, L2, ..., R0, R1, R2: [0, 2, 3, 1, 4]
    out = mf.matchesGridFiltering(lf, (300, 300), rf, (300, 300), m, 3)
    assert out["i"].tolist() == [0, 2, 3, 1, 4]
    assert len(mf.matchesGridFiltering(lf, (300, 300), rf, (300, 300), m[:0], 3)) == 0


def test_grid_filtering_for_all_pairs_properties():
    rng = np.random.default_rng(3)
    n = 400
    feats = {v: {"sift": np.column_stack([rng.uniform(0, 4000, n), rng.uniform(0, 3000, n), rng.uniform(0.5, 30, n), rng.uniform(-3, 3, n)]).astype(np.float32)} for v in (1, 2, 3)}
    gm = {}
    for pair in ((1, 2), (1, 3), (2, 3)):
        m = np.zeros(250, MATCH_DTYPE)
        m["i"] = rng.permutation(n)[:250]; m["j"] = rng.permutation(n)[:250]; m["ratio"] = rng.random(250)
        gm[pair] = {"sift": m}
    sizes = {1: (4000, 3000), 2: (4000, 3000), 3: (4000, 3000)}
    for grid in (False, True):
        out = mf.matchesGridFilteringForAllPairs(gm, sizes, feats, grid, 0)
        for pair in gm:
            a, b = out[pair]["sift"], gm[pair]["sift"]
            assert len(a) == len(b) and sorted(zip(a["i"].tolist(), a["j"].tolist())) == sorted(zip(b["i"].tolist(), b["j"].tolist()))   # a permutation
            if not grid:
                s = (feats[pair[0]]["sift"][a["i"], 2] + feats[pair[1]]["sift"][a["j"], 2]) / 2
                assert np.all(np.diff(s) <= 1e-6)                                                                                      # decreasing scale
        cut = mf.matchesGridFilteringForAllPairs(gm, sizes, feats, grid, 100)
        assert all(len(cut[p]["sift"]) == 100 and np.array_equal(cut[p]["sift"], out[p]["sift"][:100]) for p in gm)
