"""Pair-list construction / IO against the reference's own tests
(matchingImageCollection/pairBuilder_test.cpp:35-60, ImagePairListIO_test.cpp:18-60)."""
import os
import tempfile

from alicevision_b200 import pairs as P


def test_exhaustive_pairs_reference_case():
    ids = [12, 54, 89, 65]
    got = P.exhaustivePairs(ids)
    assert len(got) == 6 and all(a < b for a, b in got) and set(got) == {(12, 54), (12, 65), (12, 89), (54, 65), (54, 89), (65, 89)}
    # range: only pairs whose first image is in the chunk (pairBuilder.cpp:28-36)
    assert P.exhaustivePairs(ids, 1, 2) == [(54, 65), (54, 89), (65, 89)]
    assert P.exhaustivePairs(ids, 7, 2) == []
    assert P.exhaustivePairs(ids, 0, 1) == [(12, 54), (12, 65), (12, 89)]


def test_save_pairs_golden_string():
    assert P.savePairs({(0, 2), (0, 4), (0, 5), (8, 2), (0, 1), (5, 9)}) == "0 1 2 4 5\n5 9\n8 2\n"
    assert P.savePairs([]) == ""


def test_load_multiple_pairs_per_line_and_normalisation():
    assert P.loadPairs(" 0 2 4 5\n        0 1\n        5 9\n") == sorted({(0, 2), (0, 4), (0, 5), (0, 1), (5, 9)})
    assert P.loadPairs("3 1\n") == [(1, 3)]              # normalised to I<J (ImagePairListIO.cpp:57)
    assert P.loadPairs("4 4\n") is None                   # image sees itself
    assert P.loadPairs("7\n") is None                     # fewer than two ids
    assert P.loadPairs("0 1\n2 3\n4 5\n", 1, 1) == [(2, 3)]


def test_roundtrip_file():
    gt = {(0, 1), (1, 2), (2, 0)}
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "pairsT_IO.txt")
        assert P.savePairsToFile(f, gt)
        assert P.loadPairsFromFile(f) == [(0, 1), (0, 2), (1, 2)]


def test_other_pair_generators():
    """imageMatching/ImageMatching.cpp:145-206 for plain view ids."""
    from alicevision_b200 import pairs as P
    paths = {10: "/d/img_003.jpg", 11: "/d/img_001.jpg", 12: "/d/img_002.jpg", 7: "/d/img_004.jpg"}      # path order: 11, 12, 10, 7
    assert P.generateSequentialMatches(paths, 1) == [(7, 10), (10, 12), (11, 12)]
    assert P.generateSequentialMatches(paths, 2) == [(7, 10), (7, 12), (10, 11), (10, 12), (11, 12)]
    assert P.generateSequentialMatches(paths, 0) == [] and P.generateSequentialMatches({}, 3) == []
    assert P.generateAllMatchesInOneMap([5, 1, 3]) == [(1, 3), (1, 5), (3, 5)] == P.exhaustivePairs([5, 1, 3])
    assert P.generateAllMatchesBetweenTwoMap([2, 9], [1, 9]) == [(2, 1), (2, 9), (9, 1), (9, 9)]


def test_match_filters():
    """matching/io.cpp:82-130."""
    import numpy as np
    import pytest
    from alicevision_b200 import regions_io as rio
    from alicevision_b200.matching import MATCH_DTYPE
    def mk(n):
        m = np.zeros(n, MATCH_DTYPE); m["i"] = np.arange(n); m["j"] = np.arange(n)[::-1]
        return m
    pm = {(0, 1): {"sift": mk(30), "akaze": mk(4)}, (0, 2): {"sift": mk(7)}, (2, 3): {"akaze": mk(50)}}
    a = {k: dict(v) for k, v in pm.items()}
    rio.filterMatchesByViews(a, {0, 1, 2})
    assert sorted(a) == [(0, 1), (0, 2)]
    b = {k: dict(v) for k, v in pm.items()}
    rio.filterTopMatches(b, 20, 5)
    assert len(b[(0, 1)]["sift"]) == 20 and np.array_equal(b[(0, 1)]["sift"]["i"], np.arange(20)) and len(b[(0, 1)]["akaze"]) == 0
    assert len(b[(0, 2)]["sift"]) == 7 and len(b[(2, 3)]["akaze"]) == 20
    c = {k: dict(v) for k, v in pm.items()}
    rio.filterTopMatches(c, 0, 0)
    assert all(len(c[k][d]) == len(pm[k][d]) for k in pm for d in pm[k])
    with pytest.raises(RuntimeError):
        rio.filterTopMatches(c, 5, 10)
    d = {k: dict(v) for k, v in pm.items()}
    rio.filterMatchesByDesc(d, ["sift"])
    assert sorted(d) == [(0, 1), (0, 2)] and list(d[(0, 1)]) == ["sift"]
