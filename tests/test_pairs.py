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
