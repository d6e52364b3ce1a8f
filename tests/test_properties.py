"""Property-based tests (hypothesis) of the host-side logic around the path: pair sharding, pair-list text IO, .feat text round trip,
convertAllMatchesToPairList against a literal Python transliteration of imageMatching/ImageMatching.cpp:107-143.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from alicevision_b200 import matching, pairs as pairs_io, regions_io as rio, voctree

pair_lists = st.lists(st.tuples(st.integers(0, 60), st.integers(0, 60)), min_size=0, max_size=200)


@settings(max_examples=60, deadline=None)
@given(pair_lists, st.integers(1, 9))
def test_shard_pairs_is_a_partition_by_database_image(pl, world):
    p = np.array(pl, np.uint32).reshape(-1, 2)
    s = matching.shard_pairs(p, world)
    assert len(s) == len(p) and (len(s) == 0 or (s.min() >= 0 and s.max() < world))
    owner = {}
    for (i, _), k in zip(p.tolist(), s.tolist()):
        assert owner.setdefault(i, k) == k                         # all pairs of one database image on one shard
    firsts = sorted(owner)
    # round robin over ascending first ids, direction alternating every round
    for k, f in enumerate(firsts):
        rnd, pos = divmod(k, world)
        assert owner[f] == (pos if rnd % 2 == 0 else world - 1 - pos)


@settings(max_examples=60, deadline=None)
@given(pair_lists)
def test_pair_list_text_round_trip(pl):
    pl = [(a, b) for a, b in pl if a != b]
    text = pairs_io.savePairs(pl)
    back = pairs_io.loadPairs(text) if text else []
    # savePairs keeps (I, J) as given; loadPairs orders each pair I < J (ImagePairListIO.cpp:57)
    assert back == sorted({(min(a, b), max(a, b)) for a, b in pl})


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(*[st.floats(-1e6, 1e6, allow_nan=False, width=32)] * 4), min_size=0, max_size=50))
def test_feat_text_round_trip_is_the_ostream_format(rows):
    import os, tempfile
    f = np.array(rows, np.float32).reshape(-1, 4)
    with tempfile.TemporaryDirectory() as t:
        path = os.path.join(t, "a.feat")
        rio.saveFeatsToFile(path, f)
        text = open(path).read()
        back = rio.loadFeatsFromFile(path)
    want_text = "".join("%g %g %g %g\n" % tuple(float(v) for v in r) for r in f)      # default ostream format == %g
    assert text == want_text
    want = np.array([[np.float32(float("%g" % float(v))) for v in r] for r in f], np.float32).reshape(-1, 4)
    assert np.array_equal(back.view(np.uint32), want.view(np.uint32))


def _convert_reference(all_matches, num_matches):
    """imageMatching/ImageMatching.cpp:107-143, literally."""
    out = {}
    if num_matches == 0:
        num_matches = len(all_matches)
    for curr in sorted(all_matches):
        best = set()
        for m in all_matches[curr]:
            if m == curr:
                continue
            if m < curr:
                if m in out and curr not in out[m]:
                    best.add(m)
            else:
                best.add(m)
            if len(best) == num_matches:
                break
        if best:
            out[curr] = best
    return [(i, j) for i in sorted(out) for j in sorted(out[i])]


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 12), st.integers(0, 6), st.randoms(use_true_random=False))
def test_convert_all_matches_to_pair_list(n, num_matches, rnd):
    ids = sorted(rnd.sample(range(100), n))
    keep = n if num_matches == 0 else min(num_matches, n)
    match_ids = np.array([rnd.sample(ids, keep) for _ in ids], np.uint32).reshape(n, keep)
    got = voctree.convertAllMatchesToPairList(np.array(ids, np.uint32), match_ids, num_matches)
    want = _convert_reference({i: match_ids[k].tolist() for k, i in enumerate(ids)}, num_matches)
    assert [tuple(p) for p in got.tolist()] == want


def test_checked_f32_to_u8_conversion_of_the_staging_path():
    """hostconv.hpp through its test hook (host code, no GPU): the AVX2 implementation and the scalar one agree on acceptance and on the bytes, for every
    length (vector body + tail), and reject anything that is not an integer in 0..255 wherever it sits."""
    import ctypes as C
    from alicevision_b200 import matching
    lib = matching.load_library()
    lib.b200m_debug_convert_f32_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rng = np.random.default_rng(11)

    def conv(a, which):
        out = np.zeros(max(a.size, 1), np.uint8)
        rc = lib.b200m_debug_convert_f32_u8(a.ctypes.data, out.ctypes.data, a.size, which)
        return rc, out[: a.size]

    for n in (0, 1, 31, 32, 33, 127, 1024, 1025, 8192 * 128):
        a = rng.integers(0, 256, n).astype(np.float32)
        for which in (0, 1):
            rc, out = conv(a, which)
            assert rc == 1 and np.array_equal(out, a.astype(np.uint8))
    base = rng.integers(0, 256, 4099).astype(np.float32)
    for bad in (0.5, -1.0, 256.0, 255.5, 1e20, -1e20, np.nan, np.inf, 1e-30):
        for pos in (0, 7, 31, 32, 1000, 2047, 4096, 4098):
            a = base.copy(); a[pos] = bad
            assert conv(a, 0)[0] == 0 and conv(a, 1)[0] == 0, (bad, pos)
    a = base.copy(); a[5] = -0.0; a[6] = 255.0; a[7] = 0.0          # -0.0 is the integer 0
    for which in (0, 1):
        rc, out = conv(a, which)
        assert rc == 1 and out[5] == 0 and out[6] == 255
