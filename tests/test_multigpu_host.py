"""CPU test of the N>1 host logic: two gloo ranks shard the pair list exactly as bench.py does under torchrun,
and the union of the shards is the whole list, disjoint and balanced, with no data-path collective needed."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_img, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    from alicevision_b200 import synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pairs = synth.exhaustive_pairs(n_img)
    mine = bench.shard_pairs(pairs, rank, world, "rows")
    # what bench.py reduces: max of per-rank times, sum of per-rank pair counts
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    n = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    q.put((rank, mine.tolist(), t.item(), n.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharding_two_ranks_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_img, world = 23, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_img, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {r: set(map(tuple, m)) for r, m, _, _ in got}
    total = n_img * (n_img - 1) // 2
    assert shards[0].isdisjoint(shards[1]) and len(shards[0] | shards[1]) == total
    assert abs(len(shards[0]) - len(shards[1])) <= n_img          # balanced within one row
    assert all(t == 2.0 and n == total for _, _, t, n in got)     # MAX / SUM reductions
    # every database image (first index) lives on exactly one rank -> its descriptors are reused from L2
    for r in (0, 1):
        assert {i for i, _ in shards[r]}.isdisjoint({i for i, _ in shards[1 - r]})


def test_shard_sizes_weak_scaling_table():
    sys.path.insert(0, ROOT)
    import bench
    from alicevision_b200 import synth
    for world, n_img in bench.IMAGES_FOR_GPUS.items():
        pairs = synth.exhaustive_pairs(n_img)
        sizes = [len(bench.shard_pairs(pairs, r, world, "rows")) for r in range(world)]
        assert sum(sizes) == len(pairs)
        assert max(sizes) - min(sizes) <= n_img
        assert abs(np.mean(sizes) - 4950) / 4950 < 0.02


def test_c_sharding_properties():
    """b200m_shard_pairs (host function of the C ABI, used by bench.py under torchrun and by b200m_multi_match in-process):
    disjoint cover, all pairs of one database image on one shard, balanced for exhaustive and for sparse lists."""
    sys.path.insert(0, ROOT)
    from alicevision_b200 import matching, synth
    for n_img, world in ((23, 2), (100, 4), (282, 8), (7, 8)):
        pairs = synth.exhaustive_pairs(n_img)
        s = matching.shard_pairs(pairs, world)
        assert s.min() >= 0 and s.max() < world and len(s) == len(pairs)
        owner = {}
        for (i, _), k in zip(pairs.tolist(), s.tolist()):
            assert owner.setdefault(i, k) == k
        sizes = np.bincount(s, minlength=world)
        if n_img > 2 * world:
            assert sizes.max() - sizes.min() <= n_img
    vt = synth.voctree_like_pairs(300, k=20)
    s = matching.shard_pairs(vt, 8)
    sizes = np.bincount(s, minlength=8)
    assert sizes.sum() == len(vt) and sizes.max() < 1.35 * sizes.mean()
    assert len(matching.shard_pairs(np.zeros((0, 2), np.uint32), 4)) == 0


def test_2d_block_sharding_properties():
    """b200m_shard_pairs_2d (what bench.py and b200m_multi_match use): disjoint cover, balanced, and a shard needs only part of
    the views - half of them at 8 shards, at most three quarters at 4 - where sharding by database image needs nearly all."""
    sys.path.insert(0, ROOT)
    import bench
    from alicevision_b200 import matching, synth
    for world, n_img in bench.IMAGES_FOR_GPUS.items():
        pairs = synth.exhaustive_pairs(n_img)
        s = matching.shard_pairs_2d(pairs, world)
        assert s.min() >= 0 and s.max() < world and len(s) == len(pairs)
        sizes = np.bincount(s, minlength=world)
        assert sizes.sum() == len(pairs) and sizes.max() <= 1.03 * sizes.mean()
        assert abs(sizes.mean() - 4950) / 4950 < 0.02
        views = [len(np.unique(pairs[s == r])) for r in range(world)]
        limit = {1: 1.0, 2: 1.0, 4: 0.76, 8: 0.51}[world]
        assert max(views) <= limit * n_img + 1, (world, views)
        mine = [bench.shard_pairs(pairs, r, world, "2d") for r in range(world)]
        assert sum(len(m) for m in mine) == len(pairs)
        for m in mine:                                   # PairSet order inside a shard: the pairs of one database image stay adjacent
            assert m.tolist() == sorted(m.tolist())
    vt = synth.voctree_like_pairs(1000, k=50)            # BASELINE configs[2]
    for world, frac in ((4, 0.76), (8, 0.51)):
        s = matching.shard_pairs_2d(vt, world)
        sizes = np.bincount(s, minlength=world)
        assert sizes.sum() == len(vt) and sizes.max() <= 1.03 * sizes.mean()
        assert max(len(np.unique(vt[s == r])) for r in range(world)) <= frac * 1000 + 1
    # degenerate inputs: more shards than images, one pair, no pair
    assert matching.shard_pairs_2d(np.array([[3, 9]], np.uint32), 8).tolist()[0] in range(8)
    assert len(matching.shard_pairs_2d(np.zeros((0, 2), np.uint32), 4)) == 0
    s = matching.shard_pairs_2d(synth.exhaustive_pairs(5), 8)
    assert len(s) == 10 and s.min() >= 0 and s.max() < 8
