#!/usr/bin/env python
"""bench.py — image-pairs matched/sec on the descriptor-matching hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--features M] [--images I]
                    [--dtype f32|u8|bin] [--cpu-seconds S]

One "step" = one pass of the hot path over the rank's shard of the pair list (N=1: BASELINE configs[1],
100 synthetic images x 8192 SIFT features, exhaustive 4950 pairs).  Printed by rank 0 as ONE JSON line:

  value        whole-job pairs/s, descriptors resident in HBM, CUDA-event time of the enqueued work (max over ranks)
  e2e          the same pairs through the reference-facing C-ABI call chain with HOST buffers:
               b200m_upload_view (H2D) + b200m_match_pairs(STAGE_FULL) (kernels, D2H, host finishing)
  roofline     dominant kernel (tcgen05 distance GEMM + fused top-2): 2*M^2*128 FLOP per pair / its own device time
  cpu_baseline the reference's CPU brute force (oracle/_ref, else the port) on a bounded sample of the same pairs

Multi-GPU (torchrun, one rank per GPU): pairs are independent, so the pair list is dealt round-robin by
database image and there is NO collective on the data path; torch.distributed is used for the barrier and the
max-over-ranks reduction of the timings only.  Scaling is weak: pairs per GPU stay ~4950 as N grows.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from alicevision_b200 import synth  # noqa: E402

IMAGES_FOR_GPUS = {1: 100, 2: 141, 4: 199, 8: 282}   # exhaustive pairs ~ 4950 * N


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--features", type=int, default=8192)
    ap.add_argument("--images", type=int, default=0, help="0 = 100 per GPU-equivalent (IMAGES_FOR_GPUS)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "u8", "bin"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--tc-variant", type=int, default=4, choices=[1, 2, 3, 4], help="4 = CTA-pair tcgen05 kernel, half-norms folded into the GEMM (default); 2/3 = CTA pair with epilogue add (8/16 epilogue warps); 1 = single-CTA kernel")
    ap.add_argument("--pairs", default="exhaustive", choices=["exhaustive", "voctree"],
                    help="voctree = BASELINE configs[2]: synthetic vocabulary-tree style list, 50 neighbours per image (use with --images 1000)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def make_workload(args, world):
    n_img = args.images or IMAGES_FOR_GPUS.get(world, 100 * world)
    if args.dtype == "bin":
        descs, xys = synth.mldb_images(n_img, args.features, seed=synth.SEED_DATA)
    else:
        descs, xys = synth.sift_images(n_img, args.features, np.float32 if args.dtype == "f32" else np.uint8, seed=synth.SEED_DATA, pool_factor=1.0)
    pairs = synth.voctree_like_pairs(n_img, k=50) if args.pairs == "voctree" else synth.exhaustive_pairs(n_img)
    return descs, xys, pairs


def shard_pairs(pairs: np.ndarray, rank: int, world: int) -> np.ndarray:
    """This rank's shard: database images (first index, as ImageCollectionMatcher_generic groups them, .cpp:45-50) dealt
    round-robin over ranks, direction alternating every round so the triangular row lengths balance.  Same host function
    (b200m_shard_pairs) the single-process multi-GPU path uses."""
    if world == 1:
        return pairs
    from alicevision_b200 import matching
    return pairs[matching.shard_pairs(pairs, world) == rank]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def host_cores() -> int:
    """Physical cores this process may run on ("all physical cores"); hyper-threads only slow the reference's
    compute-bound OpenMP loop down.  Falls back to the affinity count."""
    allowed = os.sched_getaffinity(0)
    try:
        cores, cpu, phys = set(), None, None
        for line in open("/proc/cpuinfo"):
            k, _, v = line.partition(":")
            k = k.strip()
            if k == "processor":
                cpu, phys = int(v), None
            elif k == "physical id":
                phys = int(v)
            elif k == "core id" and cpu in allowed:
                cores.add((phys, int(v)))
        return len(cores) or len(allowed)
    except Exception:
        return len(allowed)


def cpu_baseline(descs, xys, pairs, hamming, budget_s):
    """The reference's CPU brute force (+ ratio test + de-duplication) on a bounded sample of the same pair list."""
    import oracle
    ora = oracle.best()
    ora.set_num_threads(host_cores())   # torchrun exports OMP_NUM_THREADS=1; the baseline gets every physical core this process may use
    n = 0
    t0 = time.perf_counter()
    while n < len(pairs) and (time.perf_counter() - t0) < budget_s:
        i, j = int(pairs[n, 0]), int(pairs[n, 1])
        ora.regions_match(descs[i], xys[i], descs[j], xys[j], 0.8, hamming)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": ora.num_threads(), "kind": "reference" if ora.kind == "ref" else "port",
            "sample": f"first {n} pairs of the same list, {dt:.1f} s, ArrayMatcher_bruteForce + ratio test + de-duplication, "
                      f"{'g++ -O3 -msse2 -fopenmp on the reference headers' if ora.kind == 'ref' else 'C++ port of the reference'}"}


def cpu_baseline_cascade(descs, xys, pairs, budget_s):
    """The reference's CASCADE_HASHING_L2 matcher (matching/ArrayMatcher_cascadeHashing.hpp + CascadeHasher.hpp compiled from the
    reference tree into oracle/_ref) through the restated collection loop: one hashed database per image I, OpenMP over its J images
    (ImageCollectionMatcher_generic.cpp:39,68).  Whole database rows of the same pair list until the budget is used.  None when
    the compiled reference is not available (the port does not restate cascade hashing)."""
    import oracle
    if not oracle.available("ref"):
        return None
    ora = oracle.Oracle("ref")
    ora.set_num_threads(host_cores())
    firsts = np.unique(pairs[:, 0])
    n = 0; matches = 0
    t0 = time.perf_counter()
    for f in firsts:
        row = pairs[pairs[:, 0] == f]
        tot, _ = ora.collection_cascade(descs, xys, row, 0.8)
        n += len(row); matches += tot
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": ora.num_threads(), "kind": "reference",
            "sample": f"first {n} pairs (whole database rows) of the same list, {dt:.1f} s, ArrayMatcher_cascadeHashing + ratio test + de-duplication, database "
                      f"hashed once per image I, OpenMP over J as in ImageCollectionMatcher_generic; {matches} matches; g++ -O3 -msse2 -fopenmp on the "
                      "reference headers with plain-loop stand-ins for Eigen's MatrixXf/VectorXf (timing baseline, results unpinned)"}


def ncu_traffic(kernel_key: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(kernel_key)
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1449.7), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
    return 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    hamming = args.dtype == "bin"
    metric = "image-pairs matched/sec (" + ("AKAZE-MLDB 64-byte, " if hamming else "SIFT 128-D, ") + f"{args.features} feat/img)"

    if args.impl == "reference":
        # Reference arm: the reference's own CPU implementation on the box's host cores, rank 0 only.
        if rank != 0:
            return
        descs, xys, pairs = make_workload(argparse.Namespace(**{**vars(args), "images": args.images or 24}), 1)
        per_step = []
        for s in range(args.warmup + args.steps):
            b = cpu_baseline(descs, xys, pairs[s % 4::4], hamming, max(2.0, args.cpu_seconds / 2))   # ~6 s of CPU work per step
            if s >= args.warmup:
                per_step.append(b)
        v = float(np.mean([b["value"] for b in per_step])) if per_step else 0.0
        cb = dict(per_step[-1]) if per_step else {"kind": "port", "cores": 0, "sample": ""}
        cb["value"] = v
        print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32" if args.dtype == "f32" else args.dtype, "data": "synthetic",
                          "config": {"workload": f"bounded sample of: {args.features} feat/img exhaustive pairs, BRUTE_FORCE_{'HAMMING' if hamming else 'L2'} on CPU"},
                          "cpu_baseline": cb, "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    from alicevision_b200 import EMatcherType, ImageCollectionMatcherB200, matching

    if world > 1:
        # one engine process per GPU shares the host: split the cores between the ranks' finishing pools and keep fewer staging
        # copies in flight per rank (4 ranks: e2e 160 k -> 178 k pairs/s, profiles/r01d_multi_rank_host_settings.md)
        os.environ.setdefault("B200M_HOST_THREADS", str(max(12, host_cores() // world)))   # 12 was already too few at 4 ranks
        os.environ.setdefault("B200M_UP_LAG", "6")

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    descs, xys, pairs = make_workload(args, world)
    mine = shard_pairs(pairs, rank, world)
    ctx = matching.Context(local)
    ctx.set_tc_variant(args.tc_variant)
    m = ImageCollectionMatcherB200(0.8, False, EMatcherType.BRUTE_FORCE_HAMMING_B200 if hamming else EMatcherType.BRUTE_FORCE_L2_B200, ctx)
    views = {i: (descs[i], xys[i]) for i in range(len(descs))}
    m.upload(views)

    # ---- value: descriptors resident in HBM, kernel-boundary throughput --------------------------------------
    for _ in range(args.warmup):
        m.match_uploaded(mine, matching.STAGE_DEVICE)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0w = time.time()
    gpu_ms = 0.0; search_ms = 0.0; launches = 0; records = 0
    for _ in range(args.steps):
        m.match_uploaded(mine, matching.STAGE_DEVICE)
        gpu_ms += ctx.last_gpu_ms(); search_ms += ctx.last_search_kernel_ms(); launches += ctx.last_launches(); records += ctx.last_records()
    barrier()
    t1w = time.time()
    clocks = sampler.stop(t0w, t1w) if sampler else None
    tc_pairs = ctx.last_tc_pairs(); errs = ctx.exactness_errors()

    # ---- e2e: host buffers -> upload -> match -> D2H -> host finishing ------------------------------------------
    e2e_s = 0.0; h2d = 0; d2h = 0
    if not args.no_e2e:
        m.clear(); m.Match(views, mine)            # warm-up of the full chain
        barrier()
        te = time.perf_counter()
        for _ in range(max(1, min(args.steps, 3))):
            m.clear()
            res = m.Match(views, mine)
        barrier()
        e2e_steps = max(1, min(args.steps, 3))
        e2e_s = (time.perf_counter() - te) / e2e_steps
        h2d = int(sum(d.nbytes + x.nbytes * 0 for d, x in zip(descs, xys)))
        d2h = int(ctx.last_records() * 16 + len(mine) * 8)
        del res

    # ---- reduce over ranks (max time, total pairs) ------------------------------------------------------------------
    stats = torch.tensor([gpu_ms / args.steps, search_ms / args.steps, (t1w - t0w) * 1e3 / args.steps, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(len(mine)), float(launches), float(records), float(errs), float(tc_pairs), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step, ms_search, ms_wall, ms_e2e = stats.tolist()
    n_pairs, launches, records, errs, tc_pairs, h2d, d2h = tot.tolist()

    if rank == 0:
        M = args.features
        flop_pair = 2.0 * M * M * 128
        peak, peak_src = peaks()
        achieved = n_pairs * flop_pair / (ms_search * 1e-3) / 1e12 / world if not hamming else None
        out = {
            "metric": metric, "value": n_pairs / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32-popcount" if hamming else "f16 (fp16 operands, fp32 accumulate; exact on integer-valued SIFT)", "data": "synthetic",
            "config": {"workload": f"{len(descs)} synthetic images x {M} {'MLDB 64-byte' if hamming else 'SIFT 128-D ' + args.dtype} features, "
                                   f"{'vocabulary-tree style (50 neighbours/image)' if args.pairs == 'voctree' else 'exhaustive'} "
                                   f"{int(n_pairs)} ordered pairs, BRUTE_FORCE_{'HAMMING' if hamming else 'L2'}, ratio 0.8",
                       "pairs_per_gpu": n_pairs / world, "sharding": "pairs dealt round-robin by database image, no collective on the data path",
                       "l2_policy": f"inputs larger than L2 ({len(descs) * M * (64 if hamming else 256) / 1e6:.0f} MB of resident descriptors vs 126 MB L2)",
                       "tensor_core_pairs": tc_pairs, "exactness_errors": errs, "wall_ms_per_step": ms_wall, "records_per_step": records / args.steps},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": (n_pairs / (ms_e2e * 1e-3)) if ms_e2e > 0 else None, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "what": "b200m_clear_views + b200m_upload_views_async (every view, from host memory) + b200m_match_pairs(STAGE_FULL): H2D overlapped with the "
                            "first pairs, kernels, D2H, host de-duplication, {(I,J): matches} map; wall clock"},
        }
        if hamming:
            hb = (n_pairs / world) * 2.0 * M * 64 / (ms_search * 1e-3) / 1e9
            hp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
            tr = ncu_traffic("hamming_top2_kernel")
            out["roofline"] = {"bound": "hbm", "achieved": hb, "peak": hp, "unit": "GB/s", "frac": hb / hp, "traffic": tr["dram_bytes_per_launch"] if tr else None,
                               "traffic_note": tr["note"] if tr else None,
                               "note": "popc-issue bound by construction (M^2*16 popc32 per pair vs 2*M*64 bytes); HBM fraction reported because the north star asks"}
        else:
            tr = ncu_traffic("l2_top2_tc2_kernel") if (args.tc_variant == 4 and M == 8192) else None
            out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                               "traffic": tr["dram_bytes_per_launch"] if tr else None, "traffic_note": tr["note"] if tr else None,
                               "kernel": {1: "tc::l2_top2_tc_kernel", 2: "tc2::l2_top2_tc2_kernel<8,false> (cta_group::2)", 3: "tc2::l2_top2_tc2_kernel<16,false> (cta_group::2)", 4: "tc2::l2_top2_tc2_kernel<8,true> (cta_group::2, K=128+16)"}[args.tc_variant], "peak_source": peak_src, "flop_per_pair": flop_pair,
                               "kernel_ms_per_step": ms_search}
        if not args.no_cpu and world == 1:      # contract: the CPU baseline is timed on rank 0 at N=1 only
            out["cpu_baseline"] = cpu_baseline(descs, xys, pairs, hamming, args.cpu_seconds)
            if not hamming:                     # the north star also names the reference's cascade-hashing matcher (scalar descriptors only)
                cb = cpu_baseline_cascade(descs, xys, pairs, max(2.0, args.cpu_seconds / 3))
                if cb:
                    out["cpu_baseline_cascade_hashing"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
