// K1 (2-CTA) — the same distance GEMM + fused top-2 as l2_tc.cuh, issued as cta_group::2 MMAs by CTA pairs.
//
// Why: ncu of the 1-CTA kernel (profiles/r01_ncu_l2_top2_tc_kernel.md) shows the tensor pipe 55 % busy with nothing
// else saturated: per 1024-cycle tile one SM must read 96 KB of operands from shared memory for the MMA and accept
// 64 KB of TMA writes (= 160 KB at 128 B/clk -> >= 1250 cycles), with only two 64 KB stages to hide the L2 latency.
// A CTA pair executing one M=256 x N=256 x K=16 MMA halves both: each CTA stages only ITS 128 rows of the 256-row
// database tile (32 KB per tile, 4-stage ring = 4 tiles of prefetch) and the MMA reads 64 KB per tile per SM.
//
// Work item = 256 consecutive queries of J (128 per CTA of the pair) x the whole database image I.
//   every CTA : warp 0 TMA producer (own Q tile, own half of each DB tile, full 256-entry half-norm chunk),
//               warp 2 TMEM allocator (cta_group::2), then, with warp 3, the EXACTNESS PASS of the previous work item:
//               the epilogue queues its candidates in shared memory and these two otherwise idle warps re-score them
//               (verify.cuh rescore_candidate) and append the final records while the next item streams through the
//               tensor pipe, so the match list is complete when the kernel ends: no second pass over the pairs;
//               warps 4.. epilogue on its own 128 accumulator rows
//   leader CTA: warp 1 issues tcgen05.mma.cta_group::2; completions are multicast to both CTAs' barriers
// Barrier homes: q_full / db_full / tm_empty live in the LEADER (TMA bytes of both CTAs and the epilogue arrivals of
// both CTAs are credited there); q_empty / db_empty / tm_full / nb_full / nb_empty are per CTA.
//
// MODE (template): what the epilogue keeps per query row and what leaves the kernel
//   MODE_MATCH  integer-valued descriptors (exact accumulators): two smallest chunk minima + chunk of the best, ratio pre-test,
//               candidates re-scored exactly (in-kernel when `fused`) -> final match records.                       [the hot path]
//   MODE_KNN    the same data, ArrayMatcher surface: chunk of the second best is kept too and EVERY query leaves a candidate
//               (dense, slot = query row) for knn_finalize_kernel, which needs both neighbours' indices.
//   MODE_REAL   real-valued fp32 descriptors: the GEMM on the fp16-ROUNDED rows is only a FILTER.  The query's half-norm is folded in
//               as well (A tile = the query's own limbs, loaded per item), so the accumulator is ||a~ - b~||^2 / 2 >= 0 and the chunk id
//               rides in the low mantissa bits of the chunk minimum: the FOUR smallest packed minima are kept with 7 integer min/max
//               per chunk.  |sqrt(d~) - sqrt(d)| <= ||a - a~|| + ||b - b~|| bounds what the rounding can do, so the rows outside the
//               best chunks are provably out of the top-2 whenever the next chunk minimum is far enough; the exact distances (the
//               reference's fp32 SSE order, from the original fp32 rows) come from re-scoring those chunks (verify.cuh rescore_real),
//               and the rare query the bound cannot decide goes to the exact_rows fallback.  Results are bit-identical to the
//               reference whatever the rounding did.
#pragma once
#include "l2_tc.cuh"
#include "verify.cuh"

namespace b200m {
namespace tc2 {

constexpr int BM = 128;                 // queries per CTA (256 per pair)
constexpr int BN = 256;                 // database rows per tile (128 staged per CTA)
constexpr int Q_BYTES = BM * 128 * 2;   // 32 KB
constexpr int DBH_BYTES = 128 * 128 * 2;  // 32 KB: this CTA's half of a database tile (two K-blocks of 64)
constexpr int AUG_BYTES = 128 * 16 * 2;   // 4 KB: 16 augmentation columns of the same 128 rows (AUG only)
constexpr int NB_BYTES = BN * 4;
constexpr int MAX_EPI_WARPS = 16;
enum : int { MODE_MATCH = 0, MODE_KNN = 1, MODE_REAL = 2 };

// Shared-memory layout. AUG = the database half-norm is folded into the GEMM as a 9th K-step: each database row carries
// 16 extra fp16 columns [b0, l0, l1, 0...] with ||b||^2/2 = 0.5*b0 + l0 + 2048*l1 (all three exact in fp16), every query
// row the constants [-0.5, -1, -2048, 0...]; with B negated by the instruction descriptor the accumulator becomes
// h = ||b||^2/2 - a.b directly.  That removes the half-norm ring, 32 LDS.128 and 64 FADD2 per thread and tile from
// the epilogue for 12.5 % more tensor work.
template <bool AUG, bool REAL = false> struct Lay {
  // database smem stages per CTA.  r02: a 4th stage for the AUG layout (one more tile of TMA lead: 3717 instead of 2347 cycles between a box
  // being issued and its use) changed neither the per-tile period nor the bench value - the feed is not the limit - so the AUG kernels keep
  // 3 stages and leave the shared memory to L1 (the re-scoring warps' row loads) and to the upload-stream kernels that co-run.
  static constexpr int NS = AUG ? 3 : 4;
  static constexpr int MRG_GROUPS = AUG ? 1 : 3;                     // column groups merged through shared memory (8 epilogue warps: 1; the 16-warp variant: 3)
  static constexpr int STAGE = DBH_BYTES + (AUG ? AUG_BYTES : 0);
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_DB = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_NB = OFF_DB + NS * STAGE;                 // non-AUG: half-norm ring; AUG: the constant A tile (4 KB); REAL: the query's own limb tile, 2 buffers
  static constexpr int OFF_MRG = OFF_NB + (AUG ? (REAL ? 2 : 1) * AUG_BYTES : NS * NB_BYTES);
  static constexpr int OFF_VQ = OFF_MRG + 2 * MRG_GROUPS * BM * 16;  // candidate queue: 2 buffers x 128 Cand (32 slots per epilogue quadrant)
  static constexpr int OFF_VQX = OFF_VQ + 2 * BM * 16;               // REAL: 5th word of the queued candidates (p4), 2 x 128 x 4 B
  static constexpr int OFF_VQN = OFF_VQX + (REAL ? 2 * BM * 4 : 0);  // 2 x (4 per-quadrant counts + pair index), 32 B each
  static constexpr int OFF_BAR = OFF_VQN + 2 * 32;
  static constexpr int NUM_BARS = 2 + 2 + NS + NS + 2 + 2 + NS + NS + 2 + 2;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_BYTES = OFF_TMEM + 16;
};
constexpr int SMEM_BYTES = Lay<false>::SMEM_BYTES;

// Packed four-smallest list of the real-valued filter (non-negative float bit patterns order like unsigned integers).
struct Top4 { uint32_t p1, p2, p3, p4; };
__device__ __forceinline__ void top4_insert(Top4& s, uint32_t pk) {
  s.p4 = min(s.p4, max(s.p3, pk));
  s.p3 = min(s.p3, max(s.p2, pk));
  s.p2 = min(s.p2, max(s.p1, pk));
  s.p1 = min(s.p1, pk);
}
__device__ __forceinline__ void fold_chunk_real(const float* h, uint32_t gid, Top4& s) {
  float a = ptx::fmin3(h[0], h[1], h[2]);
  float b = ptx::fmin3(h[3], h[4], h[5]);
  float c = ptx::fmin3(h[6], h[7], h[8]);
  float d = ptx::fmin3(h[9], h[10], h[11]);
  float e = ptx::fmin3(h[12], h[13], h[14]);
  const float cm = fmaxf(fminf(ptx::fmin3(a, b, c), ptx::fmin3(d, e, h[15])), 0.f);   // d~/2 >= 0 up to accumulation rounding
  top4_insert(s, (__float_as_uint(cm) & ~REAL_IDMASK) | gid);
}
// MODE_KNN: like tc::fold_chunk, plus the chunk of the second smallest minimum
struct Top2g { float m1, m2; uint32_t g1, g2; };
__device__ __forceinline__ void fold_chunk_knn(const float* h, uint32_t gid, Top2g& s) {
  float a = ptx::fmin3(h[0], h[1], h[2]);
  float b = ptx::fmin3(h[3], h[4], h[5]);
  float c = ptx::fmin3(h[6], h[7], h[8]);
  float d = ptx::fmin3(h[9], h[10], h[11]);
  float e = ptx::fmin3(h[12], h[13], h[14]);
  const float cm = fminf(ptx::fmin3(a, b, c), ptx::fmin3(d, e, h[15]));
  const bool lt1 = cm < s.m1, lt2 = cm < s.m2;
  s.g2 = lt1 ? s.g1 : (lt2 ? gid : s.g2);
  s.m2 = fminf(s.m2, fmaxf(s.m1, cm));
  s.g1 = lt1 ? gid : s.g1;
  s.m1 = fminf(s.m1, cm);
}

// The real-valued re-scoring reads 32 fp32 rows per candidate (16 KB, mostly from HBM) and is latency-bound: two warps cannot keep up with
// the tensor pipe (measured: 135 ms per step instead of 60), so MODE_REAL adds six re-scoring warps behind the epilogue warps.
constexpr int REAL_EXTRA_VERIFY_WARPS = 6;
template <int EPI_WARPS, int MODE> constexpr int block_threads() { return 128 + EPI_WARPS * 32 + (MODE == MODE_REAL ? REAL_EXTRA_VERIFY_WARPS * 32 : 0); }

template <int EPI_WARPS, bool AUG, int MODE = MODE_MATCH>   // 8 or 16 epilogue warps per CTA (128 or 64 accumulator columns per warp); AUG: see Lay
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(block_threads<EPI_WARPS, MODE>(), 1)
l2_top2_tc2_kernel(const ViewDev* __restrict__ views, const PairDev* __restrict__ pairs, const WorkItem* __restrict__ items, int n_items,
                   Cand* __restrict__ cands, int* __restrict__ cand_count, float ratio_sq, long long* __restrict__ trace_buf, int dbg, unsigned int* __restrict__ err_count, int fused,
                   uint32_t* __restrict__ candx, FbSink fb) {
  long long* trace = (blockIdx.x == 0) ? trace_buf : nullptr;   // dbg (ablation, debug only): 1 = skip epilogue math, 2 = also skip TMEM loads
  constexpr bool REAL = MODE == MODE_REAL, KNN = MODE == MODE_KNN;
  constexpr int VERIFY_WARPS_K = REAL ? 2 + REAL_EXTRA_VERIFY_WARPS : 2;   // warps 2, 3 (+ the warps behind the epilogue warps)
  static_assert(!(REAL || KNN) || (AUG && EPI_WARPS == 8), "MODE_REAL / MODE_KNN exist for the default variant only");
  using L = Lay<AUG, REAL>;
  constexpr int NS = L::NS, OFF_Q = L::OFF_Q, OFF_DB = L::OFF_DB, OFF_NB = L::OFF_NB, OFF_MRG = L::OFF_MRG, OFF_VQ = L::OFF_VQ, OFF_VQX = L::OFF_VQX,
                OFF_VQN = L::OFF_VQN, OFF_BAR = L::OFF_BAR, OFF_TMEM = L::OFF_TMEM, STAGE = L::STAGE;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) { asm volatile("trap;"); }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                  // [2]  leader
  uint64_t* q_empty = q_full + 2;           // [2]  per CTA (multicast commit)
  uint64_t* db_full = q_empty + 2;          // [NS] leader
  uint64_t* db_empty = db_full + NS;        // [NS] per CTA (multicast commit)
  uint64_t* tm_full = db_empty + NS;        // [2]  per CTA (multicast commit)
  uint64_t* tm_empty = tm_full + 2;         // [2]  leader, 16 arrivals (8 epilogue warps x 2 CTAs)
  uint64_t* nb_full = tm_empty + 2;         // [NS] per CTA (half-norm ring, same index as the database stage)
  uint64_t* nb_empty = nb_full + NS;        // [NS] per CTA, 8 arrivals
  uint64_t* vq_full = nb_empty + NS;        // [2]  per CTA, 4 arrivals (the epilogue quadrants that emit candidates)
  uint64_t* vq_empty = vq_full + 2;         // [2]  per CTA, 2 arrivals (the two exactness-pass warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();      // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1);  ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&tm_full[i], 1); ptx::mbar_init(&tm_empty[i], 2 * EPI_WARPS);
    }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&vq_full[i], 4); ptx::mbar_init(&vq_empty[i], VERIFY_WARPS_K); }
    for (int i = 0; i < NS; ++i) {
      ptx::mbar_init(&db_full[i], 1); ptx::mbar_init(&db_empty[i], 1);
      ptx::mbar_init(&nb_full[i], 1); ptx::mbar_init(&nb_empty[i], EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_2sm(tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  if (AUG && !REAL && warp >= 4 && warp < 8) {
    // the constant augmentation tile of the A operand: 128 rows x 16 fp16 in the 32-byte-swizzled K-major layout
    // (16-byte chunk index XOR ((row >> 2) & 1)); logical chunk 0 = [-0.5, -1, -2048, 0, 0, 0, 0, 0], chunk 1 = zeros
    const int r = (warp - 4) * 32 + lane;
    uint4* rowp = reinterpret_cast<uint4*>(smem + OFF_NB + r * 32);
    const uint32_t c01 = (uint32_t)__half_as_ushort(__float2half_rn(-0.5f)) | ((uint32_t)__half_as_ushort(__float2half_rn(-1.f)) << 16);
    const uint32_t c23 = (uint32_t)__half_as_ushort(__float2half_rn(-2048.f));
    const int sw = (r >> 2) & 1;
    rowp[sw] = make_uint4(c01, c23, 0u, 0u);
    rowp[sw ^ 1] = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
  }
  __syncwarp();                      // barrier.cluster is .aligned: every warp must reach it converged
  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // barriers of BOTH CTAs are initialised before anyone signals across the pair
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t qb = 0, qph = 0, st = 0, sph = 0, tt = 0;
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        const WorkItem w = items[it];
        const PairDev p = pairs[w.pair];
        const ViewDev* vi = views + p.view_i;
        const ViewDev* vj = views + p.view_j;
        const int qrow = ((int)w.qtile * 2 + (int)rank) * BM;
        ptx::mbar_wait(&q_empty[qb], qph ^ 1);
        if (rank == 0) ptx::mbar_arrive_expect_tx(&q_full[qb], 2 * (Q_BYTES + (REAL ? AUG_BYTES : 0)));
        uint8_t* qs = smem + OFF_Q + qb * Q_BYTES;
        ptx::tma_load_2d_2sm(qs, &vj->tmap128, &q_full[qb], 0, qrow);
        ptx::tma_load_2d_2sm(qs + BM * 128, &vj->tmap128, &q_full[qb], 64, qrow);
        if (REAL) ptx::tma_load_2d_2sm(smem + OFF_NB + qb * AUG_BYTES, &vj->tmap_augq, &q_full[qb], 0, qrow);   // the query rows' own half-norm limbs
        qb ^= 1; if (qb == 0) qph ^= 1;
        const int ntiles = ((int)p.m_i + BN - 1) / BN;
        for (int t = 0; t < ntiles; ++t) {
          ptx::mbar_wait(&db_empty[st], sph ^ 1);
          if (!AUG) ptx::mbar_wait(&nb_empty[st], sph ^ 1);         // released by the epilogue NS tiles ago: never on the critical path
          ptx::trace_stamp(trace, 0, tt, 0);
          uint8_t* ds = smem + OFF_DB + st * STAGE;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&db_full[st], 2 * STAGE);
          const int drow = t * BN + (int)rank * 128;
          ptx::tma_load_2d_2sm(ds, &vi->tmap128, &db_full[st], 0, drow);
          ptx::tma_load_2d_2sm(ds + 128 * 128, &vi->tmap128, &db_full[st], 64, drow);
          if (AUG) {
            ptx::tma_load_2d_2sm(ds + DBH_BYTES, &vi->tmap_aug, &db_full[st], 0, drow);
          } else {
            ptx::mbar_arrive_expect_tx(&nb_full[st], NB_BYTES);
            ptx::bulk_load_1d(smem + OFF_NB + st * NB_BYTES, vi->nbh + (size_t)t * BN, NB_BYTES, &nb_full[st]);
          }
          ptx::trace_stamp(trace, 0, tt, 1); ++tt;
          if (++st == NS) { st = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    // The WHOLE warp runs this loop converged and only the tcgen05.mma / tcgen05.commit instructions are predicated on one elected lane.
    // r02 (profiles/r02_k1_issue_bound.md): with `if (lane == 0)` around the loop the compiler kept the shared-memory descriptors in vector
    // registers of a divergent thread - ~17 dependent ALU / R2UR instructions in front of every UTCHMMA, ~1040 cycles of issue per
    // 1152-cycle tile: the kernel was bound by this thread, not by the tensor pipe.  Warp-uniform control flow lets the descriptor
    // arithmetic live in the uniform datapath.
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(2 * BM, BN, false, true);   // M=256 across the pair, D = A * (-B)^T
      const uint32_t q_addr = ptx::smem_u32(smem + OFF_Q);
      const uint32_t db_addr = ptx::smem_u32(smem + OFF_DB);
      const bool issuer = ptx::elect_one();
      long long* mtrace = issuer ? trace : nullptr;
      uint32_t qb = 0, qph = 0, st = 0, sph = 0, ac = 0, aph = 0, tt = 0;
      for (int it = cluster_id; it < n_items; it += n_clusters) {
        const WorkItem w = items[it];
        const PairDev p = pairs[w.pair];
        const int ntiles = ((int)p.m_i + BN - 1) / BN;
        ptx::mbar_wait(&q_full[qb], qph);
        bool db_ready = false, tm_ready = false;   // barriers of the coming tile already observed (warp-uniform)
        for (int t = 0; t < ntiles; ++t) {
          if (mtrace != nullptr) {     // debug only (CTA 0 with tracing on): did the look-ahead polls succeed; loop-top time
            if (tt < ptx::TRACE_TILES) mtrace[((size_t)0 * ptx::TRACE_TILES + tt) * 4 + 2] = (db_ready ? 1 : 0) | (tm_ready ? 2 : 0);
            ptx::trace_stamp(mtrace, 1, tt, 0);
          }
          if (!db_ready) ptx::mbar_wait(&db_full[st], sph);
          if (!tm_ready) ptx::mbar_wait(&tm_empty[ac], aph ^ 1);
          ptx::trace_stamp(mtrace, 1, tt, 1);
          ptx::tc_fence_after();
          const uint32_t a_base = q_addr + qb * Q_BYTES, b_base = db_addr + st * STAGE, d_addr = tmem_base + ac * BN;
          // The NEXT tile's barriers are polled while this tile's MMAs are queued (the database slot after the 4th MMA, the TMEM stage
          // before the last two); a failed poll falls back to a blocking wait at the top of the next tile.
          const uint32_t nst = (st + 1 == NS) ? 0 : st + 1, nsph = (st + 1 == NS) ? (sph ^ 1) : sph;
          const uint32_t nac = ac ^ 1, naph = (nac == 0) ? (aph ^ 1) : aph;
          const bool more = (t + 1 < ntiles);
          db_ready = tm_ready = false;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t da = ptx::umma_desc_k_sw128(a_base + (k >> 2) * (BM * 128) + (k & 3) * 32);
            const uint64_t db = ptx::umma_desc_k_sw128(b_base + (k >> 2) * (128 * 128) + (k & 3) * 32);
            if (issuer) ptx::umma_f16_ss_2sm(d_addr, da, db, idesc, k > 0 ? 1u : 0u);
            if (k == 3 && more) db_ready = __all_sync(0xffffffffu, ptx::mbar_try_wait(&db_full[nst], nsph));
            if (k == (AUG ? 6 : 5) && more) tm_ready = __all_sync(0xffffffffu, ptx::mbar_try_wait(&tm_empty[nac], naph ^ 1));
          }
          if (AUG) {  // 9th K-step: constants x half-norm limbs
            const uint64_t da = ptx::umma_desc_k_sw32(ptx::smem_u32(smem + OFF_NB) + (REAL ? qb * AUG_BYTES : 0));
            const uint64_t db = ptx::umma_desc_k_sw32(b_base + DBH_BYTES);
            if (issuer) ptx::umma_f16_ss_2sm(d_addr, da, db, idesc, 1u);
          }
          if (issuer) {
            ptx::umma_commit_2sm_mc(&db_empty[st], 3);
            ptx::umma_commit_2sm_mc(&tm_full[ac], 3);
          }
          __syncwarp();
          ptx::trace_stamp(mtrace, 1, tt, 2); ++tt;
          st = nst; sph = nsph; ac = nac; aph = naph;
        }
        if (issuer) ptx::umma_commit_2sm_mc(&q_empty[qb], 3);
        __syncwarp();
        qb ^= 1; if (qb == 0) qph ^= 1;
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 4 + EPI_WARPS) {
    // ------------------------------------------------------------------ exactness pass of the previous item (2 warps)
    // fused == 0 (short database images: an item lasts only a few tiles, two warps cannot hide the re-scoring latency):
    // the epilogue writes its candidates to global memory and the stand-alone exactness kernel handles them.
    uint32_t par = 0, vph = 0;
    for (int it = cluster_id; fused && it < n_items; it += n_clusters) {
      ptx::mbar_wait(&vq_full[par], vph);
      const uint32_t* hdr = reinterpret_cast<const uint32_t*>(smem + OFF_VQN + par * 32);
      const PairDev p = pairs[hdr[4]];
      const __half* db16 = views[p.view_i].h16;
      const __half* q16 = views[p.view_j].h16;
      const Cand* queue = reinterpret_cast<const Cand*>(smem + OFF_VQ) + par * BM;
      const uint32_t* queuex = reinterpret_cast<const uint32_t*>(smem + OFF_VQX) + par * BM;
      // MATCH: warp 2 serves quadrants 0-1, warp 3 quadrants 2-3.  REAL (8 warps): two warps per quadrant, alternate entries.
      const int vw = warp < 4 ? warp - 2 : warp - (4 + EPI_WARPS) + 2;
      const int quad_lo = REAL ? (vw & 3) : vw * 2, quad_hi = REAL ? quad_lo + 1 : quad_lo + 2;
      const int e_first = REAL ? (vw >> 2) : 0, e_step = REAL ? 2 : 1;
      for (int quad = quad_lo; quad < quad_hi; ++quad) {
        const int n = (int)hdr[quad];
        for (int e = e_first; e < n; e += e_step) {
          const Cand k = queue[quad * 32 + e];
          Rec rec;
          int verdict;                                                          // 0 drop, 1 keep, 2 undecided -> exact_rows fallback
          if (REAL) verdict = rescore_real(views[p.view_i], views[p.view_j], p.m_i, k, queuex[quad * 32 + e], ratio_sq, lane, err_count, rec);
          else verdict = rescore_candidate(db16, q16, p.m_i, k, ratio_sq, lane, err_count, rec) ? 1 : 0;
          if (verdict == 1 && lane == 0) {
            const int slot = atomicAdd(&cand_count[hdr[4]], 1);
            cands[p.cand_base + slot] = Cand{rec.j, rec.i, rec.d1, rec.d2};     // final record, (query, database row) order like the exact kernels
          }
          if (REAL && verdict == 2 && lane == 0) fb_push(fb, hdr[4], k.q, 0u, 0u, err_count);   // no slot of its own: the fallback appends
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&vq_empty[par]);
      par ^= 1; if (par == 0) vph ^= 1;
    }
  } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue (EPI_WARPS warps per CTA, own 128 rows)
    constexpr int NQ = EPI_WARPS / 4;            // column groups per stage (2 or 4)
    constexpr int COLS = BN / NQ;                // accumulator columns per warp (128 or 64)
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
    const int colq = (warp - 4) >> 2;            // which column group
    const int row = quad * 32 + lane;
    uint32_t ac = 0, aph = 0, st = 0, sph = 0, par = 0, vqph = 0, tt = 0;
    long long* etrace = (lane == 0 && quad == 0 && colq < 2) ? trace : nullptr;
    for (int it = cluster_id; it < n_items; it += n_clusters) {
      const WorkItem w = items[it];
      const PairDev p = pairs[w.pair];
      const int ntiles = ((int)p.m_i + BN - 1) / BN;
      tc::Top2 s{INFINITY, INFINITY, 0u};
      Top2g sk{INFINITY, INFINITY, 0xFFFFu, 0xFFFFu};                          // MODE_KNN
      Top4 s4{0x7F800000u | REAL_IDMASK, 0x7F800000u | REAL_IDMASK, 0x7F800000u | REAL_IDMASK, 0x7F800000u | REAL_IDMASK};   // MODE_REAL: +inf, id = none
      for (int t = 0; t < ntiles; ++t) {
        if (!AUG && t == 0) ptx::mbar_wait(&nb_full[st], sph);     // later tiles: already observed at the end of the previous tile
        ptx::trace_stamp(etrace, 2 + colq, tt, 0);
        ptx::mbar_wait(&tm_full[ac], aph);
        ptx::trace_stamp(etrace, 2 + colq, tt, 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ac * BN + colq * COLS;
        const float4* nb4 = reinterpret_cast<const float4*>(smem + OFF_NB + st * NB_BYTES) + colq * (COLS / 4);
        const uint32_t gbase = (uint32_t)t * (BN / tc::CHUNK) + colq * (COLS / tc::CHUNK);
        uint32_t ra[32], rb[32];
        if (dbg < 2) {
        ptx::tmem_ld_32x32b_x32(taddr, ra);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          uint32_t (&cur)[32] = (c & 1) ? rb : ra;
          uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
          if (c + 1 < COLS / 32) ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, nxt);
          else {
            // every TMEM load of this stage has completed (wait::ld at the end of the previous round): hand the stage back to the MMA
            // issuer BEFORE folding the last 32 columns - the serial chain MMA(t) -> epilogue(t) -> MMA(t+2) is what bounds the kernel
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(&tm_empty[ac], 0);     // leader: this CTA's rows of TMEM stage ac are drained
          }
          if (dbg == 0) {
            float h[32];
            if (AUG) {
#pragma unroll
              for (int e = 0; e < 32; ++e) h[e] = __uint_as_float(cur[e]);     // the accumulator already is h
            } else {
              tc::add_halfnorms(cur, nb4 + c * 8, h);
            }
            if (REAL) { fold_chunk_real(h, gbase + c * 2, s4); fold_chunk_real(h + 16, gbase + c * 2 + 1, s4); }
            else if (KNN) { fold_chunk_knn(h, gbase + c * 2, sk); fold_chunk_knn(h + 16, gbase + c * 2 + 1, sk); }
            else { tc::fold_chunk(h, gbase + c * 2, s); tc::fold_chunk(h + 16, gbase + c * 2 + 1, s); }
          } else {
            float keepalive = __uint_as_float(cur[0]);
#pragma unroll
            for (int e = 1; e < 32; ++e) keepalive = fminf(keepalive, __uint_as_float(cur[e]));   // 31 ops: keeps the loads alive
            s.m1 = fminf(s.m1, keepalive);
          }
          if (c + 1 < COLS / 32) ptx::tmem_ld_wait();
        }
        }
        ptx::trace_stamp(etrace, 2 + colq, tt, 2); ++tt;
        if (dbg >= 2) {                                     // ablation mode without TMEM loads: nothing released the stage above
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(&tm_empty[ac], 0);
        }
        __syncwarp();
        if (!AUG && lane == 0) ptx::mbar_arrive(&nb_empty[st]);      // local: the half-norm buffer may be overwritten
        ac ^= 1; if (ac == 0) aph ^= 1;
        if (++st == NS) { st = 0; sph ^= 1; }
        // the next tile's half-norms landed long ago; observing their barrier now (while the tensor pipe is still busy with
        // that tile) takes ~100 cycles off the serial MMA -> epilogue -> MMA chain
        if (!AUG && t + 1 < ntiles) ptx::mbar_wait(&nb_full[st], sph);
      }
      // merge the column groups of each query row, pre-test, emit candidates
      float4* mrg = reinterpret_cast<float4*>(smem + OFF_MRG) + par * L::MRG_GROUPS * BM;
      if (colq > 0) {
        if (REAL) mrg[(colq - 1) * BM + row] = make_float4(__uint_as_float(s4.p1), __uint_as_float(s4.p2), __uint_as_float(s4.p3), __uint_as_float(s4.p4));
        else if (KNN) mrg[(colq - 1) * BM + row] = make_float4(sk.m1, sk.m2, __uint_as_float(sk.g1), __uint_as_float(sk.g2));
        else mrg[(colq - 1) * BM + row] = make_float4(s.m1, s.m2, __uint_as_float(s.g1), 0.f);
      }
      ptx::named_bar_sync(1, EPI_WARPS * 32);
      if (colq == 0) {
        const uint32_t q = (w.qtile * 2 + rank) * BM + row;
        if (REAL) {
#pragma unroll
          for (int o_ = 0; o_ < NQ - 1; ++o_) {
            const float4 o = mrg[o_ * BM + row];
            top4_insert(s4, __float_as_uint(o.x)); top4_insert(s4, __float_as_uint(o.y)); top4_insert(s4, __float_as_uint(o.z)); top4_insert(s4, __float_as_uint(o.w));
          }
          bool keep = false;
          if (q < p.m_j) {
            // superset pre-test on bounds of the TRUE (reference) distances: lower(best) < r^2 * upper(second smallest chunk minimum)
            const RealBound rb = real_bound(views[p.view_i], views[p.view_j], q);
            const float d1t = 2.f * __uint_as_float(s4.p1 & ~REAL_IDMASK), d2t = 2.f * __uint_as_float(s4.p2 & ~REAL_IDMASK);
            keep = rb.lower(d1t) < __fmul_rn(ratio_sq, rb.upper(d2t));
          }
          const uint32_t mask = __ballot_sync(0xffffffffu, keep);
          const int pos = __popc(mask & ((1u << lane) - 1));
          if (fused) {
            ptx::mbar_wait(&vq_empty[par], vqph ^ 1);
            Cand* queue = reinterpret_cast<Cand*>(smem + OFF_VQ) + par * BM + quad * 32;
            uint32_t* queuex = reinterpret_cast<uint32_t*>(smem + OFF_VQX) + par * BM + quad * 32;
            if (keep) { queue[pos] = Cand{q, s4.p1, __uint_as_float(s4.p2), __uint_as_float(s4.p3)}; queuex[pos] = s4.p4; }
            uint32_t* hdr = reinterpret_cast<uint32_t*>(smem + OFF_VQN + par * 32);
            if (lane == 0) { hdr[quad] = (uint32_t)__popc(mask); if (quad == 0) hdr[4] = w.pair; }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&vq_full[par]);
          } else if (mask) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&cand_count[w.pair], __popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) { cands[p.cand_base + base + pos] = Cand{q, s4.p1, __uint_as_float(s4.p2), __uint_as_float(s4.p3)}; candx[p.cand_base + base + pos] = s4.p4; }
          }
        } else if (KNN) {
#pragma unroll
          for (int o_ = 0; o_ < NQ - 1; ++o_) {
            const float4 o = mrg[o_ * BM + row];
            const uint32_t og1 = __float_as_uint(o.z), og2 = __float_as_uint(o.w);
            if (o.x < sk.m1) {
              if (sk.m1 <= o.y) { sk.m2 = sk.m1; sk.g2 = sk.g1; } else { sk.m2 = o.y; sk.g2 = og2; }
              sk.m1 = o.x; sk.g1 = og1;
            } else if (o.x < sk.m2) { sk.m2 = o.x; sk.g2 = og1; }
          }
          if (q < p.m_j) {                                   // dense: every query leaves its two chunks for knn_finalize_kernel
            const float na = views[p.view_j].nrm[q];
            cands[p.cand_base + q] = Cand{q, (sk.g1 & 0xFFFFu) | (sk.g2 << 16), fmaf(2.f, sk.m1, na), fmaf(2.f, sk.m2, na)};
          }
        } else {
        float m1 = s.m1, m2 = s.m2; uint32_t g1 = s.g1;
#pragma unroll
        for (int o_ = 0; o_ < NQ - 1; ++o_) {
          const float4 o = mrg[o_ * BM + row];
          m2 = fminf(fmaxf(m1, o.x), fminf(m2, o.y));
          g1 = (o.x < m1) ? __float_as_uint(o.z) : g1;
          m1 = fminf(m1, o.x);
        }
        bool keep = false;
        float d1 = 0.f, d2 = 0.f;
        if (q < p.m_j) {
          const float na = views[p.view_j].nrm[q];
          d1 = fmaf(2.f, m1, na);
          d2 = fmaf(2.f, m2, na);
          keep = d1 < __fmul_rn(ratio_sq, d2);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (fused) {
          // hand the survivors of the pre-test to the exactness-pass warps through shared memory (32 slots per quadrant)
          ptx::mbar_wait(&vq_empty[par], vqph ^ 1);        // they finished the item that used this buffer two items ago
          Cand* queue = reinterpret_cast<Cand*>(smem + OFF_VQ) + par * BM + quad * 32;
          if (keep) queue[__popc(mask & ((1u << lane) - 1))] = Cand{q, g1, d1, d2};
          uint32_t* hdr = reinterpret_cast<uint32_t*>(smem + OFF_VQN + par * 32);
          if (lane == 0) { hdr[quad] = (uint32_t)__popc(mask); if (quad == 0) hdr[4] = w.pair; }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&vq_full[par]);
        } else if (mask) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&cand_count[w.pair], __popc(mask));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (keep) cands[p.cand_base + base + __popc(mask & ((1u << lane) - 1))] = Cand{q, g1, d1, d2};
        }
        }
      }
      par ^= 1; if (par == 0) vqph ^= 1;
    }
  }

  __syncwarp();                      // lanes 1-31 of the single-lane role warps wait here for lane 0
  ptx::tc_fence_before();
  ptx::cluster_sync_all();          // the peer may still be signalling / reading this CTA's shared memory until here
  if (warp == 2) ptx::tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace tc2
}  // namespace b200m
