// K1 — squared-L2 nearest-neighbour search as a tcgen05 distance GEMM with a fused row-wise top-2.
//
// Replaces, for one (I = database, J = query) image pair, the reference's
//   ArrayMatcher_bruteForce::SearchNeighbours  (matching/ArrayMatcher_bruteForce.hpp:98-142)
//   + feature::L2_Vectorized                   (feature/metric.hpp:48-139)
//   + partial_sort top-2                       (stl/indexedSort.hpp:40-55)
//   + the pre-filter half of NNdistanceRatio   (matching/filters.hpp:35-67)
// on integer-valued descriptors (|v| <= 1024, ||v||^2 < 2^22), where every product, partial sum and
// distance is an exact integer (or half-integer) below 2^24 in fp32.
//
//   ||a-b||^2 = ||a||^2 + 2*h,   h = ||b||^2/2 - a.b
//
// The accumulator holds -a.b (B negated by the instruction descriptor); the epilogue adds the database
// half-norm (packed FADD2), reduces 16-column chunks with 3-input min trees (FMNMX3), and keeps the two
// smallest CHUNK MINIMA plus the chunk id of the best per query row.  That is < 1 ALU op per element
// instead of ~3 for an element-wise top-2; the price is that the second value is only an upper bound
// when best and second-best share a chunk, which the exactness pass (verify.cuh) repairs by re-scoring
// the 16 rows of the winning chunk for the few queries that pass the pre-test.
//
// CTA = 12 warps, 1 CTA/SM, persistent over work items (128 queries x whole database image):
//   warp 0  TMA producer   Q tile (128x128 fp16, once per item); DB tiles (256 rows) as two K-halves of
//                          256x64 fp16 = 32 KB, ONE TMA box each, through a 4-slot ring (a slot is refilled as soon
//                          as the four MMAs that read it retire -> 1.5 tiles of latency tolerance)
//   warp 1  MMA issuer     8 x tcgen05.mma.kind::f16 M128 N256 K16 per DB tile into one of 2 TMEM stages
//   warp 2  TMEM allocator (512 columns)
//   warp 3  half-norm producer (cp.async.bulk of 256 floats per tile, 4-slot ring)
//   warps 4-11 epilogue    warp w reads TMEM lanes 32*(w%4).., columns 128*((w-4)/4).. of the stage
// Measured pipeline behaviour that shaped this layout (tools/gpu_trace.py, profiles/): issuing one TMA costs the
// producer thread ~100-150 cycles regardless of size and the data lands ~600 cycles later, so few large boxes
// and early slot release matter more than bandwidth.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace b200m {
namespace tc {

constexpr int BM = 128;            // queries per work item  (UMMA M, TMEM lanes)
constexpr int BN = 256;            // database rows per tile (UMMA N, TMEM columns per stage)
constexpr int KD = 128;            // descriptor length
constexpr int CHUNK = 16;          // columns per chunk minimum
constexpr int NS = 4;              // database ring slots (each = one K-half of a tile: 256 rows x 64 fp16)
constexpr int NBS = 4;             // half-norm ring slots
constexpr int Q_BYTES = BM * KD * 2;        // 32 KB
constexpr int SLOT_BYTES = BN * 64 * 2;     // 32 KB
constexpr int NB_BYTES = BN * 4;
constexpr int NUM_THREADS = 384;
constexpr int EPI_THREADS = 256;

constexpr int OFF_Q = 0;
constexpr int OFF_DB = OFF_Q + 2 * Q_BYTES;
constexpr int OFF_NB = OFF_DB + NS * SLOT_BYTES;
constexpr int OFF_MRG = OFF_NB + NBS * NB_BYTES;      // 2 x 128 x float4
constexpr int OFF_BAR = OFF_MRG + 2 * BM * 16;
constexpr int NUM_BARS = 2 + 2 + NS + NS + 2 + 2 + NBS + NBS;
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM + 16;

struct Top2 { float m1, m2; uint32_t g1; };

// Fold one 16-column chunk (already h = acc + nbh) into the running top-2 of chunk minima.
__device__ __forceinline__ void fold_chunk(const float* h, uint32_t gid, Top2& s) {
  float a = ptx::fmin3(h[0], h[1], h[2]);
  float b = ptx::fmin3(h[3], h[4], h[5]);
  float c = ptx::fmin3(h[6], h[7], h[8]);
  float d = ptx::fmin3(h[9], h[10], h[11]);
  float e = ptx::fmin3(h[12], h[13], h[14]);
  float cm = fminf(ptx::fmin3(a, b, c), ptx::fmin3(d, e, h[15]));
  s.m2 = fminf(s.m2, fmaxf(s.m1, cm));
  s.g1 = (cm < s.m1) ? gid : s.g1;
  s.m1 = fminf(s.m1, cm);
}

// h[0..31] = accumulator columns + database half-norms, two columns per FADD2.
__device__ __forceinline__ void add_halfnorms(const uint32_t (&acc)[32], const float4* __restrict__ nb4, float (&h)[32]) {
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const float4 nb = nb4[v];
    ptx::add_f32x2(h[4 * v + 0], h[4 * v + 1], __uint_as_float(acc[4 * v + 0]), __uint_as_float(acc[4 * v + 1]), nb.x, nb.y);
    ptx::add_f32x2(h[4 * v + 2], h[4 * v + 3], __uint_as_float(acc[4 * v + 2]), __uint_as_float(acc[4 * v + 3]), nb.z, nb.w);
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
l2_top2_tc_kernel(const ViewDev* __restrict__ views, const PairDev* __restrict__ pairs, const WorkItem* __restrict__ items, int n_items,
                  Cand* __restrict__ cands, int* __restrict__ cand_count, float ratio_sq, long long* __restrict__ trace_buf) {
  long long* trace = (blockIdx.x == 0) ? trace_buf : nullptr;
  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B operand tiles need 1024-B alignment
  if ((ptx::smem_u32(smem) & 1023u) != 0) { asm volatile("trap;"); }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = q_full + 2;          // [2]
  uint64_t* db_full = q_empty + 2;         // [NS]
  uint64_t* db_empty = db_full + NS;       // [NS]
  uint64_t* tm_full = db_empty + NS;       // [2]
  uint64_t* tm_empty = tm_full + 2;        // [2], 8 arrivals
  uint64_t* nb_full = tm_empty + 2;        // [NBS]
  uint64_t* nb_empty = nb_full + NBS;      // [NBS], 8 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1);  ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&tm_full[i], 1); ptx::mbar_init(&tm_empty[i], EPI_THREADS / 32);
    }
    for (int i = 0; i < NS; ++i) { ptx::mbar_init(&db_full[i], 1); ptx::mbar_init(&db_empty[i], 1); }
    for (int i = 0; i < NBS; ++i) { ptx::mbar_init(&nb_full[i], 1); ptx::mbar_init(&nb_empty[i], EPI_THREADS / 32); }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: Q tiles and database K-halves
    if (lane == 0) {
      uint32_t qb = 0, qph = 0, st = 0, sph = 0, tt = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = items[it];
        const PairDev p = pairs[w.pair];
        const ViewDev* vi = views + p.view_i;
        const ViewDev* vj = views + p.view_j;
        ptx::mbar_wait(&q_empty[qb], qph ^ 1);
        ptx::mbar_arrive_expect_tx(&q_full[qb], Q_BYTES);
        uint8_t* qs = smem + OFF_Q + qb * Q_BYTES;
        ptx::tma_load_2d(qs, &vj->tmap128, &q_full[qb], 0, (int)w.qtile * BM);
        ptx::tma_load_2d(qs + BM * 128, &vj->tmap128, &q_full[qb], 64, (int)w.qtile * BM);
        qb ^= 1; if (qb == 0) qph ^= 1;
        const int ntiles = ((int)p.m_i + BN - 1) / BN;
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            ptx::mbar_wait(&db_empty[st], sph ^ 1);
            if (kb == 0) ptx::trace_stamp(trace, 0, tt, 0);
            ptx::mbar_arrive_expect_tx(&db_full[st], SLOT_BYTES);
            ptx::tma_load_2d(smem + OFF_DB + st * SLOT_BYTES, &vi->tmap256, &db_full[st], kb * 64, t * BN);
            if (++st == NS) { st = 0; sph ^= 1; }
          }
          ptx::trace_stamp(trace, 0, tt, 1); ++tt;
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ half-norm producer
    if (lane == 0) {
      uint32_t ns = 0, nph = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = items[it];
        const PairDev p = pairs[w.pair];
        const float* nbh = views[p.view_i].nbh;
        const int ntiles = ((int)p.m_i + BN - 1) / BN;
        for (int t = 0; t < ntiles; ++t) {
          ptx::mbar_wait(&nb_empty[ns], nph ^ 1);
          ptx::mbar_arrive_expect_tx(&nb_full[ns], NB_BYTES);
          ptx::bulk_load_1d(smem + OFF_NB + ns * NB_BYTES, nbh + (size_t)t * BN, NB_BYTES, &nb_full[ns]);
          if (++ns == NBS) { ns = 0; nph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_f16(BM, BN, false, true);   // D = A * (-B)^T
      const uint32_t q_addr = ptx::smem_u32(smem + OFF_Q);
      const uint32_t db_addr = ptx::smem_u32(smem + OFF_DB);
      uint32_t qb = 0, qph = 0, st = 0, sph = 0, ac = 0, aph = 0, tt = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const WorkItem w = items[it];
        const PairDev p = pairs[w.pair];
        const int ntiles = ((int)p.m_i + BN - 1) / BN;
        ptx::mbar_wait(&q_full[qb], qph);
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            ptx::mbar_wait(&db_full[st], sph);
            if (kb == 0) {
              ptx::trace_stamp(trace, 1, tt, 0);
              ptx::mbar_wait(&tm_empty[ac], aph ^ 1);
              ptx::trace_stamp(trace, 1, tt, 1);
            }
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t ad = ptx::umma_desc_k_sw128(q_addr + qb * Q_BYTES + kb * (BM * 128) + k * 32);
              const uint64_t bd = ptx::umma_desc_k_sw128(db_addr + st * SLOT_BYTES + k * 32);
              ptx::umma_f16_ss(tmem_base + ac * BN, ad, bd, idesc, (kb | k) ? 1u : 0u);
            }
            ptx::umma_commit(&db_empty[st]);
            if (++st == NS) { st = 0; sph ^= 1; }
          }
          ptx::umma_commit(&tm_full[ac]);
          ptx::trace_stamp(trace, 1, tt, 2); ++tt;
          ac ^= 1; if (ac == 0) aph ^= 1;
        }
        ptx::umma_commit(&q_empty[qb]);
        qb ^= 1; if (qb == 0) qph ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;        // which 128 columns of the 256-column stage
    const int row = quad * 32 + lane;        // query row inside the tile
    uint32_t ac = 0, aph = 0, ns = 0, nph = 0, par = 0, tt = 0;
    long long* etrace = (lane == 0 && quad == 0) ? trace : nullptr;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
      const WorkItem w = items[it];
      const PairDev p = pairs[w.pair];
      const int ntiles = ((int)p.m_i + BN - 1) / BN;
      Top2 s{INFINITY, INFINITY, 0u};
      for (int t = 0; t < ntiles; ++t) {
        ptx::mbar_wait(&nb_full[ns], nph);
        ptx::trace_stamp(etrace, 2 + half, tt, 0);
        ptx::mbar_wait(&tm_full[ac], aph);
        ptx::trace_stamp(etrace, 2 + half, tt, 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + ac * BN + half * 128;
        const float4* nb4 = reinterpret_cast<const float4*>(smem + OFF_NB + ns * NB_BYTES) + half * 32;
        const uint32_t gbase = (uint32_t)t * (BN / CHUNK) + half * (128 / CHUNK);
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32b_x32(taddr, ra);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t (&cur)[32] = (c & 1) ? rb : ra;
          uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
          if (c < 3) ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, nxt);
          float h[32];
          add_halfnorms(cur, nb4 + c * 8, h);
          fold_chunk(h, gbase + c * 2, s);
          fold_chunk(h + 16, gbase + c * 2 + 1, s);
          if (c < 3) ptx::tmem_ld_wait();
        }
        ptx::trace_stamp(etrace, 2 + half, tt, 2); ++tt;
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) { ptx::mbar_arrive(&tm_empty[ac]); ptx::mbar_arrive(&nb_empty[ns]); }
        ac ^= 1; if (ac == 0) aph ^= 1;
        if (++ns == NBS) { ns = 0; nph ^= 1; }
      }
      // merge the two column halves of each query row, pre-test, emit candidates
      float4* mrg = reinterpret_cast<float4*>(smem + OFF_MRG) + par * BM;
      if (half == 1) mrg[row] = make_float4(s.m1, s.m2, __uint_as_float(s.g1), 0.f);
      ptx::named_bar_sync(1, EPI_THREADS);
      if (half == 0) {
        const float4 o = mrg[row];
        const float m2 = fminf(fmaxf(s.m1, o.x), fminf(s.m2, o.y));
        const uint32_t g1 = (o.x < s.m1) ? __float_as_uint(o.z) : s.g1;
        const float m1 = fminf(s.m1, o.x);
        const uint32_t q = w.qtile * BM + row;
        bool keep = false;
        float d1 = 0.f, d2 = 0.f;
        if (q < p.m_j) {
          const float na = views[p.view_j].nrm[q];
          d1 = fmaf(2.f, m1, na);
          d2 = fmaf(2.f, m2, na);
          keep = d1 < __fmul_rn(ratio_sq, d2);
        }
        const uint32_t mask = __ballot_sync(0xffffffffu, keep);
        if (mask) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&cand_count[w.pair], __popc(mask));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (keep) cands[p.cand_base + base + __popc(mask & ((1u << lane) - 1))] = Cand{q, g1, d1, d2};
        }
      }
      par ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace b200m
