// ORACLE (port) — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under
// alicevision_b200/ may include, link or call this file.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load the library built from it.
//
// A from-scratch CPU restatement of the reference algorithm for the descriptor-matching hot path,
// in plain C-style C++ (std::partial_sort / std::set are used on purpose: the reference's tie order
// and its non-strict-weak-order de-duplication are *defined* by libstdc++'s behaviour).
// Every function cites the reference lines it follows.  Parity is PINNED: tests/test_oracle.py
// checks this port (a) against the reference's own known-answer tests and (b) against
// oracle/_ref/libref_oracle.so, which is the reference's headers compiled verbatim.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <vector>
#include <omp.h>

namespace {

enum { DT_F32 = 0, DT_U8 = 1, DT_BIN = 2 };

// feature/metric.hpp:27-44  L2_Simple: sequential sum of squared differences, float accumulator
// (numeric/Accumulator.hpp:13-47: uchar/short/int/float all accumulate in float).
template <class T> float l2_simple(const T* a, const T* b, int n) {
  float r = 0.f;
  for (int i = 0; i < n; ++i) { float d = (float)(a[i] - b[i]); r += d * d; }
  return r;
}
// feature/metric.hpp:48-80  L2_Vectorized (generic): per group of 4, r += d0^2+d1^2+d2^2+d3^2; tail one by one.
// For uchar the differences are formed in int (integer promotion) and converted to float on assignment.
float l2_vec_u8(const uint8_t* a, const uint8_t* b, int n) {
  float r = 0.f; int i = 0;
  for (; i + 3 < n; i += 4) {
    float d0 = (float)((int)a[i] - (int)b[i]), d1 = (float)((int)a[i + 1] - (int)b[i + 1]);
    float d2 = (float)((int)a[i + 2] - (int)b[i + 2]), d3 = (float)((int)a[i + 3] - (int)b[i + 3]);
    r += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; i < n; ++i) { float d = (float)((int)a[i] - (int)b[i]); r += d * d; }
  return r;
}
// feature/metric.hpp:94-123,128-139  L2_Vectorized<float> = l2_sse: four independent lanes, each
// cum_l += (a-b)*(a-b) with separate mul and add (build with -ffp-contract=off), result
// ((f0+f1)+f2)+f3; size % 4 != 0 -> warning + 0 (:118-122).
float l2_vec_f32(const float* a, const float* b, int n) {
  if (n % 4 != 0) return 0.f;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int i = 0; i < n; i += 4) {
    float t0 = a[i] - b[i], t1 = a[i + 1] - b[i + 1], t2 = a[i + 2] - b[i + 2], t3 = a[i + 3] - b[i + 3];
    t0 = t0 * t0; t1 = t1 * t1; t2 = t2 * t2; t3 = t3 * t3;
    s0 = s0 + t0; s1 = s1 + t1; s2 = s2 + t2; s3 = s3 + t3;
  }
  return ((s0 + s1) + s2) + s3;
}
// feature/Hamming.hpp:113-129 (64-bit words when size%8==0), :130-139 (32-bit), :141-148 (byte LUT)
uint32_t hamming_u8(const uint8_t* a, const uint8_t* b, int n) {
  uint32_t r = 0;
  if (n % 8 == 0) {
    for (int i = 0; i < n; i += 8) { uint64_t x, y; std::memcpy(&x, a + i, 8); std::memcpy(&y, b + i, 8); r += (uint32_t)__builtin_popcountll(x ^ y); }
  } else if (n % 4 == 0) {
    for (int i = 0; i < n; i += 4) { uint32_t x, y; std::memcpy(&x, a + i, 4); std::memcpy(&y, b + i, 4); r += (uint32_t)__builtin_popcount(x ^ y); }
  } else {
    for (int i = 0; i < n; ++i) r += (uint32_t)__builtin_popcount((unsigned)(a[i] ^ b[i]));
  }
  return r;
}

template <class D> struct Packet { D val; int index; };                       // stl/indexedSort.hpp:13-18
template <class D> bool operator<(const Packet<D>& A, const Packet<D>& B) { return A.val < B.val; }  // :27-31 (value only)

// matching/ArrayMatcher_bruteForce.hpp:98-142  SearchNeighbours
template <class S, class D, class Metric>
int knn(const S* db, int n_db, const S* q, int n_q, int dim, int nn, int32_t* iq, int32_t* idb, D* dist, Metric metric) {
  if (n_db < 1) return 0;                       // Build false (:44-48) -> SearchNeighbours false (:100-103)
  if (nn > n_db || n_q < 1) return 0;           // :105-108
#pragma omp parallel for schedule(dynamic)
  for (int qi = 0; qi < n_q; ++qi) {            // :117
    std::vector<Packet<D>> pk((size_t)n_db);    // :120,:132
    const S* qp = q + (size_t)qi * dim;
    for (int i = 0; i < n_db; ++i) { pk[i].val = metric(qp, db + (size_t)i * dim, dim); pk[i].index = i; }  // :123-127, indexedSort.hpp:45-49
    const int m = std::min(nn, n_db);           // :130
    std::partial_sort(pk.begin(), pk.begin() + m, pk.end());  // indexedSort.hpp:54
    for (int k = 0; k < m; ++k) { dist[(size_t)qi * nn + k] = pk[k].val; iq[(size_t)qi * nn + k] = qi; idb[(size_t)qi * nn + k] = pk[k].index; }  // :135-139
  }
  return 1;
}

struct Match { uint32_t i, j; float ratio, dist; };   // matching/IndMatch.hpp:25-65 (ALICEVISION_DEBUG_MATCHING is always defined, :18)

struct ByIJ { bool operator()(const Match& a, const Match& b) const { return a.i < b.i || (a.i == b.i && a.j < b.j); } };  // IndMatch.hpp:46

// matching/IndMatch.hpp:52-58  getDeduplicated: std::set on (i,j), first inserted wins, in-order walk
void dedup_ij(std::vector<Match>& v) {
  std::set<Match, ByIJ> s(v.begin(), v.end());
  v.assign(s.begin(), s.end());
}

// matching/IndMatchDecorator.hpp:20-55: matches decorated with the left/right feature positions
struct Deco { float x1, y1, x2, y2; Match m; };
// :46-49 equality on all four coordinates; :34-44 "less": not a strict weak order — kept literally.
struct DecoLess {
  bool operator()(const Deco& a, const Deco& b) const {
    if (a.x1 == b.x1 && a.y1 == b.y1 && a.x2 == b.x2 && a.y2 == b.y2) return false;
    if (a.x1 < b.x1) return a.y1 < b.y1;
    else if (a.x1 > b.x1) return a.y1 < b.y1;
    return a.x1 < b.x1;   // x1 equal -> false
  }
};
// :57-69 decoration order = input order; :84-98 std::set range construction then in-order copy-back
void dedup_deco(std::vector<Match>& v, const float* xyL, const float* xyR) {
  std::vector<Deco> d; d.reserve(v.size());
  for (const Match& m : v) d.push_back(Deco{xyL[2 * m.i], xyL[2 * m.i + 1], xyR[2 * m.j], xyR[2 * m.j + 1], m});
  std::set<Deco, DecoLess> s(d.begin(), d.end());
  v.clear();
  for (const Deco& e : s) v.push_back(e.m);
}

// matching/filters.hpp:35-67  NNdistanceRatio: keep group g iff d[0] < fratio*d[1]; ratio = d[0]/d[1]
// (for unsigned distances: comparison promotes to float, the division is INTEGER division -> 0 or 1).
template <class D> void nn_ratio(const D* d, int n, int nn, float fratio, std::vector<int>& keep, std::vector<float>& ratios) {
  keep.clear(); ratios.clear();
  for (int g = 0; g < n / nn; ++g) {
    const D d0 = d[(size_t)g * nn], d1 = d[(size_t)g * nn + 1];
    if (d0 < fratio * d1) { keep.push_back(g); ratios.push_back((float)(d0 / d1)); }
  }
}

// matching/RegionsMatcher.hpp:126-176  RegionsMatcher::Match (+ RegionsMatcher.cpp:30-39 guards,
// createRegionsMatcher type dispatch RegionsMatcher.cpp:54-176).  Returns -1 where the reference returns false
// with an empty list; the match list otherwise.
int regions_match(int dtype, int hamming, int dim, const void* di, const float* xyi, int ni, const void* dj, const float* xyj, int nj, float ratio,
                  std::vector<Match>& out) {
  out.clear();
  if (nj == 0) return -1;                                        // RegionsMatcher.cpp:32-33
  if ((dtype != DT_BIN && hamming) || (dtype == DT_BIN && !hamming)) return -1;  // null matcher, RegionsMatcher.cpp:61-64,35-36
  const int NN = 2;                                              // RegionsMatcher.hpp:130
  std::vector<int32_t> iq((size_t)nj * NN), idb((size_t)nj * NN);
  std::vector<int> keep; std::vector<float> ratios;
  std::vector<float> df; std::vector<uint32_t> du;
  int ok;
  if (dtype == DT_F32) { df.resize((size_t)nj * NN); ok = knn((const float*)di, ni, (const float*)dj, nj, dim, NN, iq.data(), idb.data(), df.data(), l2_vec_f32); }
  else if (dtype == DT_U8) { df.resize((size_t)nj * NN); ok = knn((const uint8_t*)di, ni, (const uint8_t*)dj, nj, dim, NN, iq.data(), idb.data(), df.data(), l2_vec_u8); }
  else { du.resize((size_t)nj * NN); ok = knn((const uint8_t*)di, ni, (const uint8_t*)dj, nj, dim, NN, iq.data(), idb.data(), du.data(), hamming_u8); }
  if (!ok) return -1;                                            // RegionsMatcher.hpp:135-136
  const float f = (dtype == DT_BIN) ? ratio : ratio * ratio;     // :150 (Square in float, numeric.hpp:130); binary: unsquared (RegionsMatcher.cpp:163)
  if (dtype == DT_BIN) nn_ratio(du.data(), nj * NN, NN, f, keep, ratios); else nn_ratio(df.data(), nj * NN, NN, f, keep, ratios);
  out.reserve(keep.size());
  for (size_t k = 0; k < keep.size(); ++k) {                     // :153-165: IndMatch(i = db index, j = query index, ratio, (float)d1)
    const size_t ix = (size_t)keep[k] * NN;
    out.push_back(Match{(uint32_t)idb[ix], (uint32_t)iq[ix], ratios[k], dtype == DT_BIN ? (float)du[ix] : df[ix]});
  }
  dedup_ij(out);                                                 // :168
  dedup_deco(out, xyi, xyj);                                     // :171-173
  return out.empty() ? -1 : (int)out.size();                     // :175
}

}  // namespace


// ---- guided matching, restated: matching/guidedMatching.hpp:206-268 (cameras == nullptr) with the accumulator :77-118 and the
// error multiview/relativePose/FundamentalError.hpp:52-64.  The 3x3 Eigen expressions are written out with the association a
// coefficient-wise evaluation gives (((a+b)+c), separate multiply / add; built with -ffp-contract=off): Eigen itself is not
// in this image, so THIS sub-expression is unpinned by compiled reference code.  DIST(i, j) is the squared descriptor distance.
template <class DistFn>
static int guided_match_loop(int model, const float* xy_l, int n_l, const float* xy_r, int n_r, const double* F, double errorTh, double distRatio,
                             DistFn DIST, uint32_t* out_ij) {
  std::vector<std::pair<uint32_t, uint32_t>> out;
  for (int i = 0; i < n_l; ++i) {
    const double x0 = (double)xy_l[2 * i], x1 = (double)xy_l[2 * i + 1];                 // GetRegionPosition -> Vec2 (double) of float coordinates
    const double Fx0 = (F[0] * x0 + F[1] * x1) + F[2] * 1.0, Fx1 = (F[3] * x0 + F[4] * x1) + F[5] * 1.0, Fx2 = (F[6] * x0 + F[7] * x1) + F[8] * 1.0;
    const double nrm = Fx0 * Fx0 + Fx1 * Fx1;                                            // F_x.head<2>().squaredNorm()
    double bd = std::numeric_limits<double>::max(), sbd = std::numeric_limits<double>::max(); std::size_t idx = 0;   // distanceRatio(), :86-90
    for (int j = 0; j < n_r; ++j) {
      const double y0 = (double)xy_r[2 * j], y1 = (double)xy_r[2 * j + 1];
      double geomErr;
      if (model == 1) {                                                                  // HomographyAsymmetricError, HomographyError.hpp:23-31
        const double e0 = Fx0 / Fx2, e1 = Fx1 / Fx2;                                     // x2_est = x2h_est.head<2>() / x2h_est[2]
        const double d0 = y0 - e0, d1 = y1 - e1;
        geomErr = d0 * d0 + d1 * d1;                                                     // (x2 - x2_est).squaredNorm()
      } else {
        const double dot = (Fx0 * y0 + Fx1 * y1) + Fx2 * 1.0;                            // F_x.dot(y)
        geomErr = (dot * dot) / nrm;                                                     // Square(.) / squaredNorm, FundamentalError.hpp:62
      }
      if (geomErr < errorTh) {                                                           // guidedMatching.hpp:252
        const double dist = DIST(i, j);
        if (dist < bd) { idx = (std::size_t)j; sbd = dist; std::swap(bd, sbd); }         // update, :95-110
        else if (dist < sbd) sbd = dist;
      }
    }
    if (sbd != std::numeric_limits<double>::max() && bd < distRatio * sbd) out.push_back({(uint32_t)i, (uint32_t)idx});   // isValid :115-118, :259-263
  }
  std::sort(out.begin(), out.end());                                                     // IndMatch::getDeduplicated, :267
  out.erase(std::unique(out.begin(), out.end()), out.end());
  for (size_t k = 0; k < out.size(); ++k) { out_ij[2 * k] = out[k].first; out_ij[2 * k + 1] = out[k].second; }
  return (int)out.size();
}

extern "C" {

struct PortMatch { uint32_t i, j; float ratio, dist; };

int port_num_threads() { return omp_get_max_threads(); }
void port_set_num_threads(int n) { omp_set_num_threads(n); }

double port_metric(int which, int dtype, const void* a, const void* b, int n) {
  if (which == 2) return (double)hamming_u8((const uint8_t*)a, (const uint8_t*)b, n);
  if (dtype == DT_F32) return which == 0 ? (double)l2_simple((const float*)a, (const float*)b, n) : (double)l2_vec_f32((const float*)a, (const float*)b, n);
  return which == 0 ? (double)l2_simple((const uint8_t*)a, (const uint8_t*)b, n) : (double)l2_vec_u8((const uint8_t*)a, (const uint8_t*)b, n);
}

int port_knn_f32(int metric, const float* db, int n_db, const float* q, int n_q, int dim, int nn, int32_t* iq, int32_t* idb, float* dist) {
  if (metric == 0) return knn(db, n_db, q, n_q, dim, nn, iq, idb, dist, l2_simple<float>);
  return knn(db, n_db, q, n_q, dim, nn, iq, idb, dist, l2_vec_f32);
}
int port_knn_u8(const uint8_t* db, int n_db, const uint8_t* q, int n_q, int dim, int nn, int32_t* iq, int32_t* idb, float* dist) {
  return knn(db, n_db, q, n_q, dim, nn, iq, idb, dist, l2_vec_u8);
}
int port_knn_hamming(const uint8_t* db, int n_db, const uint8_t* q, int n_q, int nbytes, int nn, int32_t* iq, int32_t* idb, uint32_t* dist) {
  return knn(db, n_db, q, n_q, nbytes, nn, iq, idb, dist, hamming_u8);
}
// ArrayMatcher_bruteForce.hpp:63-85 SearchNeighbour: first minimum (std::min_element). bit0 = Build ok, bit1 = search ok.
int port_nn1_f32(const float* db, int n_db, const float* q, int dim, int32_t* idx, float* dist) {
  *idx = -1; *dist = -1.f;
  if (n_db < 1) return 0;
  int best = 0; float bd = l2_simple(q, db, dim);
  for (int i = 1; i < n_db; ++i) { float d = l2_simple(q, db + (size_t)i * dim, dim); if (d < bd) { bd = d; best = i; } }
  *idx = best; *dist = bd;
  return 3;
}

int port_nn_ratio_f32(const float* dist, int n, int nn, float fratio, int32_t* keep, float* ratios) {
  std::vector<int> k; std::vector<float> r; nn_ratio(dist, n, nn, fratio, k, r);
  for (size_t i = 0; i < k.size(); ++i) { keep[i] = k[i]; ratios[i] = r[i]; }
  return (int)k.size();
}
int port_nn_ratio_u32(const uint32_t* dist, int n, int nn, float fratio, int32_t* keep, float* ratios) {
  std::vector<int> k; std::vector<float> r; nn_ratio(dist, n, nn, fratio, k, r);
  for (size_t i = 0; i < k.size(); ++i) { keep[i] = k[i]; ratios[i] = r[i]; }
  return (int)k.size();
}

int port_indmatch_dedup(PortMatch* m, int n) {
  std::vector<Match> v(n); std::memcpy(v.data(), m, sizeof(Match) * n);
  dedup_ij(v); std::memcpy(m, v.data(), sizeof(Match) * v.size());
  return (int)v.size();
}
int port_decorator_dedup(PortMatch* m, int n, const float* xyL, int, const float* xyR, int) {
  std::vector<Match> v(n); std::memcpy(v.data(), m, sizeof(Match) * n);
  dedup_deco(v, xyL, xyR); std::memcpy(m, v.data(), sizeof(Match) * v.size());
  return (int)v.size();
}

int port_regions_match(int dtype, int hamming, int dim, const void* di, const float* xyi, int ni, const void* dj, const float* xyj, int nj, float ratio,
                       PortMatch* out) {
  std::vector<Match> v;
  int r = regions_match(dtype, hamming, dim, di, xyi, ni, dj, xyj, nj, ratio, v);
  if (!v.empty()) std::memcpy(out, v.data(), sizeof(Match) * v.size());
  return r;
}

// matchingImageCollection/ImageCollectionMatcher_generic.cpp:30-123 (see ref_oracle.cpp for the output convention)
int port_collection_match(int dtype, int hamming, int dim, int n_views, const void* const* desc, const float* const* xy, const int32_t* nfeat,
                          const uint32_t* pairs, int n_pairs, float ratio, int cross, uint32_t* pair_out, int32_t* counts, PortMatch* out, long cap) {
  (void)n_views;
  std::set<std::pair<uint32_t, uint32_t>> ps;                           // PairSet (types.hpp:23)
  for (int p = 0; p < n_pairs; ++p) ps.insert({pairs[2 * p], pairs[2 * p + 1]});
  std::map<size_t, std::vector<size_t>> grouped;                        // :45-50
  for (auto& p : ps) grouped[p.first].push_back(p.second);
  long n_out = 0; int visited = 0;
  for (auto& g : grouped) {                                             // :53
    const size_t I = g.first;
    for (size_t J : g.second) {
      pair_out[2 * visited] = (uint32_t)I; pair_out[2 * visited + 1] = (uint32_t)J;
      if (nfeat[I] == 0 || nfeat[J] == 0) { counts[visited++] = 0; continue; }   // :59-63, :74-78
      std::vector<Match> v, vc;
      regions_match(dtype, hamming, dim, desc[I], xy[I], nfeat[I], desc[J], xy[J], nfeat[J], ratio, v);            // :81
      if (cross) {                                                      // :83-111
        regions_match(dtype, hamming, dim, desc[J], xy[J], nfeat[J], desc[I], xy[I], nfeat[I], ratio, vc);
        std::set<std::pair<int, int>> check;
        for (auto& m : vc) check.insert({(int)m.i, (int)m.j});
        std::vector<Match> kept;
        for (auto& m : v) if (check.count({(int)m.j, (int)m.i})) kept.push_back(m);
        v.swap(kept);
      }
      if (n_out + (long)v.size() > cap) return -1;
      if (!v.empty()) std::memcpy(out + n_out, v.data(), sizeof(Match) * v.size());
      n_out += (long)v.size();
      counts[visited++] = (int)v.size();                                // :116-119 (0 -> pair not inserted)
    }
  }
  return visited;
}

// ---- file formats around the path, restated with the reference's own stream idiom ------------------------------------
// matches.txt, MatchExporter::saveTxt (matching/io.cpp:281-306): blocks are given in PairwiseMatches map order; every
// block is one (pair, descType) with its matches; consecutive blocks of the same pair share the "I J\n nbDescType" header.
int port_save_matches_txt(const char* path, int n_blocks, const uint32_t* pair_ids, const char* const* desc_names, const int64_t* offsets,
                          const Match* matches) {
  std::ofstream stream(path, std::ios::out);
  if (!stream.is_open()) return -1;
  for (int b = 0; b < n_blocks;) {
    int e = b;
    while (e < n_blocks && pair_ids[2 * e] == pair_ids[2 * b] && pair_ids[2 * e + 1] == pair_ids[2 * b + 1]) ++e;
    const std::size_t I = pair_ids[2 * b], J = pair_ids[2 * b + 1];
    stream << I << " " << J << '\n' << std::size_t(e - b) << '\n';                    // :295
    for (int k = b; k < e; ++k) {
      stream << std::string(desc_names[k]) << " " << std::size_t(offsets[k + 1] - offsets[k]) << '\n';   // :298
      for (int64_t m = offsets[k]; m < offsets[k + 1]; ++m) stream << matches[m].i << " " << matches[m].j << "\n";   // IndMatch.hpp:67
    }
    b = e;
  }
  return stream.good() ? 0 : -2;
}
// matching::LoadMatchFile (io.cpp:41-71): returns the number of (pair, descType) blocks; outputs are filled up to the caps.
int port_load_matches_txt(const char* path, int cap_blocks, uint32_t* pair_ids, char* desc_names /* 32 bytes per block */, int64_t* offsets,
                          long cap_matches, Match* matches) {
  std::ifstream stream(path);
  if (!stream.is_open()) return -1;
  std::size_t I = 0, J = 0, nbDescType = 0;
  int nb = 0; long nm = 0;
  offsets[0] = 0;
  while (stream >> I >> J >> nbDescType) {
    for (std::size_t i = 0; i < nbDescType; ++i) {
      std::string descTypeStr; std::size_t nbMatches = 0;
      stream >> descTypeStr >> nbMatches;
      if (nb >= cap_blocks || nm + (long)nbMatches > cap_matches) return -2;
      for (std::size_t k = 0; k < nbMatches; ++k) { Match m{0, 0, 0.f, 0.f}; stream >> m.i >> m.j; matches[nm++] = m; }
      pair_ids[2 * nb] = (uint32_t)I; pair_ids[2 * nb + 1] = (uint32_t)J;
      std::strncpy(desc_names + 32 * nb, descTypeStr.c_str(), 31); desc_names[32 * nb + 31] = 0;
      offsets[++nb] = nm;
    }
  }
  return nb;
}
// .feat, saveFeatsToFile / loadFeatsFromFile (feature/PointFeature.hpp:78-122)
int port_save_feat(const char* path, const float* feats, int n) {
  std::ofstream file(path);
  if (!file.is_open()) return -1;
  for (int i = 0; i < n; ++i) file << feats[4 * i] << " " << feats[4 * i + 1] << " " << feats[4 * i + 2] << " " << feats[4 * i + 3] << "\n";
  return file.good() ? 0 : -2;
}
int port_load_feat(const char* path, float* feats, int cap) {
  std::ifstream in(path);
  if (!in.is_open()) return -1;
  int n = 0; float a, b, c, d;
  while (in >> a >> b >> c >> d) { if (n < cap) { feats[4 * n] = a; feats[4 * n + 1] = b; feats[4 * n + 2] = c; feats[4 * n + 3] = d; } ++n; }
  return n;
}
// .desc, saveDescsToBinFile / loadDescsFromBinFile (feature/Descriptor.hpp:244-307)
int port_save_desc(const char* path, const void* data, long rows, int row_bytes) {
  std::ofstream file(path, std::ios::out | std::ios::binary);
  if (!file.is_open()) return -1;
  const std::size_t cardDesc = (std::size_t)rows;
  file.write((const char*)&cardDesc, sizeof(std::size_t));
  for (long r = 0; r < rows; ++r) file.write((const char*)data + (size_t)r * row_bytes, row_bytes);
  return file.good() ? 0 : -2;
}
long port_load_desc(const char* path, void* out, long cap_rows, int row_bytes) {
  std::ifstream in(path, std::ios::in | std::ios::binary);
  if (!in.is_open()) return -1;
  std::size_t cardDesc = 0;
  in.read((char*)&cardDesc, sizeof(std::size_t));
  for (long r = 0; r < (long)cardDesc && r < cap_rows; ++r) in.read((char*)out + (size_t)r * row_bytes, row_bytes);
  return (long)cardDesc;
}

// guided matching with the restated metrics: SquaredMetric (feature/Regions.hpp:128-141) = L2_Vectorized for scalar regions,
// SquaredHamming (feature/Hamming.hpp:174-185: h * h in unsigned, returned as double) for binary ones
int port_guided_match(int dtype, int model, const void* desc_l, const float* xy_l, int n_l, const void* desc_r, const float* xy_r, int n_r, const double* F,
                      double errorTh, double distRatio, uint32_t* out_ij) {
  if (dtype == DT_F32) {
    const float* a = (const float*)desc_l; const float* b = (const float*)desc_r;
    return guided_match_loop(model, xy_l, n_l, xy_r, n_r, F, errorTh, distRatio, [a, b](int i, int j) { return (double)l2_vec_f32(a + (size_t)i * 128, b + (size_t)j * 128, 128); }, out_ij);
  }
  const uint8_t* a = (const uint8_t*)desc_l; const uint8_t* b = (const uint8_t*)desc_r;
  if (dtype == DT_U8)
    return guided_match_loop(model, xy_l, n_l, xy_r, n_r, F, errorTh, distRatio, [a, b](int i, int j) { return (double)l2_vec_u8(a + (size_t)i * 128, b + (size_t)j * 128, 128); }, out_ij);
  return guided_match_loop(model, xy_l, n_l, xy_r, n_r, F, errorTh, distRatio,
                           [a, b](int i, int j) { const uint32_t h = hamming_u8(a + (size_t)i * 64, b + (size_t)j * 64, 64); return (double)(h * h); }, out_ij);
}

}  // extern "C"
