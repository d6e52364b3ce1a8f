// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under
// alicevision_b200/ may include, link or call this file.
//
// This translation unit #includes the reference's OWN headers, untouched, from
// /root/reference/src (through the stub include dir oracle/shim, which only replaces
// Eigen/Boost-facing plumbing headers) and exposes them through a C ABI so the
// pytest suite can run the real `ArrayMatcher_bruteForce`, `NNdistanceRatio`,
// `RegionsMatcher<...>::Match`, `IndMatch::getDeduplicated` and `IndMatchDecorator`
// on the same inputs as the CUDA path.  It is built by oracle/Makefile into
// oracle/_ref/libref_oracle.so (git-ignored, travels to the GPU box prebuilt).
//
// Also compiled verbatim, as a CPU TIMING BASELINE only (the north star names it): matching/ArrayMatcher_cascadeHashing.hpp +
// matching/CascadeHasher.hpp on top of the shim's plain-loop MatrixXf/VectorXf (not Eigen's kernels => its results are unpinned).
// Reference code exercised (all verbatim, nothing copied into this repo):
//   matching/ArrayMatcher_bruteForce.hpp:42-142   Build / SearchNeighbour(s)
//   feature/metric.hpp:27-139                     L2_Simple, L2_Vectorized (+SSE float)
//   feature/Hamming.hpp:76-172                    Hamming<unsigned char>
//   stl/indexedSort.hpp:40-55                     partial_sort on (val,index) packets
//   matching/filters.hpp:35-67                    NNdistanceRatio
//   matching/RegionsMatcher.hpp:83-177            RegionsMatcher<ArrayMatcherT>::Match
//   matching/IndMatch.hpp:52-58                   IndMatch::getDeduplicated
//   matching/IndMatchDecorator.hpp:20-98          IndMatchDecorator<float>
// Restated here because the reference .cpp files need Boost/FLANN/real Eigen:
//   matching/RegionsMatcher.cpp:54-176            createRegionsMatcher (4 brute-force cases)
//   matchingImageCollection/ImageCollectionMatcher_generic.cpp:30-123  the pair loop
#include <aliceVision/matching/ArrayMatcher_bruteForce.hpp>
#include <aliceVision/matching/ArrayMatcher_cascadeHashing.hpp>
#include <aliceVision/matching/RegionsMatcher.hpp>
#include <aliceVision/matching/IndMatch.hpp>
#include <aliceVision/matching/IndMatchDecorator.hpp>
#include <aliceVision/matching/filters.hpp>
#include <aliceVision/feature/metric.hpp>
#include <aliceVision/feature/Hamming.hpp>
#include <aliceVision/feature/regionsFactory.hpp>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <vector>
#include <omp.h>

using namespace aliceVision;
using namespace aliceVision::matching;
using namespace aliceVision::feature;

namespace {

// 16-byte aligned copy: l2_sse uses _mm_load_ps (metric.hpp:105-106).
template <class T> struct AlignedBuf {
  T* p = nullptr;
  AlignedBuf(const void* src, size_t n) {
    size_t bytes = std::max<size_t>(n * sizeof(T), 16);
    if (posix_memalign((void**)&p, 64, bytes) != 0) p = nullptr;
    if (p && src && n) std::memcpy(p, src, n * sizeof(T));
  }
  ~AlignedBuf() { free(p); }
};

enum { DT_F32 = 0, DT_U8 = 1, DT_BIN = 2 };

template <class Matcher, class Scalar, class Dist>
int knn_impl(const void* db, int n_db, const void* q, int n_q, int dim, int nn, int32_t* idx_q, int32_t* idx_db, Dist* dist) {
  std::mt19937 rng;
  AlignedBuf<Scalar> a(db, (size_t)std::max(n_db, 0) * dim), b(q, (size_t)std::max(n_q, 0) * dim);
  Matcher m;
  if (!m.Build(rng, a.p, n_db, dim)) return 0;
  IndMatches ind;
  std::vector<typename Matcher::DistanceType> d;
  if (!m.SearchNeighbours(b.p, n_q, &ind, &d, (size_t)nn)) return 0;
  for (size_t k = 0; k < ind.size(); ++k) {
    idx_q[k] = (int32_t)ind[k]._i;   // query index (ArrayMatcher_bruteForce.hpp:138)
    idx_db[k] = (int32_t)ind[k]._j;  // database index
    dist[k] = (Dist)d[k];
  }
  return 1;
}

template <class RegionsT, class Scalar>
std::unique_ptr<RegionsT> make_regions(const void* desc, const float* xy, int n) {
  std::unique_ptr<RegionsT> r(new RegionsT());
  r->Features().reserve(n);
  r->Descriptors().resize(n);
  for (int i = 0; i < n; ++i) r->Features().emplace_back(xy[2 * i], xy[2 * i + 1], 1.0f, 0.0f);
  if (n) std::memcpy((void*)r->DescriptorRawData(), desc, (size_t)n * sizeof(typename RegionsT::DescriptorT));
  return r;
}

std::unique_ptr<Regions> make_any_regions(int dtype, const void* desc, const float* xy, int n) {
  switch (dtype) {
    case DT_F32: return make_regions<SIFT_Float_Regions, float>(desc, xy, n);
    case DT_U8: return make_regions<SIFT_Regions, unsigned char>(desc, xy, n);
    case DT_BIN: return make_regions<AKAZE_BinaryRegions, unsigned char>(desc, xy, n);
  }
  return nullptr;
}

// Restatement of createRegionsMatcher (RegionsMatcher.cpp:54-176), brute-force cases only:
//   uchar scalar  + BRUTE_FORCE_L2      -> bruteForce<uchar, L2_Vectorized<uchar>>, squared=true  (:74-79)
//   float scalar  + BRUTE_FORCE_L2      -> bruteForce<float, L2_Vectorized<float>>, squared=true  (:103-108)
//   binary uchar  + BRUTE_FORCE_HAMMING -> bruteForce<uchar, Hamming<uchar>>,       squared=false (:159-164)
// invalid combinations return null (:61-64).
std::unique_ptr<IRegionsMatcher> create_matcher(std::mt19937& rng, const Regions& regions, int hamming) {
  std::unique_ptr<IRegionsMatcher> out;
  if (regions.IsScalar() && hamming) return out;
  if (regions.IsBinary() && !hamming) return out;
  if (regions.IsScalar()) {
    if (regions.Type_id() == typeid(unsigned char).name())
      out.reset(new RegionsMatcher<ArrayMatcher_bruteForce<unsigned char, L2_Vectorized<unsigned char>>>(rng, regions, true));
    else if (regions.Type_id() == typeid(float).name())
      out.reset(new RegionsMatcher<ArrayMatcher_bruteForce<float, L2_Vectorized<float>>>(rng, regions, true));
  } else if (regions.IsBinary() && regions.Type_id() == typeid(unsigned char).name()) {
    out.reset(new RegionsMatcher<ArrayMatcher_bruteForce<unsigned char, Hamming<unsigned char>>>(rng, regions, false));
  }
  return out;
}

// RegionsDatabaseMatcher::Match (RegionsMatcher.cpp:30-39)
bool db_match(IRegionsMatcher* m, float ratio, const Regions& q, IndMatches& out) {
  if (q.RegionCount() == 0) return false;
  if (!m) return false;
  return m->Match(ratio, q, out);
}

}  // namespace

// ---- the reference's own region file IO (feature/Regions.hpp:166-179 -> PointFeature.hpp:88-122, Descriptor.hpp:244-307) ----
// feats: n x 4 (x, y, scale, orientation).  dtype as above.
template <class RegionsT, class T>
static int save_regions_t(const void* desc, const float* feats, int n, const char* featPath, const char* descPath) {
  RegionsT r;
  const T* d = static_cast<const T*>(desc);
  for (int i = 0; i < n; ++i) {
    r.Features().emplace_back(feats[4 * i], feats[4 * i + 1], feats[4 * i + 2], feats[4 * i + 3]);
    typename RegionsT::DescriptorT v;
    for (int k = 0; k < (int)RegionsT::DescriptorT::static_size; ++k) v[k] = d[(size_t)i * RegionsT::DescriptorT::static_size + k];
    r.Descriptors().push_back(v);
  }
  try { r.Save(featPath, descPath); } catch (const std::exception&) { return -1; }
  return n;
}
template <class RegionsT, class T>
static int load_regions_t(const char* featPath, const char* descPath, void* desc, float* feats, int cap) {
  RegionsT r;
  try { r.Load(featPath, descPath); } catch (const std::exception&) { return -1; }
  const int n = (int)r.RegionCount();
  const int nf = (int)r.Features().size();
  T* d = static_cast<T*>(desc);
  for (int i = 0; i < std::min(n, cap); ++i)
    for (int k = 0; k < (int)RegionsT::DescriptorT::static_size; ++k) d[(size_t)i * RegionsT::DescriptorT::static_size + k] = r.Descriptors()[i][k];
  for (int i = 0; i < std::min(nf, cap); ++i) {
    const auto& f = r.Features()[i];
    feats[4 * i] = f.x(); feats[4 * i + 1] = f.y(); feats[4 * i + 2] = f.scale(); feats[4 * i + 3] = f.orientation();
  }
  return n == nf ? n : -2;
}


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

struct RefMatch { uint32_t i, j; float ratio, dist; };  // == matching::IndMatch (IndMatch.hpp:60-64)

int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

// ---- metric known-answer hooks (feature/metric_test.cpp) ------------------------------------
double ref_metric(int which, int dtype, const void* a, const void* b, int n) {
  // which: 0 = L2_Simple, 1 = L2_Vectorized, 2 = Hamming<unsigned char>
  if (which == 2) return (double)Hamming<unsigned char>()((const unsigned char*)a, (const unsigned char*)b, (size_t)n);
  if (dtype == DT_F32) {
    AlignedBuf<float> x(a, n), y(b, n);
    return which == 0 ? (double)L2_Simple<float>()(x.p, y.p, (size_t)n) : (double)L2_Vectorized<float>()(x.p, y.p, (size_t)n);
  }
  const unsigned char* x = (const unsigned char*)a; const unsigned char* y = (const unsigned char*)b;
  return which == 0 ? (double)L2_Simple<unsigned char>()(x, y, (size_t)n) : (double)L2_Vectorized<unsigned char>()(x, y, (size_t)n);
}

// ---- ArrayMatcher_bruteForce::SearchNeighbours ----------------------------------------------
// metric: 0 = L2_Simple (class default, used by matching_test.cpp), 1 = L2_Vectorized (pipeline)
int ref_knn_f32(int metric, const float* db, int n_db, const float* q, int n_q, int dim, int nn, int32_t* iq, int32_t* idb, float* dist) {
  if (metric == 0) return knn_impl<ArrayMatcher_bruteForce<float>, float, float>(db, n_db, q, n_q, dim, nn, iq, idb, dist);
  return knn_impl<ArrayMatcher_bruteForce<float, L2_Vectorized<float>>, float, float>(db, n_db, q, n_q, dim, nn, iq, idb, dist);
}
int ref_knn_u8(const uint8_t* db, int n_db, const uint8_t* q, int n_q, int dim, int nn, int32_t* iq, int32_t* idb, float* dist) {
  return knn_impl<ArrayMatcher_bruteForce<unsigned char, L2_Vectorized<unsigned char>>, unsigned char, float>(db, n_db, q, n_q, dim, nn, iq, idb, dist);
}
int ref_knn_hamming(const uint8_t* db, int n_db, const uint8_t* q, int n_q, int nbytes, int nn, int32_t* iq, int32_t* idb, uint32_t* dist) {
  return knn_impl<ArrayMatcher_bruteForce<unsigned char, Hamming<unsigned char>>, unsigned char, uint32_t>(db, n_db, q, n_q, nbytes, nn, iq, idb, dist);
}
// SearchNeighbour (single 1-NN, first minimum on ties; ArrayMatcher_bruteForce.hpp:63-85). Returns 0 when not built.
int ref_nn1_f32(const float* db, int n_db, const float* q, int dim, int32_t* idx, float* dist) {
  std::mt19937 rng;
  AlignedBuf<float> a(db, (size_t)std::max(n_db, 0) * dim), b(q, dim);
  ArrayMatcher_bruteForce<float> m;
  bool built = m.Build(rng, a.p, n_db, dim);
  int i = -1; float d = -1.f;
  bool ok = m.SearchNeighbour(b.p, &i, &d);
  *idx = i; *dist = d;
  return (built ? 1 : 0) | (ok ? 2 : 0);
}

// ---- NNdistanceRatio (filters.hpp:35-67) -----------------------------------------------------
int ref_nn_ratio_f32(const float* dist, int n, int nn, float fratio, int32_t* keep, float* ratios) {
  std::vector<float> d(dist, dist + n); std::vector<int> k; std::vector<float> r;
  NNdistanceRatio(d.begin(), d.end(), nn, k, fratio, &r);
  for (size_t i = 0; i < k.size(); ++i) { keep[i] = k[i]; ratios[i] = r[i]; }
  return (int)k.size();
}
int ref_nn_ratio_u32(const uint32_t* dist, int n, int nn, float fratio, int32_t* keep, float* ratios) {
  std::vector<unsigned int> d(dist, dist + n); std::vector<int> k; std::vector<float> r;
  NNdistanceRatio(d.begin(), d.end(), nn, k, fratio, &r);
  for (size_t i = 0; i < k.size(); ++i) { keep[i] = k[i]; ratios[i] = r[i]; }
  return (int)k.size();
}

// ---- the two de-duplications -----------------------------------------------------------------
int ref_indmatch_dedup(RefMatch* m, int n) {
  IndMatches v; v.reserve(n);
  for (int k = 0; k < n; ++k) v.emplace_back(m[k].i, m[k].j, m[k].ratio, m[k].dist);
  IndMatch::getDeduplicated(v);
  for (size_t k = 0; k < v.size(); ++k) m[k] = RefMatch{v[k]._i, v[k]._j, v[k]._distanceRatio, v[k]._distance};
  return (int)v.size();
}
int ref_decorator_dedup(RefMatch* m, int n, const float* xyL, int nL, const float* xyR, int nR) {
  IndMatches v; v.reserve(n);
  for (int k = 0; k < n; ++k) v.emplace_back(m[k].i, m[k].j, m[k].ratio, m[k].dist);
  PointFeatures L, R; L.reserve(nL); R.reserve(nR);
  for (int k = 0; k < nL; ++k) L.emplace_back(xyL[2 * k], xyL[2 * k + 1], 1.f, 0.f);
  for (int k = 0; k < nR; ++k) R.emplace_back(xyR[2 * k], xyR[2 * k + 1], 1.f, 0.f);
  IndMatchDecorator<float> deco(v, L, R);
  deco.getDeduplicated(v);
  for (size_t k = 0; k < v.size(); ++k) m[k] = RefMatch{v[k]._i, v[k]._j, v[k]._distanceRatio, v[k]._distance};
  return (int)v.size();
}

// ---- RegionsMatcher<...>::Match on real reference Regions ------------------------------------
// dtype: 0 SIFT_Float_Regions (float x128), 1 SIFT_Regions (uchar x128), 2 AKAZE_BinaryRegions (64 B).
// hamming: 0 -> BRUTE_FORCE_L2, 1 -> BRUTE_FORCE_HAMMING. Returns the number of matches, or -1 when the
// reference returns false / has no matcher for the combination. `out` must hold n_j entries.
int ref_regions_match(int dtype, int hamming, const void* desc_i, const float* xy_i, int n_i, const void* desc_j, const float* xy_j, int n_j,
                      float ratio, RefMatch* out) {
  std::mt19937 rng;
  std::unique_ptr<Regions> ri = make_any_regions(dtype, desc_i, xy_i, n_i), rj = make_any_regions(dtype, desc_j, xy_j, n_j);
  std::unique_ptr<IRegionsMatcher> m = create_matcher(rng, *ri, hamming);
  IndMatches v;
  bool ok = db_match(m.get(), ratio, *rj, v);
  for (size_t k = 0; k < v.size(); ++k) out[k] = RefMatch{v[k]._i, v[k]._j, v[k]._distanceRatio, v[k]._distance};
  return ok ? (int)v.size() : (v.empty() ? -1 : (int)v.size());
}

// ---- restated ImageCollectionMatcher_generic::Match (ImageCollectionMatcher_generic.cpp:30-123) --------
// views: per view descriptor pointer / positions / count.  pairs: n_pairs x 2 view indices (any order; grouped by
// first index through std::map exactly as :45-50).  Output: for each input pair p (in the order of the sorted
// PairSet walk) pair_out[p] = (I,J), counts[p] = #matches (0 => the reference would not insert the pair),
// matches appended to `out` (capacity cap). Returns number of pairs visited, or -1 on overflow.
int ref_collection_match(int dtype, int hamming, int n_views, const void* const* desc, const float* const* xy, const int32_t* counts_per_view,
                         const uint32_t* pairs, int n_pairs, float ratio, int cross, uint32_t* pair_out, int32_t* counts, RefMatch* out,
                         long cap) {
  std::mt19937 rng;
  std::vector<std::unique_ptr<Regions>> regs(n_views);
  for (int v = 0; v < n_views; ++v) regs[v] = make_any_regions(dtype, desc[v], xy[v], counts_per_view[v]);
  PairSet ps;
  for (int p = 0; p < n_pairs; ++p) ps.insert(Pair(pairs[2 * p], pairs[2 * p + 1]));
  std::map<size_t, std::vector<size_t>> grouped;
  for (const Pair& p : ps) grouped[p.first].push_back(p.second);
  long n_out = 0; int visited = 0;
  for (auto& g : grouped) {
    const size_t I = g.first;
    const Regions& regionsI = *regs.at(I);
    if (regionsI.RegionCount() == 0) {
      for (size_t J : g.second) { pair_out[2 * visited] = (uint32_t)I; pair_out[2 * visited + 1] = (uint32_t)J; counts[visited++] = 0; }
      continue;
    }
    std::unique_ptr<IRegionsMatcher> matcher = create_matcher(rng, regionsI, hamming);
    for (size_t J : g.second) {
      const Regions& regionsJ = *regs.at(J);
      pair_out[2 * visited] = (uint32_t)I; pair_out[2 * visited + 1] = (uint32_t)J;
      if (regionsJ.RegionCount() == 0 || regionsI.Type_id() != regionsJ.Type_id()) { counts[visited++] = 0; continue; }
      IndMatches vec;
      db_match(matcher.get(), ratio, regionsJ, vec);
      if (cross) {
        std::unique_ptr<IRegionsMatcher> matcherCross = create_matcher(rng, regionsJ, hamming);
        IndMatches vecCross;
        db_match(matcherCross.get(), ratio, regionsI, vecCross);
        std::map<std::pair<int, int>, IndMatch> check;
        for (IndMatch& m : vecCross) check[std::make_pair((int)m._i, (int)m._j)] = m;
        IndMatches checked;
        for (IndMatch& m : vec)
          if (check.find(std::make_pair((int)m._j, (int)m._i)) != check.end()) checked.push_back(m);
        std::swap(vec, checked);
      }
      if (n_out + (long)vec.size() > cap) return -1;
      for (auto& m : vec) out[n_out++] = RefMatch{m._i, m._j, m._distanceRatio, m._distance};
      counts[visited++] = (int)vec.size();
    }
  }
  return visited;
}

int ref_save_regions(int dtype, const void* desc, const float* feats, int n, const char* featPath, const char* descPath) {
  if (dtype == 0) return save_regions_t<SIFT_Float_Regions, float>(desc, feats, n, featPath, descPath);
  if (dtype == 1) return save_regions_t<SIFT_Regions, unsigned char>(desc, feats, n, featPath, descPath);
  return save_regions_t<AKAZE_BinaryRegions, unsigned char>(desc, feats, n, featPath, descPath);
}
int ref_load_regions(int dtype, const char* featPath, const char* descPath, void* desc, float* feats, int cap) {
  if (dtype == 0) return load_regions_t<SIFT_Float_Regions, float>(featPath, descPath, desc, feats, cap);
  if (dtype == 1) return load_regions_t<SIFT_Regions, unsigned char>(featPath, descPath, desc, feats, cap);
  return load_regions_t<AKAZE_BinaryRegions, unsigned char>(featPath, descPath, desc, feats, cap);
}
// loadDescsFromBinFile<DescriptorT, FileDescriptorT> with a type conversion (Descriptor.hpp:221-283): uchar file -> float memory
int ref_load_desc_u8_as_f32(const char* descPath, float* out, int cap) {
  std::vector<Descriptor<float, 128>> v;
  try { loadDescsFromBinFile<Descriptor<float, 128>, Descriptor<unsigned char, 128>>(descPath, v); } catch (const std::exception&) { return -1; }
  for (int i = 0; i < std::min((int)v.size(), cap); ++i) for (int k = 0; k < 128; ++k) out[(size_t)i * 128 + k] = v[i][k];
  return (int)v.size();
}

// guided matching with the reference's own Regions::SquaredDescriptorDistance (feature/Regions.hpp:198-207 -> SquaredMetric, :128-141)
int ref_guided_match(int dtype, int model, const void* desc_l, const float* xy_l, int n_l, const void* desc_r, const float* xy_r, int n_r, const double* F,
                     double errorTh, double distRatio, uint32_t* out_ij) {
  std::unique_ptr<Regions> rl = make_any_regions(dtype, desc_l, xy_l, n_l), rr = make_any_regions(dtype, desc_r, xy_r, n_r);
  const Regions* l = rl.get(); const Regions* r = rr.get();
  return guided_match_loop(model, xy_l, n_l, xy_r, n_r, F, errorTh, distRatio, [l, r](int i, int j) { return l->SquaredDescriptorDistance((size_t)i, r, (size_t)j); }, out_ij);
}

// ---- CASCADE_HASHING_L2 through the restated collection loop (ImageCollectionMatcher_generic.cpp:30-123): the matcher (= the hashed
// database) is built once per database image, and the loop over its J images runs in parallel exactly as the reference enables it
// for this matcher type (:39,:68 `#pragma omp parallel for schedule(dynamic) if (matcherType == CASCADE_HASHING_L2)`).
// Returns the total number of matches (or -1); counts[p] per visited pair.  Timing baseline: see the header comment.
long ref_collection_cascade(int dtype, int n_views, const void* const* desc, const float* const* xy, const int32_t* counts_per_view,
                            const uint32_t* pairs, int n_pairs, float ratio, unsigned seed, int32_t* counts) {
  if (dtype == 2) return -1;                                   // RegionsMatcher.cpp:63-64: binary regions need BRUTE_FORCE_HAMMING
  std::mt19937 rng(seed);
  std::vector<std::unique_ptr<Regions>> regs(n_views);
  for (int v = 0; v < n_views; ++v) regs[v] = make_any_regions(dtype, desc[v], xy[v], counts_per_view[v]);
  PairSet ps;
  for (int p = 0; p < n_pairs; ++p) ps.insert(Pair(pairs[2 * p], pairs[2 * p + 1]));
  std::map<size_t, std::vector<size_t>> grouped;
  for (const Pair& p : ps) grouped[p.first].push_back(p.second);
  long total = 0; int base = 0;
  for (auto& g : grouped) {
    const Regions& regionsI = *regs.at(g.first);
    const std::vector<size_t>& indexToCompare = g.second;
    if (regionsI.RegionCount() == 0) { for (size_t j = 0; j < indexToCompare.size(); ++j) counts[base + j] = 0; base += (int)indexToCompare.size(); continue; }
    std::unique_ptr<IRegionsMatcher> matcher;                  // createRegionsMatcher, RegionsMatcher.cpp:87-92,116-121
    if (dtype == 1) matcher.reset(new RegionsMatcher<ArrayMatcher_cascadeHashing<unsigned char, L2_Vectorized<unsigned char>>>(rng, regionsI, true));
    else matcher.reset(new RegionsMatcher<ArrayMatcher_cascadeHashing<float, L2_Vectorized<float>>>(rng, regionsI, true));
    long sub = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : sub)
    for (int j = 0; j < (int)indexToCompare.size(); ++j) {
      const Regions& regionsJ = *regs.at(indexToCompare[j]);
      IndMatches vec;
      if (regionsJ.RegionCount() != 0 && regionsI.Type_id() == regionsJ.Type_id()) matcher->Match(ratio, regionsJ, vec);
      counts[base + j] = (int)vec.size();
      sub += (long)vec.size();
    }
    total += sub; base += (int)indexToCompare.size();
  }
  return total;
}

}  // extern "C"
