// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under alicevision_b200/ may include, link or call it.
//
// CPU statement of the vocabulary-tree pair-list producer (aliceVision_imageMatching, method VocabularyTree, mode a/a).
// Built twice by oracle/Makefile:
//   -DVOC_USE_REFERENCE  -> oracle/_ref/libref_voctree.so: the reference's OWN voctree/VocabularyTree.hpp (quantize, load, save,
//                           computeSparseHistogram) and voctree/VocabularyTree.cpp (sparseDistance), compiled where they lie under
//                           /root/reference/src through oracle/shim;
//   (nothing)            -> oracle/libport_voctree.so: a restatement of the same functions, each citing its lines.
// Restated in both builds, because the reference files need Boost / sfmData:
//   voctree/Database.cpp:44-63,118-137,145-157         insert / find / computeTfIdfWeights
//   voctree/databaseIO.tcc:23-49                       populateDatabase (descriptors given in memory instead of .desc files)
//   imageMatching/ImageMatching.cpp:107-143,191-238    convertAllMatchesToPairList / generateFromVoctree (mode a/a)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <vector>

#ifdef VOC_USE_REFERENCE
#include <aliceVision/voctree/VocabularyTree.hpp>
#include <aliceVision/feature/Descriptor.hpp>
using namespace aliceVision;
using voctree::SparseHistogram;
using voctree::Word;
typedef feature::Descriptor<float, 128> DescriptorFloat;          // imageMatching/ImageMatching.hpp:22-23
typedef feature::Descriptor<unsigned char, 128> DescriptorUChar;
#define PFX(name) refv_##name
#else
typedef int32_t Word;
typedef uint32_t IndexT;
typedef std::map<Word, std::vector<IndexT>> SparseHistogram;      // voctree/VocabularyTree.hpp:55
#define PFX(name) portv_##name
#endif

namespace {

#ifndef VOC_USE_REFERENCE
// voctree/VocabularyTree.hpp:169-196 with voctree/distance.hpp:24-37 (double accumulation, i ascending)
struct Tree {
  uint32_t k = 0, levels = 0, num_words = 0, word_start = 0; int dim = 128;
  std::vector<float> centers; std::vector<uint8_t> valid;
  template <class T> Word quantize(const T* f) const {
    int32_t index = -1;
    for (unsigned level = 0; level < levels; ++level) {
      const int32_t first_child = (index + 1) * (int32_t)k;
      int32_t best_child = first_child;
      double best_distance = std::numeric_limits<double>::max();
      for (int32_t child = first_child; child < first_child + (int32_t)k; ++child) {
        if (!valid[child]) break;
        double r = 0;
        for (int i = 0; i < dim; ++i) { const double diff = (double)f[i] - (double)centers[(size_t)child * dim + i]; r += diff * diff; }
        if (r < best_distance) { best_child = child; best_distance = r; }
      }
      index = best_child;
    }
    return index - (int32_t)word_start;
  }
};
// voctree/VocabularyTree.cpp:22-258, the four methods whose loops are well defined ("weightedStrongCommonPoints" dereferences end iterators, :153-171)
float sparseDistance(const SparseHistogram& v1, const SparseHistogram& v2, const std::string& distanceMethod, const std::vector<float>& word_weights) {
  float distance = 0.f; const float epsilon = 0.001f;
  auto i1 = v1.cbegin(), i1e = v1.cend(); auto i2 = v2.cbegin(), i2e = v2.cend();
  if (distanceMethod == "classic") {
    while (i1 != i1e && i2 != i2e) {
      if (i2->first < i1->first) { distance += i2->second.size(); ++i2; }
      else if (i1->first < i2->first) { distance += i1->second.size(); ++i1; }
      else { const std::pair<std::size_t, std::size_t> val = std::minmax(i1->second.size(), i2->second.size()); distance += static_cast<float>(val.second - val.first); ++i1; ++i2; }
    }
    while (i1 != i1e) { distance += i1->second.size(); ++i1; }
    while (i2 != i2e) { distance += i2->second.size(); ++i2; }
  } else if (distanceMethod == "commonPoints") {
    float score = 0.f;
    while (i1 != i1e && i2 != i2e) {
      if (i2->first < i1->first) ++i2;
      else if (i1->first < i2->first) ++i1;
      else { score += std::min(i1->second.size(), i2->second.size()); ++i1; ++i2; }
    }
    distance = -score;
  } else if (distanceMethod == "strongCommonPoints") {
    float score = 0.f;
    while (i1 != i1e && i2 != i2e) {
      if (i2->first < i1->first) ++i2;
      else if (i1->first < i2->first) ++i1;
      else { if ((std::fabs(i1->second.size() - 1.f) < epsilon) && (std::fabs(i2->second.size() - 1.f) < epsilon)) score += 1; ++i1; ++i2; }
    }
    distance = -score;
  } else if (distanceMethod == "inversedWeightedCommonPoints") {                                   // :192-247
    float score = 0.f;
    std::map<int, int> counter;
    while (i1 != i1e && i2 != i2e) {
      if (i2->first < i1->first) ++i2;
      else if (i1->first < i2->first) ++i1;
      else { counter[i1->first] += std::min(i1->second.size(), i2->second.size()); ++i1; ++i2; }
    }
    for (const auto elem : counter) score += (1.f / elem.second) * word_weights[elem.first];    // ascending word id, float accumulation
    distance = -score;
  } else {
    return std::numeric_limits<float>::quiet_NaN();
  }
  return distance;
}
#endif

struct DocMatch {                                                  // voctree/Database.hpp:28-47
  uint32_t id; float score;
  bool operator<(const DocMatch& other) const { return score < other.score; }
};

struct Database {                                                  // voctree/Database.cpp
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> word_files;   // InvertedFile of (doc id, count)
  std::vector<float> word_weights;
  std::map<uint32_t, SparseHistogram> database;
  explicit Database(uint32_t num_words) : word_files(num_words), word_weights(num_words, 1.0f) {}
  void insert(uint32_t doc_id, const SparseHistogram& document) {                                  // :44-63
    for (auto it = document.begin(); it != document.end(); ++it) {
      auto& file = word_files[it->first];
      if (file.empty() || file.back().first != doc_id) file.push_back({doc_id, (uint32_t)it->second.size()});
      else file.back().second += it->second.size();
    }
    database[doc_id] = document;
  }
  void find(const SparseHistogram& query, std::size_t N, std::vector<DocMatch>& matches, const std::string& method) const {   // :118-137
    matches.clear();
    matches.reserve(database.size());
    for (const auto& document : database) {
#ifdef VOC_USE_REFERENCE
      const float distance = voctree::sparseDistance(query, document.second, method, word_weights);
#else
      const float distance = sparseDistance(query, document.second, method, word_weights);
#endif
      matches.push_back(DocMatch{document.first, distance});
    }
    const std::size_t nMatches = std::min(N, matches.size());
    std::partial_sort(matches.begin(), matches.begin() + nMatches, matches.end());
    matches.resize(nMatches);
  }
  void computeTfIdfWeights(float default_weight = 1.0f) {                                          // :145-157
    const float N = (float)database.size();
    for (std::size_t i = 0; i < word_files.size(); ++i) {
      const std::size_t Ni = word_files[i].size();
      word_weights[i] = Ni != 0 ? std::log(N / Ni) : default_weight;
    }
  }
};

void computeSparseHistogramLocal(const std::vector<Word>& document, SparseHistogram& v) {          // VocabularyTree.hpp:74-81
  for (std::size_t i = 0; i < document.size(); ++i) v[document[i]].push_back((IndexT)i);
}

bool write_tree(const char* path, uint32_t k, uint32_t levels, const float* centers, const uint8_t* valid, uint32_t n_nodes) {
  std::ofstream o(path, std::ios_base::binary);                    // VocabularyTree::save layout (VocabularyTree.hpp:243-256)
  if (!o.is_open()) return false;
  o.write((const char*)&k, 4); o.write((const char*)&levels, 4); o.write((const char*)&n_nodes, 4);
  o.write((const char*)centers, (size_t)n_nodes * 128 * sizeof(float)); o.write((const char*)valid, n_nodes);
  return o.good();
}

#ifdef VOC_USE_REFERENCE
typedef voctree::VocabularyTree<DescriptorFloat> TreeT;
template <class T> std::vector<Word> quantize_all(const TreeT& tree, const T* descs, long n) {
  std::vector<feature::Descriptor<T, 128>> v((size_t)n);
  for (long i = 0; i < n; ++i) for (int k = 0; k < 128; ++k) v[i][k] = descs[(size_t)i * 128 + k];
  return tree.quantize(v);                                         // the reference's own OpenMP loop (VocabularyTree.hpp:198-210)
}
#else
typedef Tree TreeT;
template <class T> std::vector<Word> quantize_all(const TreeT& tree, const T* descs, long n) {
  std::vector<Word> w((size_t)n);
#pragma omp parallel for
  for (long i = 0; i < n; ++i) w[i] = tree.quantize(descs + (size_t)i * 128);
  return w;
}
#endif

bool make_tree(TreeT& tree, uint32_t k, uint32_t levels, const float* centers, const uint8_t* valid, uint32_t n_nodes, const char* tmp_path) {
#ifdef VOC_USE_REFERENCE
  if (!write_tree(tmp_path, k, levels, centers, valid, n_nodes)) return false;
  try { tree.load(tmp_path); } catch (const std::exception&) { return false; }      // the reference's own loader
  return tree.splits() == k && tree.levels() == levels;
#else
  (void)tmp_path;
  tree.k = k; tree.levels = levels; tree.num_words = k; tree.word_start = 0;
  for (uint32_t i = 0; i + 1 < levels; ++i) { tree.word_start += tree.num_words; tree.num_words *= k; }   // setNodeCounts, :285-296
  if (tree.num_words + tree.word_start != n_nodes) return false;
  tree.centers.assign(centers, centers + (size_t)n_nodes * 128); tree.valid.assign(valid, valid + n_nodes);
  return true;
#endif
}

uint32_t tree_words(const TreeT& t) {
#ifdef VOC_USE_REFERENCE
  return t.words();
#else
  return t.num_words;
#endif
}

}  // namespace

extern "C" {

// VocabularyTree::quantize of n descriptors (dtype 0 = float, 1 = uchar; 128-D).  Returns 0, or -1 when the tree is unusable.
int PFX(quantize)(uint32_t k, uint32_t levels, const float* centers, const uint8_t* valid, uint32_t n_nodes, const char* tmp_tree_path,
                  const void* descs, long n, int dtype, int32_t* words) {
  TreeT tree;
  if (!make_tree(tree, k, levels, centers, valid, n_nodes, tmp_tree_path)) return -1;
  std::vector<Word> w = dtype == 0 ? quantize_all(tree, (const float*)descs, n) : quantize_all(tree, (const unsigned char*)descs, n);
  if (n) std::memcpy(words, w.data(), sizeof(int32_t) * (size_t)n);
  return 0;
}

// populateDatabase (uchar descriptors, first nmax when nmax != 0) + computeTfIdfWeights + generateFromVoctree (a/a) +
// convertAllMatchesToPairList.  Outputs: match_ids / scores n_docs x n_keep in ascending doc-id order of the queries,
// weights[num_words], pairs (I, J) rows.  Returns n_keep, or a negative number on failure.
long PFX(image_matching)(uint32_t k, uint32_t levels, const float* centers, const uint8_t* valid, uint32_t n_nodes, const char* tmp_tree_path,
                         int n_docs, const uint32_t* doc_ids, const unsigned char* const* descs, const long* counts, long nmax, long numImageQuery,
                         const char* method, long numMatches, uint32_t* match_ids, float* scores, float* weights, uint32_t* pairs, long cap_pairs,
                         long* n_pairs) {
  TreeT tree;
  if (!make_tree(tree, k, levels, centers, valid, n_nodes, tmp_tree_path)) return -1;
  Database db(tree_words(tree));
  std::map<uint32_t, int> order;                                      // descriptorsFiles: std::map by view id (databaseIO.tcc:29-30)
  for (int d = 0; d < n_docs; ++d) order[doc_ids[d]] = d;
  for (const auto& cur : order) {                                     // populateDatabase, databaseIO.tcc:36-47
    const int d = cur.second;
    const long n = nmax != 0 ? std::min(counts[d], nmax) : counts[d];
    std::vector<Word> doc = quantize_all(tree, descs[d], n);
    SparseHistogram newDoc;
    computeSparseHistogramLocal(doc, newDoc);
    db.insert(cur.first, newDoc);
  }
  db.computeTfIdfWeights();                                            // ImageMatching.cpp:325-331
  if (weights) std::memcpy(weights, db.word_weights.data(), sizeof(float) * db.word_weights.size());
  std::size_t nq = (std::size_t)numImageQuery;
  if (nq == 0) nq = db.database.size();                                // ImageMatching.cpp:198-201
  const std::size_t n_keep = std::min(nq, db.database.size());
  std::map<std::size_t, std::vector<std::size_t>> allMatches;          // PairList
  std::size_t row = 0;
  for (const auto& cur : order) {                                      // generateFromVoctree, :203-237 (mode a/a)
    std::vector<DocMatch> matches;
    db.find(db.database.at(cur.first), nq, matches, method);
    for (std::size_t r = 0; r < matches.size(); ++r) {
      match_ids[row * n_keep + r] = matches[r].id; scores[row * n_keep + r] = matches[r].score;
      allMatches[cur.first].push_back(matches[r].id);
    }
    if (matches.empty()) allMatches[cur.first] = {};
    ++row;
  }
  // convertAllMatchesToPairList, ImageMatching.cpp:107-143
  std::map<std::size_t, std::set<std::size_t>> outPairList;
  std::size_t nm = (std::size_t)numMatches;
  if (nm == 0) nm = allMatches.size();
  for (const auto& match : allMatches) {
    const std::size_t currImageId = match.first;
    std::set<std::size_t> bestMatches;
    for (const std::size_t currMatchId : match.second) {
      if (currMatchId == currImageId) continue;
      if (currMatchId < currImageId) {
        auto currMatches = outPairList.find(currMatchId);
        if (currMatches != outPairList.end() && currMatches->second.find(currImageId) == currMatches->second.end()) bestMatches.insert(currMatchId);
      } else {
        bestMatches.insert(currMatchId);
      }
      if (bestMatches.size() == nm) break;
    }
    if (!bestMatches.empty()) outPairList[currImageId] = bestMatches;
  }
  long n = 0;
  for (const auto& kv : outPairList)
    for (const std::size_t j : kv.second) {
      if (n < cap_pairs) { pairs[2 * n] = (uint32_t)kv.first; pairs[2 * n + 1] = (uint32_t)j; }
      ++n;
    }
  *n_pairs = n;
  return (long)n_keep;
}

}  // extern "C"
