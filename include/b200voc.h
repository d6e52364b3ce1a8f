/*
 * b200voc — C ABI of the pair-list producer that precedes the matching path: the vocabulary-tree
 * image retrieval of aliceVision_imageMatching, which decides WHICH image pairs b200m_match_pairs is given
 * (BASELINE configs[2]).  Reference code replaced:
 *
 *   voctree/VocabularyTree.hpp:169-196    VocabularyTree::quantize   (tree descent, L2 in double, first minimum wins)
 *   voctree/VocabularyTree.hpp:74-81      computeSparseHistogram
 *   voctree/VocabularyTree.cpp:22-258     sparseDistance  ("classic", "commonPoints", "strongCommonPoints", "inversedWeightedCommonPoints")
 *   voctree/Database.cpp:44-137,145-157   Database::insert / find / computeTfIdfWeights
 *   voctree/databaseIO.tcc:23-49          populateDatabase
 *   imageMatching/ImageMatching.cpp:107-143,191-238   convertAllMatchesToPairList / generateFromVoctree (mode a/a)
 *
 * GPU work: the quantisation of every descriptor (sequential double-precision sums, bit-identical to the reference) and
 * the all-against-all document scores (one pass over the inverted file instead of N^2 sparse-vector merges); the top-N
 * selection keeps the reference's std::partial_sort on the host because its tie order is libstdc++-defined.
 * Same conventions as b200match.h: 0 on success, b200m_last_error() for the message, no CPU fallback for the GPU steps.
 */
#ifndef B200VOC_H_
#define B200VOC_H_

#include <stddef.h>
#include <stdint.h>

#include "b200match.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200v_tree b200v_tree; /* voctree::VocabularyTree<Descriptor<float, dim>> */
typedef struct b200v_db b200v_db;     /* voctree::Database bound to one tree and one GPU */

/* ---- vocabulary tree -------------------------------------------------------------------------------------------- */
/* centers: n_nodes x dim floats in the reference's node order (level by level, children of node i at (i+1)*k ..), valid:
 * n_nodes flags (VocabularyTree.hpp:143-144).  n_nodes must equal k + k^2 + ... + k^levels. */
int b200v_tree_create(uint32_t k, uint32_t levels, int dim, const float* centers, const uint8_t* valid, uint32_t n_nodes, b200v_tree** out);
/* VocabularyTree::load / save (VocabularyTree.hpp:243-283): uint32 k, levels, size; size centers of dim floats; size flags */
int b200v_tree_load(const char* path, int dim, b200v_tree** out);
int b200v_tree_save(const b200v_tree* tree, const char* path);
void b200v_tree_destroy(b200v_tree* tree);
uint32_t b200v_tree_levels(const b200v_tree* tree);
uint32_t b200v_tree_splits(const b200v_tree* tree);
uint32_t b200v_tree_words(const b200v_tree* tree);

/* VocabularyTree::quantize(std::vector<DescriptorT>) on GPU `device`: words[n].  dtype = B200M_U8 or B200M_F32 rows of the
 * tree's dimension.  n == 0 is allowed. */
int b200v_quantize(int device, const b200v_tree* tree, const void* descs, int64_t n, int dtype, int32_t* words);

/* ---- database ---------------------------------------------------------------------------------------------------- */
int b200v_db_create(const b200v_tree* tree, int device, b200v_db** out);
void b200v_db_destroy(b200v_db* db);
/* populateDatabase for one view: quantizeToSparse + Database::insert.  nmax != 0 keeps the first nmax descriptors
 * (loadDescsFromBinFile's Nmax).  A doc id may be inserted once. */
int b200v_db_insert_descriptors(b200v_db* db, uint32_t doc_id, const void* descs, int64_t n, int dtype, int64_t nmax);
/* Database::insert of an already quantised document */
int b200v_db_insert_words(b200v_db* db, uint32_t doc_id, const int32_t* words, int64_t n);
int64_t b200v_db_size(const b200v_db* db);
/* the words of an inserted document in feature order (what quantize returned) */
int b200v_db_document(const b200v_db* db, uint32_t doc_id, const int32_t** words, int64_t* n);
/* Database::computeTfIdfWeights; weights (may be NULL) receives tree.words() floats */
int b200v_db_compute_tfidf(b200v_db* db, float default_weight, float* weights);

/* Database::find for EVERY inserted document against the whole database (generateFromVoctree in mode a/a):
 * numImageQuery == 0 means "all" (ImageMatching.cpp:198-201).  Outputs are row-major n_docs x n_keep with
 * n_keep = min(numImageQuery, n_docs): query_ids (ascending doc ids, the std::map walk), match ids and scores in ranked
 * order (score = sparseDistance: smaller is better, negative for the *CommonPoints methods).
 * distanceMethod: "classic", "commonPoints", "strongCommonPoints" (the reference's default) or "inversedWeightedCommonPoints" (float sum in ascending word
 * order per document pair, one thread per pair; "weightedStrongCommonPoints" is undefined behaviour in the reference, VocabularyTree.cpp:153-171, and is refused). */
int b200v_db_query_all(b200v_db* db, size_t numImageQuery, const char* distanceMethod, uint32_t* query_ids, uint32_t* match_ids, float* scores,
                       size_t* n_keep);
/* the raw all-against-all integer score matrix of the last b200v_db_query_all (n_docs x n_docs, row = query): for tests / reuse */
int b200v_db_last_scores(const b200v_db* db, const int32_t** scores, int64_t* n_docs);
/* GPU time (CUDA events) of the scoring step of the last query, ms */
double b200v_db_last_gpu_ms(const b200v_db* db);

/* imageMatching::convertAllMatchesToPairList (ImageMatching.cpp:107-143) on the output of b200v_db_query_all, flattened to
 * (I, J) rows in OrderedPairList order (what main_imageMatching writes and loadPairs reads back).  Call with pairs == NULL to
 * get the count. */
int b200v_convert_matches_to_pairs(const uint32_t* query_ids, const uint32_t* match_ids, size_t n_docs, size_t n_keep, size_t numMatches,
                                   uint32_t* pairs, int64_t cap_pairs, int64_t* n_pairs);

#ifdef __cplusplus
}
#endif
#endif /* B200VOC_H_ */
