/*
 * b200io — C ABI of the file formats on either side of the descriptor-matching path: what
 * aliceVision_featureMatching reads (per-view .feat / .desc written by featureExtraction) and writes (matches.txt).
 * Host code only (no CUDA); lives in the same shared library as b200match.h.
 *
 *   .desc   feature/Descriptor.hpp:244-307   loadDescsFromBinFile / saveDescsToBinFile: size_t count, then count rows of
 *                                            static_size x bin_type (uchar[128] SIFT, float[128] SIFT_FLOAT, uchar[64] AKAZE_MLDB)
 *   .feat   feature/PointFeature.hpp:78-122  loadFeatsFromFile / saveFeatsToFile: one "x y scale orientation" line per
 *                                            feature, written with the default ostream float format (6 significant digits)
 *   matches matching/io.cpp:41-71,281-306    "I J\n nbDescType\n descType nbMatches\n i j\n ..." per image pair
 *
 * All functions return 0 on success and a B200IO_* code otherwise (the reference throws std::runtime_error on the same
 * conditions: file cannot be opened / written); b200io_last_error() gives a thread-local message.
 */
#ifndef B200IO_H_
#define B200IO_H_

#include <stddef.h>
#include <stdint.h>

#include "b200match.h" /* b200m_match == matching::IndMatch */

#ifdef __cplusplus
extern "C" {
#endif

enum { B200IO_OK = 0, B200IO_ERR_ARG = 1, B200IO_ERR_OPEN = 2, B200IO_ERR_FORMAT = 3, B200IO_ERR_WRITE = 4 };

const char* b200io_last_error(void);

/* ---- .desc (binary) ------------------------------------------------------------------------------------------ */
/* Number of descriptors announced by the file header (Descriptor.hpp:258-259). */
int b200io_desc_count(const char* path, int64_t* count);
/* loadDescsFromBinFile<DescriptorT, FileDescriptorT>: the file holds rows of `dim` elements of `file_dtype`
 * (B200M_F32 or B200M_U8; binary descriptors are B200M_U8 rows of 64), `out` receives min(count, cap_rows) rows of
 * `dim` elements of `out_dtype`, converted element-wise with a plain C++ cast and no rescaling (convertDesc, :209-219).
 * rows_read = rows stored.  A short file leaves the remaining rows zero, like the reference's unchecked reads. */
int b200io_load_desc(const char* path, int dim, int file_dtype, int out_dtype, void* out, int64_t cap_rows, int64_t* rows_read);
/* saveDescsToBinFile (:288-307) */
int b200io_save_desc(const char* path, const void* data, int64_t rows, int dim, int dtype);

/* ---- .feat (text) ---------------------------------------------------------------------------------------------- */
/* loadFeatsFromFile: feats = [x, y, scale, orientation] per feature.  With feats == NULL only counts.  Parsing stops at
 * the first token that is not a number (istream_iterator semantics); a trailing incomplete record is dropped. */
int b200io_load_feat(const char* path, float* feats, int64_t cap, int64_t* count);
/* saveFeatsToFile */
int b200io_save_feat(const char* path, const float* feats, int64_t count);

/* ---- matches.txt ----------------------------------------------------------------------------------------------- */
/* MatchExporter::saveTxt (matching/io.cpp:281-306) for n_pairs image pairs and n_desc descriptor types:
 * pair_ids[2*p] = I, [2*p+1] = J (the caller passes them in PairwiseMatches map order, i.e. sorted);
 * desc_names[d] = EImageDescriberType_enumToString (e.g. "sift"), in MatchesPerDescType map order;
 * offsets[d] has n_pairs+1 entries into matches[d].  A (pair, descType) with zero matches is not listed, a pair with
 * no listed type is not written (the reference's map never holds them: ImageCollectionMatcher_generic.cpp:116-119).
 * Written to a temporary file in the same directory and renamed, like the reference (:284-304). */
int b200io_save_matches_txt(const char* path, int64_t n_pairs, const uint32_t* pair_ids, int n_desc, const char* const* desc_names,
                            const int64_t* const* offsets, const b200m_match* const* matches);

typedef struct b200io_matches b200io_matches;
/* matching::LoadMatchFile (io.cpp:27-78): blocks in file order; distance_ratio / distance of the loaded matches are 0
 * (IndMatch's operator>> reads i and j only, IndMatch.hpp:69). */
int b200io_load_matches_txt(const char* path, b200io_matches** out);
int64_t b200io_matches_num_blocks(const b200io_matches* m);
/* block b: pair (I, J), descriptor type name, match range [begin, end) into the array returned by b200io_matches_data */
int b200io_matches_block(const b200io_matches* m, int64_t b, uint32_t* I, uint32_t* J, const char** desc_name, int64_t* begin, int64_t* end);
const b200m_match* b200io_matches_data(const b200io_matches* m);
void b200io_matches_free(b200io_matches* m);

#ifdef __cplusplus
}
#endif
#endif /* B200IO_H_ */
