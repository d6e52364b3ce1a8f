// Host-side file formats around the matching path (include/b200io.h): .desc / .feat readers and writers and the
// matches.txt exporter / importer, byte-compatible with the reference's stream code
// (feature/Descriptor.hpp:244-307, feature/PointFeature.hpp:78-122, matching/io.cpp:27-78,281-306).
// The text paths format / parse whole buffers (integers by hand, floats with "%g" == the default ostream format) and
// spread large exports over threads, because at ~10^7 matches per run the reference's operator<< loop becomes the wall.
#include "../../include/b200io.h"

#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

namespace {

thread_local std::string g_io_err;
int io_fail(int code, const std::string& msg) { g_io_err = msg; return code; }

struct File {
  FILE* f = nullptr;
  File(const char* path, const char* mode) { f = std::fopen(path, mode); }
  ~File() { if (f) std::fclose(f); }
  File(const File&) = delete;
  File& operator=(const File&) = delete;
};

bool read_all(const char* path, std::string& out) {
  File fp(path, "rb");
  if (!fp.f) return false;
  char buf[1 << 16];
  size_t n;
  while ((n = std::fread(buf, 1, sizeof(buf), fp.f)) > 0) out.append(buf, n);
  return true;
}

inline char* put_u64(char* p, uint64_t v) {
  char tmp[24]; int n = 0;
  do { tmp[n++] = char('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

inline size_t elem_size(int dtype) { return dtype == B200M_F32 ? 4 : 1; }

// skips white space; parses one unsigned integer; returns false at end of input or on a non-digit
inline bool next_u64(const char*& p, const char* end, uint64_t& v) {
  while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
  if (p >= end || *p < '0' || *p > '9') return false;
  v = 0;
  while (p < end && *p >= '0' && *p <= '9') v = v * 10 + uint64_t(*p++ - '0');
  return true;
}
inline bool next_token(const char*& p, const char* end, std::string& tok) {
  while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
  if (p >= end) return false;
  const char* s = p;
  while (p < end && !(*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
  tok.assign(s, p);
  return true;
}

}  // namespace

struct b200io_matches {
  struct Block { uint32_t I, J; std::string desc; int64_t begin, end; };
  std::vector<Block> blocks;
  std::vector<b200m_match> data;
};

extern "C" {

const char* b200io_last_error(void) { return g_io_err.c_str(); }

// ------------------------------------------------------------------------------------------------ .desc
int b200io_desc_count(const char* path, int64_t* count) {
  if (!path || !count) return io_fail(B200IO_ERR_ARG, "null argument");
  File fp(path, "rb");
  if (!fp.f) return io_fail(B200IO_ERR_OPEN, std::string("Can't load descriptor binary file, can't open '") + path + "' !");
  uint64_t n = 0;
  if (std::fread(&n, sizeof(n), 1, fp.f) != 1) n = 0;      // the reference reads into a zero-initialised size_t (Descriptor.hpp:258-259)
  *count = (int64_t)n;
  return B200IO_OK;
}

int b200io_load_desc(const char* path, int dim, int file_dtype, int out_dtype, void* out, int64_t cap_rows, int64_t* rows_read) {
  if (!path || dim < 1 || cap_rows < 0 || (cap_rows > 0 && !out) || (file_dtype != B200M_F32 && file_dtype != B200M_U8) ||
      (out_dtype != B200M_F32 && out_dtype != B200M_U8))
    return io_fail(B200IO_ERR_ARG, "bad arguments");
  File fp(path, "rb");
  if (!fp.f) return io_fail(B200IO_ERR_OPEN, std::string("Can't load descriptor binary file, can't open '") + path + "' !");
  uint64_t n = 0;
  if (std::fread(&n, sizeof(n), 1, fp.f) != 1) n = 0;
  const int64_t rows = std::min<int64_t>((int64_t)n, cap_rows);
  const size_t frow = (size_t)dim * elem_size(file_dtype), orow = (size_t)dim * elem_size(out_dtype);
  if (rows > 0) std::memset(out, 0, (size_t)rows * orow);
  if (file_dtype == out_dtype) {
    (void)!std::fread(out, frow, (size_t)rows, fp.f);
  } else {
    std::vector<unsigned char> buf(frow * 4096);
    for (int64_t r0 = 0; r0 < rows; r0 += 4096) {
      const size_t want = (size_t)std::min<int64_t>(4096, rows - r0);
      const size_t got = std::fread(buf.data(), frow, want, fp.f);
      for (size_t r = 0; r < got; ++r)
        for (int k = 0; k < dim; ++k) {
          if (file_dtype == B200M_U8) reinterpret_cast<float*>(out)[(size_t)(r0 + r) * dim + k] = float(buf[r * frow + k]);
          else reinterpret_cast<unsigned char*>(out)[(size_t)(r0 + r) * dim + k] = (unsigned char)(reinterpret_cast<const float*>(buf.data())[r * dim + k]);
        }
      if (got < want) break;
    }
  }
  if (rows_read) *rows_read = rows;
  return B200IO_OK;
}

int b200io_save_desc(const char* path, const void* data, int64_t rows, int dim, int dtype) {
  if (!path || rows < 0 || dim < 1 || (rows > 0 && !data) || (dtype != B200M_F32 && dtype != B200M_U8 && dtype != B200M_BIN))
    return io_fail(B200IO_ERR_ARG, "bad arguments");
  File fp(path, "wb");
  if (!fp.f) return io_fail(B200IO_ERR_OPEN, std::string("Can't save descriptor binary file, can't open '") + path + "' !");
  const uint64_t n = (uint64_t)rows;
  bool ok = std::fwrite(&n, sizeof(n), 1, fp.f) == 1;
  const size_t row = (size_t)dim * elem_size(dtype);
  if (rows > 0) ok = ok && std::fwrite(data, row, (size_t)rows, fp.f) == (size_t)rows;
  ok = ok && std::fflush(fp.f) == 0;
  if (!ok) return io_fail(B200IO_ERR_WRITE, std::string("Can't save descriptor binary file, '") + path + "' is incorrect !");
  return B200IO_OK;
}

// ------------------------------------------------------------------------------------------------ .feat
int b200io_load_feat(const char* path, float* feats, int64_t cap, int64_t* count) {
  if (!path || !count || cap < 0) return io_fail(B200IO_ERR_ARG, "bad arguments");
  std::string txt;
  if (!read_all(path, txt)) return io_fail(B200IO_ERR_OPEN, std::string("Can't load features file, can't open '") + path + "' !");
  const char* p = txt.c_str();            // NUL-terminated: strtof stops there
  int64_t n = 0;
  for (;;) {
    float rec[4];
    int k = 0;
    for (; k < 4; ++k) {
      char* e = nullptr;
      errno = 0;
      rec[k] = std::strtof(p, &e);
      if (e == p) break;                  // no conversion: end of input or garbage -> istream_iterator stops here
      p = e;
    }
    if (k < 4) break;
    if (feats && n < cap) std::memcpy(feats + 4 * n, rec, sizeof(rec));
    ++n;
  }
  *count = n;
  return B200IO_OK;
}

int b200io_save_feat(const char* path, const float* feats, int64_t count) {
  if (!path || count < 0 || (count > 0 && !feats)) return io_fail(B200IO_ERR_ARG, "bad arguments");
  File fp(path, "wb");
  if (!fp.f) return io_fail(B200IO_ERR_OPEN, std::string("Can't save features file, can't open '") + path + "' !");
  std::string buf;
  buf.reserve(1 << 20);
  char line[128];
  bool ok = true;
  for (int64_t i = 0; i < count; ++i) {
    const float* f = feats + 4 * i;
    const int len = std::snprintf(line, sizeof(line), "%g %g %g %g\n", (double)f[0], (double)f[1], (double)f[2], (double)f[3]);
    buf.append(line, (size_t)len);
    if (buf.size() > (1u << 20) - 256) { ok = ok && std::fwrite(buf.data(), 1, buf.size(), fp.f) == buf.size(); buf.clear(); }
  }
  ok = ok && std::fwrite(buf.data(), 1, buf.size(), fp.f) == buf.size() && std::fflush(fp.f) == 0;
  if (!ok) return io_fail(B200IO_ERR_WRITE, std::string("Can't save features file, '") + path + "' is incorrect !");
  return B200IO_OK;
}

// ------------------------------------------------------------------------------------------------ matches.txt
int b200io_save_matches_txt(const char* path, int64_t n_pairs, const uint32_t* pair_ids, int n_desc, const char* const* desc_names,
                            const int64_t* const* offsets, const b200m_match* const* matches) {
  if (!path || n_pairs < 0 || n_desc < 0 || (n_pairs > 0 && !pair_ids) || (n_desc > 0 && (!desc_names || !offsets || !matches)))
    return io_fail(B200IO_ERR_ARG, "bad arguments");
  for (int d = 0; d < n_desc; ++d) if (!desc_names[d] || !offsets[d]) return io_fail(B200IO_ERR_ARG, "bad arguments");
  // temporary file next to the target, renamed at the end (matching/io.cpp:284-304)
  std::string target(path), tmp;
  {
    const size_t slash = target.find_last_of('/');
    const size_t dot = target.find_last_of('.');
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash);
    const std::string stem = has_ext ? target.substr(0, dot) : target, ext = has_ext ? target.substr(dot) : "";
    char uniq[64];
    std::snprintf(uniq, sizeof(uniq), ".%ld_%lx", (long)getpid(), (unsigned long)(uintptr_t)&target);
    tmp = stem + uniq + ext;
  }
  File fp(tmp.c_str(), "wb");
  if (!fp.f) return io_fail(B200IO_ERR_OPEN, "can't open '" + tmp + "' for writing");
  // format blocks of pairs in parallel (bounded memory), write them in order
  const int nthreads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const int64_t PAIRS_PER_TASK = 64;
  bool ok = true;
  for (int64_t base = 0; base < n_pairs && ok; base += PAIRS_PER_TASK * nthreads) {
    const int ntask = (int)std::min<int64_t>(nthreads, (n_pairs - base + PAIRS_PER_TASK - 1) / PAIRS_PER_TASK);
    std::vector<std::vector<char>> out(ntask);
    auto work = [&](int t) {
      const int64_t p0 = base + t * PAIRS_PER_TASK, p1 = std::min(n_pairs, p0 + PAIRS_PER_TASK);
      size_t need = 0;
      for (int64_t p = p0; p < p1; ++p) {
        need += 64;
        for (int d = 0; d < n_desc; ++d) need += 64 + std::strlen(desc_names[d]) + 22 * (size_t)(offsets[d][p + 1] - offsets[d][p]);
      }
      std::vector<char>& buf = out[t];
      buf.resize(need);
      char* w = buf.data();
      for (int64_t p = p0; p < p1; ++p) {
        int listed = 0;
        for (int d = 0; d < n_desc; ++d) listed += offsets[d][p + 1] > offsets[d][p];
        if (!listed) continue;
        w = put_u64(w, pair_ids[2 * p]); *w++ = ' '; w = put_u64(w, pair_ids[2 * p + 1]); *w++ = '\n';
        w = put_u64(w, (uint64_t)listed); *w++ = '\n';
        for (int d = 0; d < n_desc; ++d) {
          const int64_t a = offsets[d][p], b = offsets[d][p + 1];
          if (b <= a) continue;
          const size_t ln = std::strlen(desc_names[d]);
          std::memcpy(w, desc_names[d], ln); w += ln; *w++ = ' '; w = put_u64(w, (uint64_t)(b - a)); *w++ = '\n';
          const b200m_match* m = matches[d];
          for (int64_t e = a; e < b; ++e) { w = put_u64(w, m[e].i); *w++ = ' '; w = put_u64(w, m[e].j); *w++ = '\n'; }
        }
      }
      buf.resize((size_t)(w - buf.data()));
    };
    std::vector<std::thread> th;
    for (int t = 1; t < ntask; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    for (int t = 0; t < ntask && ok; ++t) ok = out[t].empty() || std::fwrite(out[t].data(), 1, out[t].size(), fp.f) == out[t].size();
  }
  ok = ok && std::fflush(fp.f) == 0;
  std::fclose(fp.f); fp.f = nullptr;
  if (!ok) { std::remove(tmp.c_str()); return io_fail(B200IO_ERR_WRITE, "write to '" + tmp + "' failed"); }
  if (std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); return io_fail(B200IO_ERR_WRITE, std::string("can't rename to '") + path + "'"); }
  return B200IO_OK;
}

int b200io_load_matches_txt(const char* path, b200io_matches** out) {
  if (!path || !out) return io_fail(B200IO_ERR_ARG, "null argument");
  *out = nullptr;
  std::string txt;
  if (!read_all(path, txt)) return io_fail(B200IO_ERR_OPEN, std::string("can't open '") + path + "'");
  try {
    std::unique_ptr<b200io_matches> res(new b200io_matches());
    const char* p = txt.data(); const char* end = p + txt.size();
    uint64_t I, J, nb;
    while (next_u64(p, end, I) && next_u64(p, end, J) && next_u64(p, end, nb)) {     // `while (stream >> I >> J >> nbDescType)`, io.cpp:52
      for (uint64_t d = 0; d < nb; ++d) {
        std::string name; uint64_t cnt = 0;
        if (!next_token(p, end, name) || !next_u64(p, end, cnt)) return io_fail(B200IO_ERR_FORMAT, "truncated descriptor-type header");
        // a match needs at least 4 bytes ("0 0\n"): a count the rest of the file cannot hold is a corrupt / truncated file, not an
        // allocation request (the reference's stream loop would push `cnt` default matches; behind a C ABI that must not be unbounded)
        if (cnt > (uint64_t)(end - p) / 4 + 1) return io_fail(B200IO_ERR_FORMAT, "match count exceeds the remaining file size (truncated or corrupt matches file)");
        b200io_matches::Block b{(uint32_t)I, (uint32_t)J, name, (int64_t)res->data.size(), 0};
        res->data.reserve(res->data.size() + cnt);
        for (uint64_t e = 0; e < cnt; ++e) {
          uint64_t a = 0, c = 0;
          if (!next_u64(p, end, a) || !next_u64(p, end, c)) return io_fail(B200IO_ERR_FORMAT, "truncated match list");
          res->data.push_back(b200m_match{(uint32_t)a, (uint32_t)c, 0.f, 0.f});
        }
        b.end = (int64_t)res->data.size();
        res->blocks.push_back(std::move(b));
      }
    }
    *out = res.release();
  } catch (const std::exception& e) {
    return io_fail(B200IO_ERR_FORMAT, std::string("loading '") + path + "' failed: " + e.what());
  }
  return B200IO_OK;
}

int64_t b200io_matches_num_blocks(const b200io_matches* m) { return m ? (int64_t)m->blocks.size() : 0; }
int b200io_matches_block(const b200io_matches* m, int64_t b, uint32_t* I, uint32_t* J, const char** desc_name, int64_t* begin, int64_t* end) {
  if (!m || b < 0 || b >= (int64_t)m->blocks.size()) return io_fail(B200IO_ERR_ARG, "bad block index");
  const auto& k = m->blocks[(size_t)b];
  if (I) *I = k.I;
  if (J) *J = k.J;
  if (desc_name) *desc_name = k.desc.c_str();
  if (begin) *begin = k.begin;
  if (end) *end = k.end;
  return B200IO_OK;
}
const b200m_match* b200io_matches_data(const b200io_matches* m) { return m && !m->data.empty() ? m->data.data() : nullptr; }
void b200io_matches_free(b200io_matches* m) { delete m; }

}  // extern "C"
