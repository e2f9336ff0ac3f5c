// On-disk formats at the edges of the hot path (host code only):
//   * headerless CSV of doubles            K/loaders/CsvDataLoader.scala:28-30 (row.split(",").map(_.toDouble))
//   * TIMIT sparse label files "row label" K/loaders/TimitFeaturesDataLoader.scala:22-42 (1-based row, 1-based label)
//   * CIFAR-10 binary records 1 + 3072 B   K/loaders/CifarLoader.scala:30-45
//   * fitted BlockLinearMapper <-> flat file, replacing the Java-serialised FittedPipeline (K/workflow/FittedPipeline.scala:18-22)
// The parsers fill caller-owned host buffers (pinned if the caller wants an asynchronous upload); the CSV parser splits the
// file by lines over hardware threads -- the reference parses with one Scala closure per line on the Spark executors.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <thread>
#include <vector>

#include "engine.h"

namespace ks {

static thread_local std::string g_io_err;

struct MappedFile {
  const char* p = nullptr;
  size_t n = 0;
  int fd = -1;
  ~MappedFile() {
    if (p && n) munmap(const_cast<char*>(p), n);
    if (fd >= 0) close(fd);
  }
  bool open_ro(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) {
      g_io_err = std::string("cannot open ") + path + ": " + strerror(errno);
      return false;
    }
    struct stat st;
    if (fstat(fd, &st) != 0) {
      g_io_err = std::string("cannot stat ") + path;
      return false;
    }
    n = static_cast<size_t>(st.st_size);
    if (n == 0) return true;
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      g_io_err = std::string("cannot mmap ") + path;
      n = 0;
      return false;
    }
    p = static_cast<const char*>(m);
    return true;
  }
};

// line starts of a text buffer (empty trailing line ignored)
static std::vector<size_t> line_starts(const char* p, size_t n) {
  std::vector<size_t> s;
  size_t i = 0;
  while (i < n) {
    s.push_back(i);
    const void* nl = memchr(p + i, '\n', n - i);
    if (!nl) break;
    i = static_cast<size_t>(static_cast<const char*>(nl) - p) + 1;
  }
  while (!s.empty()) {  // drop blank trailing lines
    size_t b = s.back();
    bool blank = true;
    for (size_t q = b; q < n && p[q] != '\n'; ++q)
      if (p[q] != ' ' && p[q] != '\r' && p[q] != '\t') { blank = false; break; }
    if (!blank) break;
    s.pop_back();
  }
  return s;
}

static int64_t count_fields(const char* p, size_t b, size_t n) {
  int64_t c = 1;
  for (size_t i = b; i < n && p[i] != '\n'; ++i) c += p[i] == ',';
  return c;
}

}  // namespace ks

using namespace ks;

extern "C" {

KS_API const char* ks_io_last_error(void) { return g_io_err.c_str(); }

KS_API int32_t ks_csv_dims(const char* path, int64_t* n_rows, int64_t* n_cols) {
  if (!path || !n_rows || !n_cols) return KS_ERR_INVALID;
  MappedFile f;
  if (!f.open_ro(path)) return KS_ERR_INVALID;
  auto ls = line_starts(f.p, f.n);
  *n_rows = static_cast<int64_t>(ls.size());
  *n_cols = ls.empty() ? 0 : count_fields(f.p, ls[0], f.n);
  return KS_OK;
}

// out is row-major [n_rows][ld]; is_f64 selects double / float elements.  Every row must have exactly n_cols fields.
static int32_t csv_read(const char* path, void* out, int64_t n_rows, int64_t n_cols, int64_t ld, bool is_f64) {
  if (!path || !out || n_rows < 0 || n_cols <= 0 || ld < n_cols) return KS_ERR_INVALID;
  MappedFile f;
  if (!f.open_ro(path)) return KS_ERR_INVALID;
  auto ls = line_starts(f.p, f.n);
  if (static_cast<int64_t>(ls.size()) != n_rows) {
    g_io_err = "csv has " + std::to_string(ls.size()) + " rows, expected " + std::to_string(n_rows);
    return KS_ERR_INVALID;
  }
  const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), n_rows / 256 + 1)));
  std::vector<std::string> errs(nt);
  auto work = [&](int ti) {
    const int64_t r0 = n_rows * ti / nt, r1 = n_rows * (ti + 1) / nt;
    std::string tok;
    for (int64_t r = r0; r < r1; ++r) {
      size_t i = ls[r];
      const size_t end = (r + 1 < n_rows) ? ls[r + 1] : f.n;
      for (int64_t c = 0; c < n_cols; ++c) {
        size_t j = i;
        while (j < end && f.p[j] != ',' && f.p[j] != '\n' && f.p[j] != '\r') ++j;
        if (j == i) {
          errs[ti] = "empty field at row " + std::to_string(r) + " column " + std::to_string(c);
          return;
        }
        tok.assign(f.p + i, j - i);
        char* ep = nullptr;
        const double v = strtod(tok.c_str(), &ep);
        if (ep == tok.c_str()) {
          errs[ti] = "not a number at row " + std::to_string(r) + " column " + std::to_string(c) + ": '" + tok + "'";
          return;
        }
        if (is_f64) static_cast<double*>(out)[r * ld + c] = v;
        else static_cast<float*>(out)[r * ld + c] = static_cast<float>(v);
        const bool last = c + 1 == n_cols;
        if (!last && (j >= end || f.p[j] != ',')) {
          errs[ti] = "row " + std::to_string(r) + " has fewer than " + std::to_string(n_cols) + " fields";
          return;
        }
        if (last && j < end && f.p[j] == ',') {
          errs[ti] = "row " + std::to_string(r) + " has more than " + std::to_string(n_cols) + " fields";
          return;
        }
        i = j + 1;
      }
    }
  };
  std::vector<std::thread> th;
  for (int ti = 1; ti < nt; ++ti) th.emplace_back(work, ti);
  work(0);
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) {
      g_io_err = e;
      return KS_ERR_INVALID;
    }
  return KS_OK;
}
KS_API int32_t ks_csv_read_f64(const char* path, double* out, int64_t n_rows, int64_t n_cols, int64_t ld) {
  return csv_read(path, out, n_rows, n_cols, ld, true);
}
KS_API int32_t ks_csv_read_f32(const char* path, float* out, int64_t n_rows, int64_t n_cols, int64_t ld) {
  return csv_read(path, out, n_rows, n_cols, ld, false);
}

// "row label" per line, both 1-based (TimitFeaturesDataLoader.scala:26-42): labels_out[row - 1] = label - 1; rows that the file
// does not mention keep -1 (the reference would fail on them at lookup time).
KS_API int32_t ks_timit_labels_read(const char* path, int32_t* labels_out, int64_t n_rows) {
  if (!path || !labels_out || n_rows < 0) return KS_ERR_INVALID;
  MappedFile f;
  if (!f.open_ro(path)) return KS_ERR_INVALID;
  for (int64_t i = 0; i < n_rows; ++i) labels_out[i] = -1;
  auto ls = line_starts(f.p, f.n);
  for (size_t li = 0; li < ls.size(); ++li) {
    const size_t end = (li + 1 < ls.size()) ? ls[li + 1] : f.n;
    std::string line(f.p + ls[li], end - ls[li]);
    char* ep = nullptr;
    const long long row = strtoll(line.c_str(), &ep, 10);
    if (ep == line.c_str()) {
      g_io_err = "bad label line " + std::to_string(li + 1);
      return KS_ERR_INVALID;
    }
    const char* q = ep;
    const long long lab = strtoll(q, &ep, 10);
    if (ep == q) {
      g_io_err = "bad label line " + std::to_string(li + 1);
      return KS_ERR_INVALID;
    }
    if (row < 1 || row > n_rows) {
      g_io_err = "label line " + std::to_string(li + 1) + ": row " + std::to_string(row) + " out of range";
      return KS_ERR_INVALID;
    }
    labels_out[row - 1] = static_cast<int32_t>(lab - 1);
  }
  return KS_OK;
}

// CIFAR-10 binary: records of 1 label byte + 3072 image bytes (channel planes R, G, B of 32 x 32, row-major inside a plane:
// RowColumnMajorByteArrayVectorizedImage, CifarLoader.scala:20-28).  images_out: [n][3072] bytes as stored; *n_out = records.
KS_API int32_t ks_cifar_read(const char* path, uint8_t* images_out, int32_t* labels_out, int64_t max_records, int64_t* n_out) {
  if (!path || !n_out) return KS_ERR_INVALID;
  MappedFile f;
  if (!f.open_ro(path)) return KS_ERR_INVALID;
  const size_t rec = 1 + 3072;
  if (f.n % rec != 0) {
    g_io_err = "file size is not a multiple of 3073 bytes";
    return KS_ERR_INVALID;
  }
  const int64_t n = static_cast<int64_t>(f.n / rec);
  *n_out = n;
  if (!images_out && !labels_out) return KS_OK;  // size query
  if (n > max_records) {
    g_io_err = "buffer holds " + std::to_string(max_records) + " records, file has " + std::to_string(n);
    return KS_ERR_INVALID;
  }
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t* r = reinterpret_cast<const uint8_t*>(f.p) + i * rec;
    if (labels_out) labels_out[i] = r[0];
    if (images_out) memcpy(images_out + i * 3072, r + 1, 3072);
  }
  return KS_OK;
}

}  // extern "C"
