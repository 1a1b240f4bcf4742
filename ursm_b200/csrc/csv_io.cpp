// Fast headerless comma-separated I/O for the scUnif.py file contract
// (scUnif.py:85-94: np.loadtxt(delimiter=",") in, np.savetxt(delimiter=",") out, i.e.
// '%.18e' per value, '\n' per row).  SURVEY 8f row 1: at BASELINE.json's sizes the text
// files are tens of GB and np.savetxt (one Python-level format call per row) dominates
// the wall time of the drop-in CLI.  Byte-exact with numpy: values are printed with the C
// library's correctly rounded "%.18e" (the same digits Python's % operator produces) and
// parsed with strtod (correctly rounded, as numpy's parser is).  Rows are split over
// std::threads.  Host-only code; it lives in the same C-ABI library for convenience.
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ursm_b200.h"

namespace ursm {
void set_error(const std::string& msg);
}

namespace {

int pick_threads(int requested, int64_t rows) {
  int n = requested > 0 ? requested : (int)std::thread::hardware_concurrency();
  if (n < 1) n = 1;
  if (n > 64) n = 64;
  if ((int64_t)n > rows) n = (int)std::max<int64_t>(1, rows);
  return n;
}

int fail(const std::string& msg) {
  ursm::set_error(msg);
  return 1;
}

// The matrices of this file contract have few distinct values (counts; E[S] = k/sample),
// while a correctly rounded 19-digit conversion costs about a microsecond in either
// direction.  Each worker thread therefore keeps a small direct-mapped cache
// value <-> text; a miss falls back to snprintf / strtod, so the bytes and the doubles
// are exactly those of the C library.
constexpr int kCacheBits = 13;
constexpr int kCacheSize = 1 << kCacheBits;
constexpr int kTokMax = 31;

struct FmtCache {  // bits of a double -> "%.18e" text
  struct Entry {
    uint64_t key;
    uint8_t len;
    char txt[kTokMax];
  };
  std::vector<Entry> e;
  FmtCache() : e(kCacheSize) {
    for (auto& x : e) {
      x.key = 0;
      x.len = 0;
    }
  }
  inline char* put(char* p, double v) {
    uint64_t k;
    memcpy(&k, &v, 8);
    Entry& x = e[(size_t)((k * 0x9E3779B97F4A7C15ull) >> (64 - kCacheBits))];
    if (x.len == 0 || x.key != k) {
      x.len = (uint8_t)snprintf(x.txt, kTokMax, "%.18e", v);
      x.key = k;
    }
    memcpy(p, x.txt, x.len);
    return p + x.len;
  }
};

struct ParseCache {  // text of a field -> double
  struct Entry {
    double val;
    uint8_t len;
    char txt[kTokMax];
  };
  std::vector<Entry> e;
  ParseCache() : e(kCacheSize) {
    for (auto& x : e) x.len = 0;
  }
  // parses the field starting at p (NUL-terminated buffer); returns the end of the number
  inline const char* get(const char* p, double* out) {
    size_t n = 0;
    uint64_t h = 1469598103934665603ull;
    while (p[n] != ',' && p[n] != '\n' && p[n] != '\0' && p[n] != ' ' && p[n] != '\r' &&
           n < (size_t)kTokMax) {
      h = (h ^ (uint8_t)p[n]) * 1099511628211ull;
      ++n;
    }
    if (n == 0 || n >= (size_t)kTokMax) {  // empty or unusually long field: no caching
      char* end = nullptr;
      *out = strtod(p, &end);
      return end;
    }
    Entry& x = e[(size_t)(h >> (64 - kCacheBits))];
    if (x.len == n && memcmp(x.txt, p, n) == 0) {
      *out = x.val;
      return p + n;
    }
    char* end = nullptr;
    const double v = strtod(p, &end);
    if (end == p + n) {  // the whole field is one number: remember it
      x.len = (uint8_t)n;
      memcpy(x.txt, p, n);
      x.val = v;
    }
    *out = v;
    return end;
  }
};

}  // namespace

extern "C" {

int ursm_csv_write_f64(const char* path, const double* data, int64_t rows, int64_t cols,
                       int32_t nthreads) {
  if (!path || (!data && rows * cols > 0) || rows < 0 || cols < 0)
    return fail("ursm_csv_write_f64: bad argument");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(std::string("ursm_csv_write_f64: cannot open ") + path + ": " + strerror(errno));
  const int nt = pick_threads(nthreads, rows);
  // '%.18e' is at most 26 characters ("-1.234567890123456789e+308") + separator
  const size_t per_value = 27;
  const int64_t rows_per_batch = std::max<int64_t>(1, (int64_t)((64u << 20) / (per_value * std::max<int64_t>(1, cols))));
  std::vector<std::string> bufs(nt);
  std::vector<FmtCache> caches(nt);
  bool ok = true;
  for (int64_t r0 = 0; r0 < rows && ok; r0 += rows_per_batch * nt) {
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
      const int64_t a = std::min(rows, r0 + rows_per_batch * t), b = std::min(rows, a + rows_per_batch);
      bufs[t].clear();
      if (a >= b) continue;
      th.emplace_back([&, t, a, b]() {
        std::string& s = bufs[t];
        s.resize((size_t)(b - a) * (size_t)cols * per_value + 1);
        char* p = &s[0];
        FmtCache& cache = caches[t];
        for (int64_t r = a; r < b; ++r) {
          const double* row = data + (size_t)r * cols;
          for (int64_t c = 0; c < cols; ++c) {
            p = cache.put(p, row[c]);
            *p++ = (c + 1 < cols) ? ',' : '\n';
          }
          if (cols == 0) *p++ = '\n';
        }
        s.resize((size_t)(p - s.data()));
      });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < nt && ok; ++t)
      if (!bufs[t].empty()) ok = fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
  }
  if (fclose(f) != 0) ok = false;
  if (!ok) return fail(std::string("ursm_csv_write_f64: write failed: ") + path);
  return 0;
}

// Shape of a headerless comma CSV: rows = non-empty lines, cols = fields of the first one.
int ursm_csv_shape(const char* path, int64_t* rows, int64_t* cols) {
  if (!path || !rows || !cols) return fail("ursm_csv_shape: bad argument");
  int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(std::string("ursm_csv_shape: cannot open ") + path + ": " + strerror(errno));
  struct stat st;
  fstat(fd, &st);
  const size_t n = (size_t)st.st_size;
  *rows = *cols = 0;
  if (n == 0) {
    close(fd);
    return 0;
  }
  const char* m = (const char*)mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return fail("ursm_csv_shape: mmap failed");
  int64_t r = 0, c = 1;
  size_t line = 0;
  while (line < n) {
    const char* nl = (const char*)memchr(m + line, '\n', n - line);
    const size_t e = nl ? (size_t)(nl - m) : n;
    bool nonempty = false;
    for (size_t i = line; i < e && !nonempty; ++i) nonempty = (m[i] != '\r' && m[i] != ' ');
    if (nonempty) {
      if (r == 0)
        for (size_t i = line; i < e; ++i) c += (m[i] == ',');
      ++r;
    }
    line = e + 1;
  }
  munmap((void*)m, n);
  *rows = r;
  *cols = r ? c : 0;
  return 0;
}

int ursm_csv_read_f64(const char* path, double* out, int64_t rows, int64_t cols, int32_t nthreads) {
  if (!path || !out || rows < 0 || cols < 0) return fail("ursm_csv_read_f64: bad argument");
  if (rows == 0 || cols == 0) return 0;
  int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(std::string("ursm_csv_read_f64: cannot open ") + path + ": " + strerror(errno));
  struct stat st;
  fstat(fd, &st);
  const size_t n = (size_t)st.st_size;
  // one extra zero page-safe byte: copy the tail so strtod never runs off the mapping
  const char* m = (const char*)mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return fail("ursm_csv_read_f64: mmap failed");
  // start offsets of the non-empty lines
  std::vector<size_t> starts;
  starts.reserve((size_t)rows + 1);
  {
    size_t line = 0;
    while (line < n) {
      const char* nl = (const char*)memchr(m + line, '\n', n - line);
      const size_t e = nl ? (size_t)(nl - m) : n;
      bool nonempty = false;
      for (size_t i = line; i < e && !nonempty; ++i) nonempty = (m[i] != '\r' && m[i] != ' ');
      if (nonempty) starts.push_back(line);
      line = e + 1;
    }
  }
  if ((int64_t)starts.size() != rows) {
    munmap((void*)m, n);
    return fail("ursm_csv_read_f64: row count changed since ursm_csv_shape");
  }
  const int nt = pick_threads(nthreads, rows);
  std::vector<int64_t> bad(nt, -1);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) {
    const int64_t a = rows * t / nt, b = rows * (t + 1) / nt;
    th.emplace_back([&, t, a, b]() {
      std::string last;
      ParseCache cache;
      for (int64_t r = a; r < b; ++r) {
        // fields are parsed in place: every line but the last one of the file ends in
        // '\n' inside the mapping, which stops strtod; the last line gets a private
        // NUL-terminated copy
        const char* p = m + starts[r];
        if (r == rows - 1) {
          last.assign(p, n - starts[r]);
          p = last.c_str();
        }
        double* dst = out + (size_t)r * cols;
        for (int64_t c = 0; c < cols; ++c) {
          while (*p == ' ') ++p;
          const char* end = cache.get(p, &dst[c]);
          if (end == p) {
            bad[t] = r;
            return;
          }
          p = end;
          while (*p == ' ' || *p == '\r') ++p;
          if (c + 1 < cols) {
            if (*p != ',') {
              bad[t] = r;
              return;
            }
            ++p;
          } else if (*p != '\0' && *p != '\n') {
            bad[t] = r;
            return;
          }
        }
      }
    });
  }
  for (auto& x : th) x.join();
  munmap((void*)m, n);
  for (int t = 0; t < nt; ++t)
    if (bad[t] >= 0) {
      char msg[256];
      snprintf(msg, sizeof msg, "ursm_csv_read_f64: %s: malformed or ragged row %lld", path,
               (long long)bad[t]);
      return fail(msg);
    }
  return 0;
}

}  // extern "C"
