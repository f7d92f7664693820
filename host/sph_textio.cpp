// sph_textio.cpp — host-parallel reader / writer for the reference's text formats (include/sph_textio.h).
//
// Reader (read_data_from_file, SUMMER_SPH.f90:594-716 | "SUMMER_SPH - Variable.f90":729-852): the file is mapped,
// the header line skipped, the rest cut into one line-aligned piece per thread; every thread parses its rows with
// std::from_chars into its own columns, and the pieces are concatenated in file order (so gas rows keep their
// `number` order, F:684).  Writer (make_save, F:719-738 | V:921-942): rows have a fixed width, so every thread
// formats a contiguous block of rows and writes it at its own offset.
#include "../include/sph_textio.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

int thread_count(int32_t threads, size_t work_items) {
  int t = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (t < 1) t = 1;
  if ((size_t)t > work_items) t = (int)std::max<size_t>(1, work_items);
  return t;
}

inline bool is_sep(char c) { return c == ' ' || c == '\t' || c == ',' || c == '\r'; }

// One real in Fortran list-directed spelling: optional sign, digits, optional exponent with e/E/d/D.
// Returns the end of the token, or nullptr when the token is not a number.
const char* parse_real(const char* p, const char* end, double& v) {
  const char* q = p;
  if (q < end && *q == '+') ++q;                       // from_chars takes '-' but not '+'
  const char* tok_end = q;
  bool has_d = false;
  while (tok_end < end && !is_sep(*tok_end) && *tok_end != '\n') { if (*tok_end == 'd' || *tok_end == 'D') has_d = true; ++tok_end; }
  if (tok_end == q) return nullptr;
  if (!has_d) {
    auto r = std::from_chars(q, tok_end, v);
    return (r.ec == std::errc() && r.ptr == tok_end) ? tok_end : nullptr;
  }
  char buf[64];
  const size_t len = (size_t)(tok_end - q);
  if (len >= sizeof(buf)) return nullptr;
  for (size_t i = 0; i < len; ++i) buf[i] = (q[i] == 'd' || q[i] == 'D') ? 'e' : q[i];
  auto r = std::from_chars(buf, buf + len, v);
  return (r.ec == std::errc() && r.ptr == buf + len) ? tok_end : nullptr;
}

struct Piece {
  std::vector<double> gas[10], sink[8];
  int64_t rows = 0;          // non-blank data rows seen
  int64_t bad_row = -1;      // first row (0-based within the piece) that could not be read
};

void parse_piece(const char* p, const char* end, bool variable, double h_fixed, double sink_radius, Piece& out) {
  const int want = variable ? 10 : 8;
  while (p < end) {
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    double v[10]; int k = 0; bool ok = true;
    const char* q = p;
    while (q < eol) {
      while (q < eol && is_sep(*q)) ++q;
      if (q >= eol) break;
      if (k == 10) break;                                   // extra columns are ignored (F:647)
      const char* r = parse_real(q, eol, v[k]);
      if (!r) { ok = false; break; }
      ++k; q = r;
    }
    if (k > 0 || !ok) {                                     // blank rows are skipped
      const int64_t row = out.rows++;
      const bool sink = ok && k >= 8 && v[6] == 0.0;        // u == 0 exactly marks a sink (F:658-659)
      if (!ok || k < 8 || (!sink && k < want)) { if (out.bad_row < 0) out.bad_row = row; }
      else if (sink) {
        const double s[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[7], sink_radius};
        for (int c = 0; c < 8; ++c) out.sink[c].push_back(s[c]);
      } else {
        for (int c = 0; c < 8; ++c) out.gas[c].push_back(v[c]);
        out.gas[8].push_back(variable ? v[8] : 0.0);        // alpha := 0 when only 8 columns are read (F:681)
        out.gas[9].push_back(variable ? v[9] : h_fixed);
      }
    }
    p = eol < end ? eol + 1 : end;
  }
}

// "%25.17E" of one value into exactly 25 characters (right aligned)
inline void put_real(char* dst, double v) {
  char tmp[40];
  int len;
  if (std::isfinite(v)) {
    auto r = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::scientific, 17);
    len = (int)(r.ptr - tmp);
    for (int i = 0; i < len; ++i) if (tmp[i] == 'e') tmp[i] = 'E';
  } else {
    const char* s = std::isnan(v) ? "NAN" : (v > 0 ? "INF" : "-INF");
    len = (int)strlen(s); memcpy(tmp, s, (size_t)len);
  }
  if (len > 25) len = 25;
  memset(dst, ' ', (size_t)(25 - len));
  memcpy(dst + (25 - len), tmp, (size_t)len);
}

}  // namespace

struct sph_ics {
  std::vector<double> gas[10], sink[8];
};

extern "C" {

const char* sph_textio_last_error(void) { return g_err.c_str(); }

int sph_ics_open(const char* path, int32_t variable_h, double h_fixed, double sink_radius, int32_t threads, sph_ics** out) {
  if (!path || !out) return fail(SPH_TEXTIO_ERR_ARG, "null argument");
  *out = nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(SPH_TEXTIO_ERR_OPEN, std::string("Error opening file: ") + path);
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); return fail(SPH_TEXTIO_ERR_OPEN, std::string("Error opening file: ") + path); }
  const size_t size = (size_t)st.st_size;
  const char* base = nullptr;
  if (size > 0) {
    base = (const char*)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (base == MAP_FAILED) { close(fd); return fail(SPH_TEXTIO_ERR_OPEN, std::string("Error opening file: ") + path); }
  }
  close(fd);
  const char* end = base + size;
  const char* body = size ? (const char*)memchr(base, '\n', size) : nullptr;     // header line (F:617)
  body = body ? body + 1 : end;
  const size_t len = (size_t)(end - body);
  const int T = thread_count(threads, len / (1 << 20) + 1);
  std::vector<Piece> pieces((size_t)T);
  std::vector<const char*> cut((size_t)T + 1);
  cut[0] = body; cut[(size_t)T] = end;
  for (int t = 1; t < T; ++t) {
    const char* c = body + len * (size_t)t / (size_t)T;
    if (c < cut[(size_t)t - 1]) c = cut[(size_t)t - 1];
    const char* nl = (const char*)memchr(c, '\n', (size_t)(end - c));
    cut[(size_t)t] = nl ? nl + 1 : end;
  }
  {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t]() { parse_piece(cut[(size_t)t], cut[(size_t)t + 1], variable_h != 0, h_fixed, sink_radius, pieces[(size_t)t]); });
    for (auto& x : th) x.join();
  }
  if (base) munmap((void*)base, size);
  int64_t rows = 0;
  for (int t = 0; t < T; ++t) {
    if (pieces[(size_t)t].bad_row >= 0) return fail(SPH_TEXTIO_ERR_PARSE, "Error reading line " + std::to_string(rows + pieces[(size_t)t].bad_row + 1));
    rows += pieces[(size_t)t].rows;
  }
  if (rows == 0) return fail(SPH_TEXTIO_ERR_EMPTY, std::string("No data found in file: ") + path);
  sph_ics* ics = new sph_ics();
  size_t ng = 0, ns = 0;
  std::vector<size_t> og((size_t)T), os((size_t)T);
  for (int t = 0; t < T; ++t) { og[(size_t)t] = ng; os[(size_t)t] = ns; ng += pieces[(size_t)t].gas[0].size(); ns += pieces[(size_t)t].sink[0].size(); }
  for (auto& v : ics->gas) v.resize(ng);
  for (auto& v : ics->sink) v.resize(ns);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t]() {
        const Piece& p = pieces[(size_t)t];
        for (int c = 0; c < 10; ++c) if (!p.gas[c].empty()) memcpy(ics->gas[c].data() + og[(size_t)t], p.gas[c].data(), p.gas[c].size() * sizeof(double));
        for (int c = 0; c < 8; ++c) if (!p.sink[c].empty()) memcpy(ics->sink[c].data() + os[(size_t)t], p.sink[c].data(), p.sink[c].size() * sizeof(double));
      });
    for (auto& x : th) x.join();
  }
  if (ns == 0) for (auto& v : ics->sink) v.assign(1, 0.0);     // dummy sink: all zeros, radius 0 (F:698-707)
  *out = ics;
  return SPH_TEXTIO_OK;
}

int sph_ics_sizes(const sph_ics* ics, int64_t* n_gas, int32_t* n_sink) {
  if (!ics) return fail(SPH_TEXTIO_ERR_ARG, "null handle");
  if (n_gas) *n_gas = (int64_t)ics->gas[0].size();
  if (n_sink) *n_sink = (int32_t)ics->sink[0].size();
  return SPH_TEXTIO_OK;
}

int sph_ics_fetch(const sph_ics* ics,
                  double* x, double* y, double* z, double* vx, double* vy, double* vz, double* u, double* m, double* alpha, double* h,
                  double* sx, double* sy, double* sz, double* svx, double* svy, double* svz, double* sm, double* sradius) {
  if (!ics) return fail(SPH_TEXTIO_ERR_ARG, "null handle");
  double* g[10] = {x, y, z, vx, vy, vz, u, m, alpha, h};
  double* s[8] = {sx, sy, sz, svx, svy, svz, sm, sradius};
  for (int c = 0; c < 10; ++c) if (g[c] && !ics->gas[c].empty()) memcpy(g[c], ics->gas[c].data(), ics->gas[c].size() * sizeof(double));
  for (int c = 0; c < 8; ++c) if (s[c] && !ics->sink[c].empty()) memcpy(s[c], ics->sink[c].data(), ics->sink[c].size() * sizeof(double));
  return SPH_TEXTIO_OK;
}

int sph_ics_close(sph_ics* ics) { delete ics; return SPH_TEXTIO_OK; }

int sph_save_write(const char* path, int32_t variable_h, int64_t n_gas,
                   const double* x, const double* y, const double* z, const double* vx, const double* vy, const double* vz,
                   const double* u, const double* m, const double* alpha, const double* h,
                   int32_t n_sink, const double* sx, const double* sy, const double* sz, const double* svx, const double* svy,
                   const double* svz, const double* sm, int32_t threads) {
  if (!path || n_gas < 0 || n_sink < 0) return fail(SPH_TEXTIO_ERR_ARG, "bad argument");
  const double* g[10] = {x, y, z, vx, vy, vz, u, m, alpha, h};
  const double* s[7] = {sx, sy, sz, svx, svy, svz, sm};
  const int ncol = variable_h ? 10 : 9;
  if (n_gas > 0) for (int c = 0; c < ncol; ++c) if (!g[c]) return fail(SPH_TEXTIO_ERR_ARG, "null gas column");
  if (n_sink > 0) for (int c = 0; c < 7; ++c) if (!s[c]) return fail(SPH_TEXTIO_ERR_ARG, "null sink column");
  const int fd = open(path, O_WRONLY | O_CREAT | O_EXCL, 0644);                  // status="new" (F:728)
  if (fd < 0) return fail(SPH_TEXTIO_ERR_EXISTS, std::string(path) + " exists or cannot be created (status=\"new\")");
  std::string hdr = " x  y  z  vx  vy vz energy mass  alpha";                    // F:729 | V:931
  if (variable_h) hdr += "  smoothing";
  hdr += "\n";
  bool ok = write(fd, hdr.data(), hdr.size()) == (ssize_t)hdr.size();
  const size_t grow = (size_t)ncol * 26, srow = 8 * 26;        // 25 characters per value, one blank between values, newline
  const size_t gas_off = hdr.size(), sink_off = gas_off + (size_t)n_gas * grow;
  const int T = thread_count(threads, (size_t)n_gas / 4096 + 1);
  std::vector<char> oks((size_t)T, 1);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t]() {
        const int64_t r0 = n_gas * t / T, r1 = n_gas * (t + 1) / T;
        const int64_t block = 8192;
        std::vector<char> buf((size_t)block * grow);
        for (int64_t a = r0; a < r1; a += block) {
          const int64_t b = std::min(r1, a + block);
          char* w = buf.data();
          for (int64_t i = a; i < b; ++i) {
            for (int c = 0; c < ncol; ++c) { put_real(w, g[c][i]); w += 25; *w++ = (c + 1 < ncol) ? ' ' : '\n'; }
          }
          const size_t bytes = (size_t)(w - buf.data());
          if (pwrite(fd, buf.data(), bytes, (off_t)(gas_off + (size_t)a * grow)) != (ssize_t)bytes) oks[(size_t)t] = 0;
        }
      });
    for (auto& xth : th) xth.join();
  }
  for (char c : oks) ok = ok && c;
  if (n_sink > 0) {
    std::vector<char> buf((size_t)n_sink * srow);
    char* w = buf.data();
    for (int32_t i = 0; i < n_sink; ++i) {
      const double row[8] = {s[0][i], s[1][i], s[2][i], s[3][i], s[4][i], s[5][i], 0.0, s[6][i]};     // x y z vx vy vz 0.0 m
      for (int c = 0; c < 8; ++c) { put_real(w, row[c]); w += 25; *w++ = (c + 1 < 8) ? ' ' : '\n'; }
    }
    ok = ok && pwrite(fd, buf.data(), buf.size(), (off_t)sink_off) == (ssize_t)buf.size();
  }
  ok = (close(fd) == 0) && ok;
  return ok ? SPH_TEXTIO_OK : fail(SPH_TEXTIO_ERR_WRITE, std::string("short write to ") + path);
}

}  // extern "C"
