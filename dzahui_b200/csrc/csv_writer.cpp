// csv_writer.cpp — the on-disk format of the hot path's output: `Writer::write` of the reference (writer.rs:99-146).
// Host code (file I/O), part of the C-ABI library.  One file per call, named <prefix><id>.csv; first line the variable
// names joined by ','; then the values in chunks of n_vars per line, each printed like Rust's `f64::to_string()`:
// shortest digits that round-trip, positional notation always (no exponent), integral values without ".0",
// "NaN" / "inf" / "-inf", and "-0" for negative zero.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/dzahui_b200.h"

namespace dzb {

// Writes the digits into out (at least RUST_F64_MAX bytes) and returns their count; no allocation.
constexpr size_t RUST_F64_MAX = 352;  // "-0." + 323 zeros + 17 digits, or 309 integer digits
size_t rust_f64_format(double v, char* out) {
  auto lit = [&](const char* t) {
    const size_t n = std::strlen(t);
    std::memcpy(out, t, n);
    return n;
  };
  if (std::isnan(v)) return lit("NaN");
  if (std::isinf(v)) return lit(v < 0 ? "-inf" : "inf");
  if (v == 0.0) return lit(std::signbit(v) ? "-0" : "0");
  char sci[64];
  const char* end = std::to_chars(sci, sci + sizeof sci, v, std::chars_format::scientific).ptr;  // shortest round-trip digits
  const char* p = sci;
  char* o = out;
  if (*p == '-') *o++ = '-', ++p;
  char digits[24];
  int nd = 0;
  for (; p < end && *p != 'e'; ++p)
    if (*p != '.') digits[nd++] = *p;
  int exp10 = 0;  // value = d.ddd x 10^exp10
  {
    ++p;  // 'e'
    const bool neg = *p == '-';
    if (*p == '-' || *p == '+') ++p;
    for (; p < end; ++p) exp10 = exp10 * 10 + (*p - '0');
    if (neg) exp10 = -exp10;
  }
  if (exp10 >= 0) {
    if (exp10 + 1 >= nd) {  // integer: pad with zeros
      std::memcpy(o, digits, (size_t)nd), o += nd;
      std::memset(o, '0', (size_t)(exp10 + 1 - nd)), o += exp10 + 1 - nd;
    } else {
      std::memcpy(o, digits, (size_t)exp10 + 1), o += exp10 + 1;
      *o++ = '.';
      std::memcpy(o, digits + exp10 + 1, (size_t)(nd - exp10 - 1)), o += nd - exp10 - 1;
    }
  } else {
    *o++ = '0', *o++ = '.';
    std::memset(o, '0', (size_t)(-exp10 - 1)), o += -exp10 - 1;
    std::memcpy(o, digits, (size_t)nd), o += nd;
  }
  return (size_t)(o - out);
}
std::string rust_f64_to_string(double v) {
  char buf[RUST_F64_MAX];
  return std::string(buf, rust_f64_format(v, buf));
}

}  // namespace dzb

extern "C" {

int dzb_format_f64(double v, char* buf, size_t cap) {
  if (!buf) return DZB_INVALID;
  const std::string s = dzb::rust_f64_to_string(v);
  if (s.size() + 1 > cap) return DZB_WRONG_DIMS;
  std::memcpy(buf, s.c_str(), s.size() + 1);
  return DZB_OK;
}

// Writer::new (writer.rs:47-86): the directory must exist (Error::NotFound -> DZB_IO); erase_prev_dir removes the
// plain files directly inside it, never nested directories.
int dzb_csv_prepare_dir(const char* dir, int erase_prev_dir) {
  if (!dir) return DZB_INVALID;
  struct stat st;
  if (stat(dir, &st) != 0 || !S_ISDIR(st.st_mode)) return DZB_IO;
  if (!erase_prev_dir) return DZB_OK;
  DIR* d = opendir(dir);
  if (!d) return DZB_IO;
  int rc = DZB_OK;
  while (dirent* e = readdir(d)) {
    if (!std::strcmp(e->d_name, ".") || !std::strcmp(e->d_name, "..")) continue;
    const std::string p = std::string(dir) + "/" + e->d_name;
    struct stat fs;
    if (lstat(p.c_str(), &fs) != 0) { rc = DZB_IO; break; }
    if (!S_ISDIR(fs.st_mode) && unlink(p.c_str()) != 0) { rc = DZB_IO; break; }
  }
  closedir(d);
  return rc;
}

// Writer::write (writer.rs:99-146)
int dzb_csv_write(const char* dir, const char* prefix, double id, const char* const* variable_names, size_t n_vars,
                  const double* vals, size_t n_vals) {
  if (!dir || !prefix || !variable_names || (!vals && n_vals) || n_vars == 0) return DZB_INVALID;
  const std::string path = std::string(dir) + "/" + prefix + dzb::rust_f64_to_string(id) + ".csv";
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) return DZB_IO;
  std::string block;
  for (size_t k = 0; k < n_vars; ++k) {
    block += variable_names[k];
    block.push_back(k + 1 < n_vars ? ',' : '\n');
  }
  if (std::fwrite(block.data(), 1, block.size(), f) != block.size()) { std::fclose(f); return DZB_IO; }
  // the values: formatted on all host threads, slab by slab (each thread a contiguous run of a slab), written in order
  const size_t SLAB = (size_t)1 << 22;
  const size_t hw = std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::string> parts;
  bool ok = true;
  for (size_t s0 = 0; s0 < n_vals && ok; s0 += SLAB) {
    const size_t s1 = std::min(n_vals, s0 + SLAB);
    const size_t nt = std::min(std::min<size_t>(hw, 32), (s1 - s0) / 65536 + 1);
    parts.assign(nt, std::string());
    auto work = [&](size_t k) {
      const size_t b = s0 + (s1 - s0) * k / nt, e = s0 + (s1 - s0) * (k + 1) / nt;
      std::string& out = parts[k];
      out.reserve((e - b) * 24);
      char buf[dzb::RUST_F64_MAX + 1];
      for (size_t i = b; i < e; ++i) {
        size_t len = dzb::rust_f64_format(vals[i], buf);
        buf[len++] = ((i + 1) % n_vars == 0 || i + 1 == n_vals) ? '\n' : ',';
        out.append(buf, len);
      }
    };
    std::vector<std::thread> pool;
    for (size_t k = 1; k < nt; ++k) pool.emplace_back(work, k);
    work(0);
    for (std::thread& t : pool) t.join();
    for (const std::string& part : parts) ok = ok && std::fwrite(part.data(), 1, part.size(), f) == part.size();
  }
  return (std::fclose(f) == 0 && ok) ? DZB_OK : DZB_IO;
}

}  // extern "C"
