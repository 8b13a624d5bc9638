// csv_writer.cpp — the on-disk format of the hot path's output: `Writer::write` of the reference (writer.rs:99-146).
// Host code (file I/O), part of the C-ABI library.  One file per call, named <prefix><id>.csv; first line the variable
// names joined by ','; then the values in chunks of n_vars per line, each printed like Rust's `f64::to_string()`:
// shortest digits that round-trip, positional notation always (no exponent), integral values without ".0",
// "NaN" / "inf" / "-inf", and "-0" for negative zero.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/dzahui_b200.h"

namespace dzb {

std::string rust_f64_to_string(double v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
  if (v == 0.0) return std::signbit(v) ? "-0" : "0";
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);  // shortest round-trip digits
  std::string sci(buf, res.ptr);
  std::string out;
  size_t pos = 0;
  if (sci[0] == '-') out.push_back('-'), pos = 1;
  const size_t epos = sci.find('e');
  std::string digits;
  for (size_t i = pos; i < epos; ++i)
    if (sci[i] != '.') digits.push_back(sci[i]);
  const int exp10 = std::atoi(sci.c_str() + epos + 1);  // value = d.ddd x 10^exp10
  const int nd = (int)digits.size();
  if (exp10 >= 0) {
    if (exp10 + 1 >= nd) {  // integer: pad with zeros
      out += digits;
      out.append((size_t)(exp10 + 1 - nd), '0');
    } else {
      out.append(digits, 0, (size_t)exp10 + 1);
      out.push_back('.');
      out.append(digits, (size_t)exp10 + 1, std::string::npos);
    }
  } else {
    out += "0.";
    out.append((size_t)(-exp10 - 1), '0');
    out += digits;
  }
  return out;
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
  for (size_t i = 0; i < n_vals; ++i) {
    block += dzb::rust_f64_to_string(vals[i]);
    block.push_back(((i + 1) % n_vars == 0 || i + 1 == n_vals) ? '\n' : ',');
    if (block.size() > (1u << 20)) {
      if (std::fwrite(block.data(), 1, block.size(), f) != block.size()) { std::fclose(f); return DZB_IO; }
      block.clear();
    }
  }
  const bool ok = std::fwrite(block.data(), 1, block.size(), f) == block.size();
  return (std::fclose(f) == 0 && ok) ? DZB_OK : DZB_IO;
}

}  // extern "C"
