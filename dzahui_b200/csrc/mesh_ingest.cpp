// mesh_ingest.cpp — see mesh_ingest.h.
#include "mesh_ingest.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <unordered_set>
#include <utility>

namespace dzb {
namespace {

// Accepts exactly what Rust's `str::parse::<f64>()` accepts: optional sign, then inf/infinity/nan (any case) or
// decimal digits with an optional fraction and exponent, at least one mantissa digit, no surrounding blanks.
// The value is correctly rounded (glibc strtod), like Rust's dec2flt.
bool parse_rust_float(const char* b, const char* e, double* value) {
  const char* p = b;
  if (p == e) return false;
  if (*p == '+' || *p == '-') ++p;
  if (p == e) return false;
  auto word = [&](const char* wd) {
    const char* q = p;
    for (; *wd; ++wd, ++q)
      if (q == e || std::tolower((unsigned char)*q) != *wd) return false;
    return q == e;
  };
  bool special = word("inf") || word("infinity") || word("nan");
  if (!special) {
    int mant = 0;
    while (p != e && std::isdigit((unsigned char)*p)) ++p, ++mant;
    if (p != e && *p == '.') {
      ++p;
      while (p != e && std::isdigit((unsigned char)*p)) ++p, ++mant;
    }
    if (!mant) return false;
    if (p != e && (*p == 'e' || *p == 'E')) {
      ++p;
      if (p != e && (*p == '+' || *p == '-')) ++p;
      int ed = 0;
      while (p != e && std::isdigit((unsigned char)*p)) ++p, ++ed;
      if (!ed) return false;
    }
    if (p != e) return false;
  }
  std::string tmp(b, e);
  *value = std::strtod(tmp.c_str(), nullptr);
  return true;
}

struct VertexLine {
  std::string token[3];
  double value[3];
};

}  // namespace

IngestStatus ingest_obj_1d(const char* path, bool has_multiplier, double multiplier, Ingested1D* out) {
  std::FILE* f = std::fopen(path, "rb");
  if (!f) {
    out->message = std::string("cannot open ") + path;
    return IngestStatus::Io;
  }
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
  std::fclose(f);

  // One pass over the lines: keep those starting with "v " and split them on single blanks (`split(" ")`),
  // exactly 3 fields after the tag, each a Rust-parsable float (mesh_builder.rs:58-84 and :159-178).
  std::vector<VertexLine> vs;
  size_t pos = 0;
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    size_t end = eol;
    if (end > pos && text[end - 1] == '\r') --end;  // BufRead::lines strips "\r\n"
    if (end - pos >= 2 && text[pos] == 'v' && text[pos + 1] == ' ') {
      VertexLine vl;
      int nf = 0;
      size_t s = pos + 2;
      bool ok = true;
      for (;;) {
        size_t sp = text.find(' ', s);
        if (sp == std::string::npos || sp > end) sp = end;
        double v;
        if (!parse_rust_float(text.data() + s, text.data() + sp, &v)) ok = false;
        if (nf < 3) {
          vl.token[nf].assign(text, s, sp - s);
          vl.value[nf] = v;
        }
        ++nf;
        if (sp == end) break;
        s = sp + 1;
      }
      if (!ok) {
        out->message = "Error while parsing vertex coordinate from obj";
        return IngestStatus::Parse;
      }
      if (nf != 3) {
        out->message = "A vertex line should contain 3 elements only";
        return IngestStatus::Parse;
      }
      vs.push_back(std::move(vl));
    }
    pos = eol + 1;
  }

  // Axis detection (check_for_constant_coordinates, :140-190): distinct coordinate STRINGS per axis, every key
  // starting with "0.0" or "-0.0" collapsed to "0.0"; an axis with one key is constant.
  std::unordered_set<std::string> keys[3];
  for (const VertexLine& vl : vs)
    for (int a = 0; a < 3; ++a) {
      const std::string& t = vl.token[a];
      if (t.compare(0, 3, "0.0") == 0 || t.compare(0, 4, "-0.0") == 0)
        keys[a].insert("0.0");
      else
        keys[a].insert(t);
    }
  const bool cx = keys[0].size() == 1, cy = keys[1].size() == 1, cz = keys[2].size() == 1;
  int kept;  // the pair [1,0] / [2,1] / [2,0] of removed axes leaves z / x / y (:217-227, :244-247)
  if (cx && cy)
    kept = 2;
  else if (cy && cz)
    kept = 0;
  else if (cz && cx)
    kept = 1;
  else {
    out->message = "Only coordinates over a line paralell to x, y or z axis are accepted. Check .obj file.";
    return IngestStatus::Parse;
  }
  if (vs.empty()) {  // unreachable above (no keys => size 0), kept for clarity: the reference would panic
    out->message = "no vertex lines";
    return IngestStatus::Parse;
  }
  out->kept_axis = kept;

  // Ordering (:253-260): insertion sort that shifts while `stored > new`; i.e. a stable ascending sort in which
  // equal keys (-0.0 / +0.0) keep file order.  NaNs break strict weak ordering, so replay the literal insertion.
  std::vector<double>& x = out->nodes;
  x.clear();
  x.reserve(vs.size());
  bool has_nan = false;
  for (const VertexLine& vl : vs) {
    x.push_back(vl.value[kept]);
    has_nan |= std::isnan(vl.value[kept]);
  }
  if (!has_nan) {
    if (!std::is_sorted(x.begin(), x.end())) std::stable_sort(x.begin(), x.end());
  } else {
    std::vector<double> sorted;
    for (double v : x) {
      sorted.push_back(v);
      long long j = (long long)sorted.size() - 2;
      while (j >= 0 && sorted[(size_t)j] > v) {
        sorted[(size_t)j + 1] = sorted[(size_t)j];
        --j;
      }
      sorted[(size_t)(j + 1)] = v;
    }
    x.swap(sorted);
  }

  const double vertices_len = 6.0 * (double)x.size();
  out->max_length = -x.front() + x.back();
  out->prom_width = (out->max_length * 6.0 / (vertices_len - 6.)) * (has_multiplier ? multiplier : 1.0);
  const float mid0 = (float)x.front() + (float)out->max_length / 2.0f;
  const float mid1 = (float)out->prom_width / 2.0f;
  out->translation[0] = -mid0;
  out->translation[1] = -mid1;
  out->translation[2] = 0.0f;
  return IngestStatus::Ok;
}

IngestStatus ingest_obj_2d(const char* path, Ingested2D* out) {
  std::FILE* f = std::fopen(path, "rb");
  if (!f) {
    out->message = std::string("cannot open ") + path;
    return IngestStatus::Io;
  }
  std::string text;
  char buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
  std::fclose(f);

  std::vector<VertexLine> vs;
  std::unordered_set<std::string> keys[3];
  size_t pos = 0;
  auto fields = [&](size_t s, size_t end, std::vector<std::pair<size_t, size_t>>* out_f) {  // str::split(" ") after the tag
    out_f->clear();
    for (;;) {
      size_t sp = text.find(' ', s);
      if (sp == std::string::npos || sp > end) sp = end;
      out_f->push_back({s, sp});
      if (sp == end) break;
      s = sp + 1;
    }
  };
  std::vector<std::pair<size_t, size_t>> fl;
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    size_t end = eol;
    if (end > pos && text[end - 1] == '\r') --end;
    if (end - pos >= 2 && text[pos] == 'v' && text[pos + 1] == ' ') {
      fields(pos + 2, end, &fl);
      VertexLine vl;
      bool ok = true;
      for (size_t k = 0; k < fl.size(); ++k) {
        double v;
        if (!parse_rust_float(text.data() + fl[k].first, text.data() + fl[k].second, &v)) ok = false;
        if (k < 3) vl.token[k].assign(text, fl[k].first, fl[k].second - fl[k].first), vl.value[k] = v;
      }
      if (!ok) {
        out->message = "Error while parsing vertex coordinate from obj";
        return IngestStatus::Parse;
      }
      if (fl.size() != 3) {
        out->message = "A vertex line should contain 3 elements only";
        return IngestStatus::Parse;
      }
      for (int a = 0; a < 3; ++a) {
        const std::string& t = vl.token[a];
        keys[a].insert((t.compare(0, 3, "0.0") == 0 || t.compare(0, 4, "-0.0") == 0) ? std::string("0.0") : t);
      }
      vs.push_back(std::move(vl));
    } else if (end - pos >= 2 && text[pos] == 'f' && text[pos + 1] == ' ') {
      fields(pos + 2, end, &fl);
      if (fl.size() != 3) {
        out->message = "Amount of face specificating elements should be 3.";
        return IngestStatus::Parse;
      }
      for (size_t k = 0; k < 3; ++k) {  // a/b/c: the first field is the vertex, exactly two more must follow
        const size_t b = fl[k].first, e = fl[k].second;
        size_t slash = text.find('/', b);
        if (slash == std::string::npos || slash > e) slash = e;
        unsigned long long v = 0;
        size_t i = b;
        if (i < slash && text[i] == '+') ++i;
        bool ok = i < slash;
        for (; i < slash; ++i) {
          if (text[i] < '0' || text[i] > '9') { ok = false; break; }
          v = v * 10 + (unsigned)(text[i] - '0');
          if (v > 0xffffffffULL) { ok = false; break; }
        }
        if (!ok) {
          out->message = "Error while parsing face coordinate";
          return IngestStatus::Parse;
        }
        size_t extra = 0;
        for (size_t j = slash; j < e; ++j) extra += text[j] == '/';
        if (extra != 2) {
          out->message = "Amount of elements per face specification should be 3 in format a/b/c.";
          return IngestStatus::Parse;
        }
        out->triangles.push_back((unsigned)v - 1u);
      }
    }
    pos = eol + 1;
  }
  int cc;  // :355-363
  if (keys[0].size() == 1) cc = 0;
  else if (keys[1].size() == 1) cc = 1;
  else if (keys[2].size() == 1) cc = 2;
  else {
    out->message = "Only coordinates over a plane paralell to x, y or z plane are accepted. Check .obj file.";
    return IngestStatus::Parse;
  }
  out->constant_axis = cc;
  double x_min = 0.0, y_min = 0.0, x_max = 0.0, y_max = 0.0;  // all start at 0.0 (:366-371)
  out->xy.reserve(2 * vs.size());
  for (const VertexLine& vl : vs) {
    const double a = vl.value[cc == 0 ? 1 : 0], b = vl.value[cc == 2 ? 1 : 2];
    if (a < x_min) x_min = a;
    if (a > x_max) x_max = a;
    if (b < y_min) y_min = b;
    if (b > y_max) y_max = b;
    out->xy.push_back(a), out->xy.push_back(b);
  }
  const double len_x = x_max - x_min, len_y = y_max - y_min;
  out->max_length = len_x > len_y ? len_x : len_y;
  out->translation[0] = (float)x_min + (float)out->max_length / 2.0f;
  out->translation[1] = (float)y_min + (float)out->max_length / 2.0f;
  out->translation[2] = 0.0f;
  return IngestStatus::Ok;
}

}  // namespace dzb
