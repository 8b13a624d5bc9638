// mesh_ingest.cpp — see mesh_ingest.h.
#include "mesh_ingest.h"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string_view>
#include <thread>
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
  if (!special) {  // the common case without a heap copy: from_chars is correctly rounded too and takes a bounded range
    const char* q = *b == '+' ? b + 1 : b;
    const std::from_chars_result r = std::from_chars(q, e, *value, std::chars_format::general);
    if (r.ec == std::errc() && r.ptr == e) return true;
  }
  std::string tmp(b, e);  // inf / nan words, and over/underflow (strtod returns +-inf / 0 like Rust)
  *value = std::strtod(tmp.c_str(), nullptr);
  return true;
}

struct VertexLine {
  double value[3];
};

// check_for_constant_coordinates (:140-190) keeps a HashMap of the coordinate STRINGS of every axis (keys starting with
// "0.0" or "-0.0" collapsed to "0.0") and only ever asks whether an axis has exactly one key: remember the first key and
// whether a different one turned up.
struct AxisKeys {
  std::string_view first;
  bool any = false, several = false;
  void add(std::string_view t) {
    if (t.substr(0, 3) == "0.0" || t.substr(0, 4) == "-0.0") t = "0.0";
    if (!any) first = t, any = true;
    else if (t != first) several = true;
  }
  bool constant() const { return any && !several; }
};


// ---- the line pass, shared by the 1-D and 2-D builders and run on several threads for big files ------------------------
// A chunk is a run of whole lines.  Every chunk parses its `v` lines (and, for 2-D, its `f` lines) independently; the
// results are concatenated in file order, and the first error in file order is the one reported, so the outcome is
// that of the single sequential pass the reference makes.
struct Chunk {
  std::vector<VertexLine> vs;
  std::vector<unsigned> triangles;
  AxisKeys keys[3];
  IngestStatus status = IngestStatus::Ok;
  const char* message = nullptr;
};

void parse_chunk(const std::string& text, size_t pos, size_t stop, bool faces, Chunk* out) {
  const char* const base = text.data();
  auto fail = [&](const char* msg) { out->status = IngestStatus::Parse, out->message = msg; };
  std::pair<size_t, size_t> fl[4];
  while (pos < stop) {
    const char* nl = static_cast<const char*>(std::memchr(base + pos, '\n', stop - pos));
    const size_t eol = nl ? (size_t)(nl - base) : stop;
    size_t end = eol;
    if (end > pos && text[end - 1] == '\r') --end;  // BufRead::lines strips "\r\n"
    const bool vline = end - pos >= 2 && text[pos] == 'v' && text[pos + 1] == ' ';
    const bool fline = faces && end - pos >= 2 && text[pos] == 'f' && text[pos + 1] == ' ';
    if (vline || fline) {
      // str::split(" ") after the tag: the first three fields are kept, the rest only counted (and, for `v`, parsed)
      size_t nf = 0, s = pos + 2;
      bool ok = true;
      VertexLine vl{};
      for (;;) {
        const char* bl = static_cast<const char*>(std::memchr(base + s, ' ', end - s));
        const size_t sp = bl ? (size_t)(bl - base) : end;
        if (vline) {
          double v = 0.0;
          if (!parse_rust_float(base + s, base + sp, &v)) ok = false;
          if (nf < 3) vl.value[nf] = v;
        }
        if (nf < 3) fl[nf] = {s, sp};
        ++nf;
        if (sp == end) break;
        s = sp + 1;
      }
      if (vline) {
        if (!ok) return fail("Error while parsing vertex coordinate from obj");
        if (nf != 3) return fail("A vertex line should contain 3 elements only");
        for (int a = 0; a < 3; ++a) out->keys[a].add(std::string_view(base + fl[a].first, fl[a].second - fl[a].first));
        out->vs.push_back(vl);
      } else {
        if (nf != 3) return fail("Amount of face specificating elements should be 3.");
        for (size_t k = 0; k < 3; ++k) {  // a/b/c: the first field is the vertex, exactly two more must follow
          const size_t b = fl[k].first, e = fl[k].second;
          const char* sl = static_cast<const char*>(std::memchr(base + b, '/', e - b));
          const size_t slash = sl ? (size_t)(sl - base) : e;
          unsigned long long v = 0;
          size_t i = b;
          if (i < slash && text[i] == '+') ++i;
          bool okf = i < slash;
          for (; i < slash; ++i) {
            if (text[i] < '0' || text[i] > '9') { okf = false; break; }
            v = v * 10 + (unsigned)(text[i] - '0');
            if (v > 0xffffffffULL) { okf = false; break; }
          }
          if (!okf) return fail("Error while parsing face coordinate");
          size_t extra = 0;
          for (size_t j = slash; j < e; ++j) extra += text[j] == '/';
          if (extra != 2) return fail("Amount of elements per face specification should be 3 in format a/b/c.");
          out->triangles.push_back((unsigned)v - 1u);
        }
      }
    }
    pos = eol + 1;
  }
}

// Reads the file and runs the line pass; vertices / triangles / per-axis keys in file order, or the first error.
IngestStatus read_and_parse(const char* path, bool faces, std::string* text, Chunk* all, std::string* message) {
  std::FILE* f = std::fopen(path, "rb");
  if (!f) {
    *message = std::string("cannot open ") + path;
    return IngestStatus::Io;
  }
  if (std::fseek(f, 0, SEEK_END) == 0) {  // one allocation of the file's size when it is known (regular files)
    const long len = std::ftell(f);
    if (len > 0) text->reserve((size_t)len);
    std::rewind(f);
  }
  char buf[1 << 18];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text->append(buf, got);
  std::fclose(f);

  const size_t size = text->size();
  size_t nthreads = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 32);
  nthreads = std::min(nthreads, size / (1u << 20) + 1);  // at least ~1 MB of text per thread
  std::vector<size_t> cut(nthreads + 1, size);
  cut[0] = 0;
  for (size_t k = 1; k < nthreads; ++k) {  // chunk boundaries on line starts
    size_t p = std::max(cut[k - 1], size * k / nthreads);
    const char* nl = p < size ? static_cast<const char*>(std::memchr(text->data() + p, '\n', size - p)) : nullptr;
    cut[k] = nl ? (size_t)(nl - text->data()) + 1 : size;
  }
  std::vector<Chunk> chunks(nthreads);
  std::vector<std::thread> pool;
  for (size_t k = 1; k < nthreads; ++k) pool.emplace_back(parse_chunk, std::cref(*text), cut[k], cut[k + 1], faces, &chunks[k]);
  parse_chunk(*text, cut[0], cut[1], faces, &chunks[0]);
  for (std::thread& t : pool) t.join();

  size_t nv = 0, nt = 0;
  for (const Chunk& c : chunks) {
    if (c.status != IngestStatus::Ok) {  // chunks before this one parsed cleanly: this is the first error of the file
      *message = c.message;
      return c.status;
    }
    nv += c.vs.size(), nt += c.triangles.size();
  }
  all->vs.reserve(nv), all->triangles.reserve(nt);
  for (const Chunk& c : chunks) {
    all->vs.insert(all->vs.end(), c.vs.begin(), c.vs.end());
    all->triangles.insert(all->triangles.end(), c.triangles.begin(), c.triangles.end());
    for (int a = 0; a < 3; ++a) {
      if (!c.keys[a].any) continue;
      all->keys[a].add(c.keys[a].first);  // (already collapsed; add() is idempotent on it)
      if (c.keys[a].several) all->keys[a].several = true;
    }
  }
  return IngestStatus::Ok;
}

}  // namespace

IngestStatus ingest_obj_1d(const char* path, bool has_multiplier, double multiplier, Ingested1D* out) {
  // One pass over the lines: keep those starting with "v " and split them on single blanks (`split(" ")`),
  // exactly 3 fields after the tag, each a Rust-parsable float (mesh_builder.rs:58-84 and :159-178).
  std::string text;
  Chunk all;
  const IngestStatus st = read_and_parse(path, false, &text, &all, &out->message);
  if (st != IngestStatus::Ok) return st;
  const std::vector<VertexLine>& vs = all.vs;
  const AxisKeys* keys = all.keys;

  // Axis detection (check_for_constant_coordinates, :140-190): an axis with one distinct coordinate string is constant.
  const bool cx = keys[0].constant(), cy = keys[1].constant(), cz = keys[2].constant();
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
  std::string text;
  Chunk all;
  const IngestStatus st = read_and_parse(path, true, &text, &all, &out->message);
  if (st != IngestStatus::Ok) return st;
  const std::vector<VertexLine>& vs = all.vs;
  const AxisKeys* keys = all.keys;
  out->triangles = std::move(all.triangles);
  int cc;  // :355-363
  if (keys[0].constant()) cc = 0;
  else if (keys[1].constant()) cc = 1;
  else if (keys[2].constant()) cc = 2;
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
