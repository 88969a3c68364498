// Host-side data formats of the hot path's two neighbours, native:
//   * ingest of the JSON files the reference's reader loads (data_reader.py:20-70,85-92) straight into
//     CSR arrays - no Python object per rating;
//   * the offline splitter that writes them (TrainValidTestSplit.py:31-219): CSV in, per-rating split,
//     per-row grouping in first-appearance order, input/target pairing, the same JSON / CSV bytes out.
// No CUDA in this file: it is compiled by the host compiler and linked into libocf_b200.so.
#include <algorithm>
#include <cctype>
#include <charconv>
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/ocf.h"

namespace ocf {
std::string& last_error();
int fail(int code, const std::string& msg);
}  // namespace ocf

namespace {

using ocf::fail;

// ---------------------------------------------------------------------------------------------
// file -> memory
// ---------------------------------------------------------------------------------------------
int read_file(const char* path, std::string* out) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(OCF_ERR_INVALID, std::string("cannot open ") + path + ": " + std::strerror(errno));
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (n < 0) {
    std::fclose(f);
    return fail(OCF_ERR_INVALID, std::string("cannot size ") + path);
  }
  out->resize((size_t)n);
  size_t got = n ? std::fread(&(*out)[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  if (got != (size_t)n) return fail(OCF_ERR_INVALID, std::string("short read of ") + path);
  return OCF_OK;
}

// ---------------------------------------------------------------------------------------------
// ids: what Python's dict lookup `items_to_densevec[item]` (data_reader.py:135) treats as one key.
// JSON numbers arrive as int or float; 153 == 153.0 and they hash alike, "153" is a different key.
// ---------------------------------------------------------------------------------------------
struct Id {
  enum Kind : uint8_t { INT, FLT, STR } kind = INT;
  int64_t i = 0;
  double d = 0.0;
  std::string s;
};

void canonicalise(Id* id) {
  if (id->kind == Id::FLT && std::nearbyint(id->d) == id->d && std::fabs(id->d) < 9.2e18) {
    id->kind = Id::INT;
    id->i = (int64_t)id->d;
  }
}

std::string show(const Id& id) {
  if (id.kind == Id::INT) return std::to_string(id.i);
  if (id.kind == Id::FLT) {
    char buf[40];
    std::snprintf(buf, sizeof buf, "%.17g", id.d);
    return buf;
  }
  return "'" + id.s + "'";
}

struct Vocab {
  int64_t size = 0;                                   // entries of the list (duplicates included)
  // dense window for integer ids (the usual case), maps for the rest
  int64_t lo = 0;
  std::vector<int32_t> window;                        // -1 = absent
  std::unordered_map<int64_t, int32_t> ints;
  std::unordered_map<uint64_t, int32_t> floats;       // bit pattern of a non-integral double
  std::unordered_map<std::string, int32_t> strs;

  void put(const Id& id, int32_t pos) {               // later duplicates win (data_reader.py:25-28)
    if (id.kind == Id::INT) ints[id.i] = pos;
    else if (id.kind == Id::FLT) { uint64_t b; std::memcpy(&b, &id.d, 8); floats[b] = pos; }
    else strs[id.s] = pos;
  }
  void freeze() {
    if (ints.empty()) return;
    int64_t mn = INT64_MAX, mx = INT64_MIN;
    for (auto& kv : ints) { mn = std::min(mn, kv.first); mx = std::max(mx, kv.first); }
    const uint64_t span = (uint64_t)mx - (uint64_t)mn;          // no signed overflow for ids near +-2^63
    if (span < (uint64_t)(1 << 28) && span < 64 * (uint64_t)ints.size() + 1024) {
      lo = mn;
      window.assign((size_t)(mx - mn + 1), -1);
      for (auto& kv : ints) window[(size_t)(kv.first - mn)] = kv.second;
    }
  }
  int32_t find(const Id& id) const {
    if (id.kind == Id::INT) {
      if (!window.empty()) {
        int64_t k = id.i - lo;
        return (k >= 0 && k < (int64_t)window.size()) ? window[(size_t)k] : -1;
      }
      auto it = ints.find(id.i);
      return it == ints.end() ? -1 : it->second;
    }
    if (id.kind == Id::FLT) {
      uint64_t b; std::memcpy(&b, &id.d, 8);
      auto it = floats.find(b);
      return it == floats.end() ? -1 : it->second;
    }
    auto it = strs.find(id.s);
    return it == strs.end() ? -1 : it->second;
  }
};

// ---------------------------------------------------------------------------------------------
// JSON reader (RFC 8259 + the NaN / Infinity literals Python's json module reads and writes)
// ---------------------------------------------------------------------------------------------
struct Json {
  const char* p;
  const char* end;
  const char* begin;
  std::string err;

  explicit Json(const std::string& text) : p(text.data()), end(text.data() + text.size()), begin(text.data()) {}

  bool bad(const std::string& what) {
    if (err.empty()) err = what + " at byte " + std::to_string((long long)(p - begin));
    return false;
  }
  void ws() {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
  }
  bool eat(char c) {
    ws();
    if (p < end && *p == c) { ++p; return true; }
    return false;
  }
  bool expect(char c) {
    if (eat(c)) return true;
    return bad(std::string("expected '") + c + "'");
  }
  char peek() {
    ws();
    return p < end ? *p : '\0';
  }
  bool literal(const char* word) {
    size_t n = std::strlen(word);
    if ((size_t)(end - p) >= n && std::memcmp(p, word, n) == 0) { p += n; return true; }
    return false;
  }

  static void utf8(uint32_t cp, std::string* out) {
    if (cp < 0x80) out->push_back((char)cp);
    else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
      out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out->push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      out->push_back((char)(0xF0 | (cp >> 18))); out->push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  bool hex4(uint32_t* v) {
    if (end - p < 4) return bad("truncated \\u escape");
    uint32_t x = 0;
    for (int k = 0; k < 4; ++k) {
      char c = p[k];
      x <<= 4;
      if (c >= '0' && c <= '9') x |= (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f') x |= (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') x |= (uint32_t)(c - 'A' + 10);
      else return bad("bad \\u escape");
    }
    p += 4;
    *v = x;
    return true;
  }
  // Reads a "..." token into out (UTF-8, escapes resolved).
  bool string(std::string* out) {
    ws();
    if (p >= end || *p != '"') return bad("expected a string");
    ++p;
    out->clear();
    for (;;) {
      const char* q = p;
      while (q < end && *q != '"' && *q != '\\') ++q;
      out->append(p, q);
      p = q;
      if (p >= end) return bad("unterminated string");
      if (*p == '"') { ++p; return true; }
      ++p;                                             // backslash
      if (p >= end) return bad("unterminated escape");
      char c = *p++;
      switch (c) {
        case '"': out->push_back('"'); break;
        case '\\': out->push_back('\\'); break;
        case '/': out->push_back('/'); break;
        case 'b': out->push_back('\b'); break;
        case 'f': out->push_back('\f'); break;
        case 'n': out->push_back('\n'); break;
        case 'r': out->push_back('\r'); break;
        case 't': out->push_back('\t'); break;
        case 'u': {
          uint32_t cp = 0;
          if (!hex4(&cp)) return false;
          if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
            const char* save = p;
            p += 2;
            uint32_t lo2 = 0;
            if (!hex4(&lo2)) return false;
            if (lo2 >= 0xDC00 && lo2 < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo2 - 0xDC00);
            else p = save;                             // lone surrogate: kept as is (like Python)
          }
          utf8(cp, out);
          break;
        }
        default: return bad("bad escape");
      }
    }
  }

  // A JSON number (or NaN / Infinity / -Infinity). is_int: no fraction or exponent and fits int64.
  bool number(double* d, int64_t* i, bool* is_int) {
    ws();
    const char* s = p;
    if (literal("NaN")) { *d = std::nan(""); *is_int = false; return true; }
    if (literal("Infinity")) { *d = HUGE_VAL; *is_int = false; return true; }
    if (literal("-Infinity")) { *d = -HUGE_VAL; *is_int = false; return true; }
    bool neg = false;
    if (p < end && *p == '-') { neg = true; ++p; }
    if (p >= end || *p < '0' || *p > '9') { p = s; return bad("expected a number"); }
    uint64_t mant = 0;
    int digits = 0, frac = 0;
    bool simple = true;                                // mantissa still exact in 64 bits
    while (p < end && *p >= '0' && *p <= '9') {
      if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant || digits) ++digits; }
      else simple = false;
      ++p;
    }
    bool integral = true;
    if (p < end && *p == '.') {
      integral = false;
      ++p;
      if (p >= end || *p < '0' || *p > '9') return bad("digits expected after '.'");
      while (p < end && *p >= '0' && *p <= '9') {
        if (digits < 19) { mant = mant * 10 + (uint64_t)(*p - '0'); if (mant || digits) ++digits; ++frac; }
        else simple = false;
        ++p;
      }
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
      integral = false;
      simple = false;
      ++p;
      if (p < end && (*p == '+' || *p == '-')) ++p;
      if (p >= end || *p < '0' || *p > '9') return bad("digits expected in exponent");
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    if (integral && simple && mant <= (uint64_t)INT64_MAX) {
      *is_int = true;
      *i = neg ? -(int64_t)mant : (int64_t)mant;
      *d = (double)*i;
      return true;
    }
    *is_int = false;
    static const double P10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    if (simple && mant < (1ull << 53) && frac <= 22) {
      // Clinger's fast path: both operands exact doubles, one correctly rounded division
      double v = (double)mant / P10[frac];
      *d = neg ? -v : v;
      return true;
    }
    std::string tmp(s, p);
    *d = std::strtod(tmp.c_str(), nullptr);            // correctly rounded (glibc)
    return true;
  }

  int depth = 0;
  struct Nest {                                         // bounds the recursion of skip_value on hostile input
    Json* j;
    explicit Nest(Json* js) : j(js) { ++j->depth; }
    ~Nest() { --j->depth; }
  };
  bool skip_value() {
    Nest nest(this);
    if (depth > 256) return bad("nesting deeper than 256 levels");
    char c = peek();
    if (c == '"') { std::string s; return string(&s); }
    if (c == '{') {
      ++p;
      if (eat('}')) return true;
      do {
        std::string k;
        if (!string(&k) || !expect(':') || !skip_value()) return false;
      } while (eat(','));
      return expect('}');
    }
    if (c == '[') {
      ++p;
      if (eat(']')) return true;
      do { if (!skip_value()) return false; } while (eat(','));
      return expect(']');
    }
    if (literal("null") || literal("true") || literal("false")) return true;
    double d; int64_t i; bool ii;
    return number(&d, &i, &ii);
  }

  bool id(Id* out) {
    char c = peek();
    if (c == '"') { out->kind = Id::STR; return string(&out->s); }
    bool is_int;
    if (!number(&out->d, &out->i, &is_int)) return bad("an id must be a number or a string");
    out->kind = is_int ? Id::INT : Id::FLT;
    canonicalise(out);
    return true;
  }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI: vocabulary (unique_items_list.json / unique_users_list.json, data_reader.py:20-28)
// ---------------------------------------------------------------------------------------------
struct ocf_vocab {
  Vocab v;
};

extern "C" int ocf_vocab_load_json(const char* path, ocf_vocab** out) {
  if (!path || !out) return fail(OCF_ERR_INVALID, "ocf_vocab_load_json: bad argument");
  *out = nullptr;
  std::string text;
  if (int rc = read_file(path, &text)) return rc;
  Json js(text);
  auto voc = std::make_unique<ocf_vocab>();
  if (!js.expect('[')) return fail(OCF_ERR_INVALID, std::string(path) + ": " + js.err);
  if (!js.eat(']')) {
    do {
      Id id;
      if (!js.id(&id)) return fail(OCF_ERR_INVALID, std::string(path) + ": " + js.err);
      if (voc->v.size >= INT32_MAX) return fail(OCF_ERR_INVALID, std::string(path) + ": more than 2^31 ids");
      voc->v.put(id, (int32_t)voc->v.size);
      voc->v.size++;
    } while (js.eat(','));
    if (!js.expect(']')) return fail(OCF_ERR_INVALID, std::string(path) + ": " + js.err);
  }
  js.ws();
  if (js.p != js.end) return fail(OCF_ERR_INVALID, std::string(path) + ": trailing data after the list");
  voc->v.freeze();
  *out = voc.release();
  return OCF_OK;
}

extern "C" int ocf_vocab_size(const ocf_vocab* vocab, int64_t* n) {
  if (!vocab || !n) return fail(OCF_ERR_INVALID, "ocf_vocab_size: bad argument");
  *n = vocab->v.size;
  return OCF_OK;
}

extern "C" int ocf_vocab_destroy(ocf_vocab* vocab) {
  delete vocab;
  return OCF_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI: rating dicts (ratingsBy{User,Item}_dict.json, ..._dicts_{train,valid,test}.json)
// ---------------------------------------------------------------------------------------------
namespace {

struct RowDict {                                       // one {"row key": [[id, rating], ...] | null} object
  std::vector<std::string> keys;                       // first-appearance order (Python dict order)
  std::vector<int64_t> start, len;                     // segment of col/val per key (the LAST value of a repeated key)
  std::vector<uint8_t> none;
  std::vector<int32_t> col;
  std::vector<float> val;
  std::unordered_map<std::string, int64_t> index;
  int64_t live = 0;                                    // ratings in live segments
};

bool parse_row_dict(Json* js, const Vocab& cols, RowDict* out, std::string* key_error) {
  if (!js->expect('{')) return false;
  if (js->eat('}')) return true;
  std::string key;
  Id id;
  do {
    if (!js->string(&key) || !js->expect(':')) return false;
    int64_t row;
    auto it = out->index.find(key);
    if (it == out->index.end()) {
      row = (int64_t)out->keys.size();
      out->index.emplace(key, row);
      out->keys.push_back(key);
      out->start.push_back(0);
      out->len.push_back(0);
      out->none.push_back(0);
    } else {
      row = it->second;                                // repeated key: the last value wins, the place stays
      out->live -= out->len[(size_t)row];
    }
    out->start[(size_t)row] = (int64_t)out->col.size();
    out->len[(size_t)row] = 0;
    out->none[(size_t)row] = 0;
    js->ws();
    if (js->literal("null")) {
      out->none[(size_t)row] = 1;
      continue;
    }
    if (!js->expect('[')) return false;
    if (!js->eat(']')) {
      do {
        if (!js->expect('[')) return false;
        if (!js->id(&id)) return false;
        if (!js->expect(',')) return false;
        double r; int64_t ri; bool is_int;
        if (!js->number(&r, &ri, &is_int)) return false;
        while (js->eat(',')) if (!js->skip_value()) return false;     // anything after [id, rating] is not read
        if (!js->expect(']')) return false;
        int32_t c = cols.find(id);
        if (c < 0) {
          *key_error = "KeyError: " + show(id) + " (row " + key + ") is not in the unique list";
          return false;
        }
        out->col.push_back(c);
        out->val.push_back((float)r);
      } while (js->eat(','));
      if (!js->expect(']')) return false;
    }
    out->len[(size_t)row] = (int64_t)out->col.size() - out->start[(size_t)row];
    out->live += out->len[(size_t)row];
  } while (js->eat(','));
  return js->expect('}');
}

}  // namespace

struct ocf_ratings {
  int paired = 0;
  RowDict a, b;                     // single: a. paired: a = inputs, b = targets (rows follow b's keys)
  std::vector<int64_t> a_row;       // paired: row of `a` holding key k of `b`
  int64_t key_bytes = 0;
};

extern "C" int ocf_ratings_load_json(const char* path, const ocf_vocab* cols, int paired, ocf_ratings** out) {
  if (!path || !cols || !out) return fail(OCF_ERR_INVALID, "ocf_ratings_load_json: bad argument");
  *out = nullptr;
  std::string text;
  if (int rc = read_file(path, &text)) return rc;
  Json js(text);
  auto r = std::make_unique<ocf_ratings>();
  r->paired = paired ? 1 : 0;
  std::string key_error;
  bool ok;
  if (paired) {
    // [input dict, target dict] (TrainValidTestSplit.py:175); the with-timestamps files nest one level deeper
    ok = js.expect('[');
    if (ok && js.peek() == '[')
      return fail(OCF_ERR_INVALID, std::string(path) + ": a '_withtimestamps_' file ([[inputs, targets], timestamps]); "
                                   "timestamps are out of scope (broken in the reference, data_reader.py:132,359,409)");
    ok = ok && parse_row_dict(&js, cols->v, &r->a, &key_error) && js.expect(',') &&
         parse_row_dict(&js, cols->v, &r->b, &key_error) && js.expect(']');
  } else {
    ok = parse_row_dict(&js, cols->v, &r->a, &key_error);
  }
  if (ok) {
    js.ws();
    if (js.p != js.end) ok = js.bad("trailing data");
  }
  if (!ok) return fail(OCF_ERR_INVALID, std::string(path) + ": " + (key_error.empty() ? js.err : key_error));
  const RowDict& rows = paired ? r->b : r->a;
  for (auto& k : rows.keys) r->key_bytes += (int64_t)k.size();
  if (paired) {
    r->a_row.resize(rows.keys.size());
    for (size_t k = 0; k < rows.keys.size(); ++k) {
      auto it = r->a.index.find(rows.keys[k]);
      if (it == r->a.index.end())
        return fail(OCF_ERR_INVALID, std::string(path) + ": KeyError: row '" + rows.keys[k] + "' has targets but no entry "
                                     "in the input dict (data_reader.py:232)");
      r->a_row[k] = it->second;
    }
  }
  *out = r.release();
  return OCF_OK;
}

extern "C" int ocf_ratings_info(const ocf_ratings* r, int64_t info[4]) {
  if (!r || !info) return fail(OCF_ERR_INVALID, "ocf_ratings_info: bad argument");
  const RowDict& rows = r->paired ? r->b : r->a;
  info[0] = (int64_t)rows.keys.size();
  info[1] = r->key_bytes;
  if (r->paired) {
    int64_t n = 0;
    for (int64_t ar : r->a_row) n += r->a.len[(size_t)ar];
    info[2] = n;
    info[3] = r->b.live;
  } else {
    info[2] = r->a.live;
    info[3] = 0;
  }
  return OCF_OK;
}

extern "C" int ocf_ratings_keys(const ocf_ratings* r, char* bytes, int64_t* offsets) {
  if (!r || !bytes || !offsets) return fail(OCF_ERR_INVALID, "ocf_ratings_keys: bad argument");
  const RowDict& rows = r->paired ? r->b : r->a;
  int64_t at = 0;
  for (size_t k = 0; k < rows.keys.size(); ++k) {
    offsets[k] = at;
    std::memcpy(bytes + at, rows.keys[k].data(), rows.keys[k].size());
    at += (int64_t)rows.keys[k].size();
  }
  offsets[rows.keys.size()] = at;
  return OCF_OK;
}

extern "C" int ocf_ratings_csr(const ocf_ratings* r, int which, int64_t* rowptr, int32_t* col, float* val,
                               uint8_t* none) {
  if (!r || !rowptr || !col || !val) return fail(OCF_ERR_INVALID, "ocf_ratings_csr: bad argument");
  if (which != 0 && !(which == 1 && r->paired)) return fail(OCF_ERR_INVALID, "ocf_ratings_csr: no such store");
  const RowDict& src = (r->paired && which == 1) ? r->b : r->a;
  const bool via = r->paired && which == 0;            // inputs of a paired file follow the target keys
  const size_t n = r->paired ? r->b.keys.size() : r->a.keys.size();
  int64_t at = 0;
  for (size_t k = 0; k < n; ++k) {
    size_t row = via ? (size_t)r->a_row[k] : k;
    rowptr[k] = at;
    int64_t len = src.len[row];
    if (len) {
      std::memcpy(col + at, src.col.data() + src.start[row], (size_t)len * sizeof(int32_t));
      std::memcpy(val + at, src.val.data() + src.start[row], (size_t)len * sizeof(float));
    }
    if (none) none[k] = src.none[row];
    at += len;
  }
  rowptr[n] = at;
  return OCF_OK;
}

extern "C" int ocf_ratings_destroy(ocf_ratings* r) {
  delete r;
  return OCF_OK;
}

// =============================================================================================
// The offline splitter (TrainValidTestSplit.py): ratings CSV -> per-rating train/valid/test split ->
// per-row dicts with paired inputs, written as the JSON / CSV bytes the reference's script writes.
// =============================================================================================
namespace {

// float.__repr__ (what json.dump and pandas' to_csv print for a float): shortest digits that round-trip,
// fixed notation for 1e-4 <= |x| < 1e16 with at least one fractional digit, else d[.ddd]e[+-]XX.
void py_float_repr(double x, std::string* out, const char* nan_text, const char* inf_text) {
  if (std::isnan(x)) { out->append(nan_text); return; }
  if (std::isinf(x)) { if (x < 0) out->push_back('-'); out->append(inf_text); return; }
  if (x == 0.0) { out->append(std::signbit(x) ? "-0.0" : "0.0"); return; }
  char buf[48];
  if (std::fabs(x) < 1e15 && x == std::trunc(x)) {     // ids and whole ratings: digits + ".0"
    auto r = std::to_chars(buf, buf + sizeof buf, (int64_t)x);
    out->append(buf, r.ptr);
    out->append(".0");
    return;
  }
  if (std::fabs(x) < 1e15 && x + x == std::trunc(x + x)) {   // half-star ratings: exact in binary, "<int>.5"
    if (x < 0) out->push_back('-');
    auto r = std::to_chars(buf, buf + sizeof buf, (int64_t)std::fabs(x));
    out->append(buf, r.ptr);
    out->append(".5");
    return;
  }
  auto res = std::to_chars(buf, buf + sizeof buf - 1, x, std::chars_format::scientific);   // shortest round-trip
  *res.ptr = '\0';
  const char* s = buf;
  if (*s == '-') { out->push_back('-'); ++s; }
  char digits[24];
  int nd = 0;
  const char* q = s;
  for (; q < res.ptr && *q != 'e'; ++q)
    if (*q != '.') digits[nd++] = *q;
  int e10 = std::atoi(q + 1);                          // value = d.ddd * 10^e10
  int decpt = e10 + 1;                                 // digits before the decimal point
  if (decpt > 16 || decpt < -3) {
    out->push_back(digits[0]);
    if (nd > 1) { out->push_back('.'); out->append(digits + 1, (size_t)(nd - 1)); }
    out->push_back('e');
    out->push_back(e10 < 0 ? '-' : '+');
    int a = e10 < 0 ? -e10 : e10;
    if (a < 10) out->push_back('0');
    out->append(std::to_string(a));
  } else if (decpt <= 0) {
    out->append("0.");
    out->append((size_t)(-decpt), '0');
    out->append(digits, (size_t)nd);
  } else if (decpt >= nd) {
    out->append(digits, (size_t)nd);
    out->append((size_t)(decpt - nd), '0');
    out->append(".0");
  } else {
    out->append(digits, (size_t)decpt);
    out->push_back('.');
    out->append(digits + decpt, (size_t)(nd - decpt));
  }
}

// json.dumps(str) with ensure_ascii=True
void json_string(const std::string& s, std::string* out) {
  static const char* HEX = "0123456789abcdef";
  auto u16 = [&](uint32_t v) {
    out->append("\\u");
    out->push_back(HEX[(v >> 12) & 15]); out->push_back(HEX[(v >> 8) & 15]);
    out->push_back(HEX[(v >> 4) & 15]); out->push_back(HEX[v & 15]);
  };
  out->push_back('"');
  size_t i = 0, n = s.size();
  while (i < n) {
    unsigned char c = (unsigned char)s[i];
    if (c == '"') { out->append("\\\""); ++i; }
    else if (c == '\\') { out->append("\\\\"); ++i; }
    else if (c == '\n') { out->append("\\n"); ++i; }
    else if (c == '\r') { out->append("\\r"); ++i; }
    else if (c == '\t') { out->append("\\t"); ++i; }
    else if (c == '\b') { out->append("\\b"); ++i; }
    else if (c == '\f') { out->append("\\f"); ++i; }
    else if (c < 0x20) { u16(c); ++i; }
    else if (c < 0x80) { out->push_back((char)c); ++i; }
    else {
      uint32_t cp; int extra;
      if ((c & 0xE0) == 0xC0) { cp = c & 0x1F; extra = 1; }
      else if ((c & 0xF0) == 0xE0) { cp = c & 0x0F; extra = 2; }
      else if ((c & 0xF8) == 0xF0) { cp = c & 0x07; extra = 3; }
      else { cp = 0xFFFD; extra = 0; }
      ++i;
      for (int k = 0; k < extra && i < n; ++k, ++i) cp = (cp << 6) | ((unsigned char)s[i] & 0x3F);
      if (cp >= 0x10000) { cp -= 0x10000; u16(0xD800 + (cp >> 10)); u16(0xDC00 + (cp & 0x3FF)); }
      else u16(cp);
    }
  }
  out->push_back('"');
}

struct Column {
  enum Type : uint8_t { INT, FLT, STR } type = INT;
  std::vector<std::string> s;                          // STR columns only; numbers live in ocf_csv::packed
};

bool parse_int_field(const char* b, const char* e, int64_t* v) {
  if (b == e) return false;
  auto r = std::from_chars((*b == '+') ? b + 1 : b, e, *v);
  return r.ec == std::errc() && r.ptr == e && !(*b == '+' && b + 1 < e && b[1] == '-');
}
bool parse_float_field(const char* b, const char* e, double* v) {
  if (b == e) { *v = std::nan(""); return true; }     // empty field = NaN (pandas)
  std::string t(b, e);
  if (t == "NaN" || t == "nan" || t == "NA" || t == "N/A" || t == "NULL" || t == "null") { *v = std::nan(""); return true; }
  char* stop = nullptr;
  errno = 0;
  *v = std::strtod(t.c_str(), &stop);
  return stop == t.c_str() + t.size() && !std::isspace((unsigned char)t[0]);
}

struct RowKind { enum K : uint8_t { INT, FLT, STR }; };

}  // namespace

struct ocf_csv {
  int n_cols = 0;
  int64_t n_rows = 0;
  Column col[4];
  uint8_t row_kind = RowKind::INT;   // what `ratings.iloc[i]` upcasts a row to (TrainValidTestSplit.py:125)
  // the numeric columns of a row side by side (int64 or double bits, 4 slots per row): the writers visit rows
  // in permuted order, so one cache line per rating instead of one per column
  std::vector<uint64_t> packed;
  int64_t int_at(int c, int64_t row) const { int64_t v; std::memcpy(&v, &packed[(size_t)row * 4 + c], 8); return v; }
  double flt_at(int c, int64_t row) const { double v; std::memcpy(&v, &packed[(size_t)row * 4 + c], 8); return v; }
};

namespace {

// One pass over the records of a CSV text (RFC 4180 quoting; header line dropped; blank lines skipped like
// pandas' skip_blank_lines): cell(column, begin, end) per field. Returns the number of records or -1 (error set).
template <typename Cell>
int64_t scan_csv(const std::string& text, int n_columns, const char* path, Cell&& cell_cb) {
  const char* p = text.data();
  const char* end = p + text.size();
  int64_t line = 0, records = 0;
  std::string cell;
  while (p < end) {
    int c = 0;
    bool any = false;
    for (;;) {                                         // one record
      const char* b = p;
      const char* e;
      if (p < end && *p == '"') {
        cell.clear();
        ++p;
        for (;;) {
          if (p >= end) { fail(OCF_ERR_INVALID, std::string(path) + ": unterminated quoted field"); return -1; }
          if (*p == '"') {
            if (p + 1 < end && p[1] == '"') { cell.push_back('"'); p += 2; continue; }
            ++p;
            break;
          }
          cell.push_back(*p++);
        }
        const char* t = p;
        while (p < end && *p != ',' && *p != '\n' && *p != '\r') ++p;
        cell.append(t, p);
        b = cell.data();
        e = b + cell.size();
        any = true;
      } else {
        while (p < end && *p != ',' && *p != '\n' && *p != '\r') ++p;
        e = p;
        if (e > b) any = true;
      }
      const bool more = p < end && *p == ',';
      if (line > 0 && (any || more || c > 0)) {
        if (c >= n_columns) {
          fail(OCF_ERR_INVALID, std::string(path) + ": line " + std::to_string(line + 1) + " has more than " +
                                std::to_string(n_columns) + " fields");
          return -1;
        }
        cell_cb(c, b, e);
      }
      ++c;
      if (more) { ++p; any = true; continue; }
      break;
    }
    if (p < end && *p == '\r') ++p;
    if (p < end && *p == '\n') ++p;
    const bool blank = !any && c == 1;
    if (line > 0) {
      if (!blank) {
        if (c != n_columns) {
          fail(OCF_ERR_INVALID, std::string(path) + ": line " + std::to_string(line + 1) + " has " + std::to_string(c) +
                                " fields, expected " + std::to_string(n_columns));
          return -1;
        }
        ++records;
      }
    } else if (c != n_columns) {
      fail(OCF_ERR_INVALID, std::string(path) + ": the header has " + std::to_string(c) + " columns, the schema needs " +
                            std::to_string(n_columns) + " (TrainValidTestSplit.py:39-69)");
      return -1;
    }
    ++line;
  }
  return records;
}

}  // namespace

extern "C" int ocf_csv_load(const char* path, int n_columns, ocf_csv** out) {
  if (!path || !out || (n_columns != 3 && n_columns != 4)) return fail(OCF_ERR_INVALID, "ocf_csv_load: bad argument");
  *out = nullptr;
  std::string text;
  if (int rc = read_file(path, &text)) return rc;
  auto csv = std::make_unique<ocf_csv>();
  csv->n_cols = n_columns;
  // pass 1: column types as pandas.read_csv infers them - int64 if every field is an integer, float64 if every
  // field is a number (empty = NaN), else strings. Nothing is stored: a Netflix-sized file has 4 * 10^8 fields.
  bool all_int[4] = {true, true, true, true}, all_num[4] = {true, true, true, true};
  int64_t rows = scan_csv(text, n_columns, path, [&](int c, const char* b, const char* e) {
    int64_t iv; double dv;
    if (all_int[c] && !parse_int_field(b, e, &iv)) all_int[c] = false;
    if (!all_int[c] && all_num[c] && !parse_float_field(b, e, &dv)) all_num[c] = false;
  });
  if (rows < 0) return OCF_ERR_INVALID;
  csv->n_rows = rows;
  for (int c = 0; c < n_columns; ++c) {
    Column& col = csv->col[c];
    col.type = all_int[c] ? Column::INT : (all_num[c] ? Column::FLT : Column::STR);
    if (col.type == Column::STR) { csv->row_kind = RowKind::STR; col.s.reserve((size_t)rows); }
    else if (col.type == Column::FLT && csv->row_kind == RowKind::INT) csv->row_kind = RowKind::FLT;
  }
  // pass 2: values into their typed homes
  csv->packed.assign((size_t)rows * 4, 0);
  int64_t r = 0;
  scan_csv(text, n_columns, path, [&](int c, const char* b, const char* e) {
    Column& col = csv->col[c];
    if (col.type == Column::INT) { int64_t v = 0; parse_int_field(b, e, &v); std::memcpy(&csv->packed[(size_t)r * 4 + c], &v, 8); }
    else if (col.type == Column::FLT) { double v = 0; parse_float_field(b, e, &v); std::memcpy(&csv->packed[(size_t)r * 4 + c], &v, 8); }
    else col.s.emplace_back(b, e);
    if (c == n_columns - 1) ++r;
  });
  *out = csv.release();
  return OCF_OK;
}

extern "C" int ocf_csv_rows(const ocf_csv* csv, int64_t* n) {
  if (!csv || !n) return fail(OCF_ERR_INVALID, "ocf_csv_rows: bad argument");
  *n = csv->n_rows;
  return OCF_OK;
}

extern "C" int ocf_csv_destroy(ocf_csv* csv) {
  delete csv;
  return OCF_OK;
}

namespace {

struct Out {                                           // buffered file writer
  FILE* f = nullptr;
  std::string buf;
  bool ok = true;
  bool open(const std::string& path) {
    f = std::fopen(path.c_str(), "wb");
    buf.reserve(1 << 20);
    return f != nullptr;
  }
  void flush() {
    if (f && !buf.empty() && std::fwrite(buf.data(), 1, buf.size(), f) != buf.size()) ok = false;
    buf.clear();
  }
  void tick() { if (buf.size() > (1 << 20) - 4096) flush(); }
  bool close() {
    flush();
    if (f && std::fclose(f) != 0) ok = false;
    f = nullptr;
    return ok;
  }
};

struct Splitter {
  const ocf_csv& csv;
  int user_col, item_col, rating_col = 2, ts_col = 3;
  bool cast_user_to_int;
  std::vector<int32_t> uid;                            // per CSV row: id of its user KEY (the dict key string)
  std::vector<std::string> ukey;
  std::string error;

  Splitter(const ocf_csv& c, bool reverse, bool cast) : csv(c), user_col(reverse ? 1 : 0), item_col(reverse ? 0 : 1), cast_user_to_int(cast) {}
  int64_t int_at(int c, int64_t row) const { return csv.int_at(c, row); }
  double flt_at(int c, int64_t row) const { return csv.flt_at(c, row); }
  // one value of a row as json.dump prints it after the row went through `ratings.iloc[i]`
  void json_value(int c, int64_t row, std::string* out) const {
    const Column& col = csv.col[c];
    if (col.type == Column::STR) { json_string(col.s[(size_t)row], out); return; }
    if (csv.row_kind == RowKind::FLT || col.type == Column::FLT) {
      py_float_repr(col.type == Column::INT ? (double)int_at(c, row) : flt_at(c, row), out, "NaN", "Infinity");
      return;
    }
    char buf[24];
    auto r = std::to_chars(buf, buf + sizeof buf, int_at(c, row));
    out->append(buf, r.ptr);
  }
  // one value as DataFrame.to_csv prints it (the column's own dtype)
  void csv_value(int c, int64_t row, std::string* out) const {
    const Column& col = csv.col[c];
    if (col.type == Column::INT) {
      char buf[24];
      auto r = std::to_chars(buf, buf + sizeof buf, int_at(c, row));
      out->append(buf, r.ptr);
      return;
    }
    if (col.type == Column::FLT) {
      if (!std::isnan(flt_at(c, row))) py_float_repr(flt_at(c, row), out, "", "inf");
      return;
    }
    const std::string& s = col.s[(size_t)row];
    if (s.find_first_of(",\"\r\n") == std::string::npos) { out->append(s); return; }
    out->push_back('"');
    for (char ch : s) { if (ch == '"') out->push_back('"'); out->push_back(ch); }
    out->push_back('"');
  }
  // the dict key of a row's user (TrainValidTestSplit.py:126-135)
  bool user_key(int64_t row, std::string* out) {
    const Column& col = csv.col[user_col];
    out->clear();
    if (cast_user_to_int) {                            // str(int(row["userId"]))
      if (col.type == Column::STR) {
        error = "ValueError: invalid literal for int(): '" + col.s[(size_t)row] + "' (schema 'movielens' casts user ids to int)";
        return false;
      }
      if (col.type == Column::INT) { out->append(std::to_string(int_at(user_col, row))); return true; }
      double d = flt_at(user_col, row);
      if (std::isnan(d) || std::isinf(d)) { error = "ValueError: cannot convert float NaN/inf to integer (user id)"; return false; }
      out->append(std::to_string((int64_t)std::trunc(d)));
      return true;
    }
    if (col.type == Column::STR) { out->append(col.s[(size_t)row]); return true; }
    if (csv.row_kind == RowKind::FLT || col.type == Column::FLT)
      py_float_repr(col.type == Column::INT ? (double)int_at(user_col, row) : flt_at(user_col, row), out, "nan", "inf");
    else
      out->append(std::to_string(int_at(user_col, row)));
    return true;
  }
  bool index_users() {
    uid.resize((size_t)csv.n_rows);
    std::unordered_map<std::string, int32_t> by_key;
    std::unordered_map<int64_t, int32_t> by_int;       // INT user columns: skip the string work per row
    const Column& col = csv.col[user_col];
    std::string key;
    for (int64_t r = 0; r < csv.n_rows; ++r) {
      if (col.type == Column::INT) {
        auto it = by_int.find(csv.int_at(user_col, r));
        if (it != by_int.end()) { uid[(size_t)r] = it->second; continue; }
      }
      if (!user_key(r, &key)) return false;
      auto ins = by_key.emplace(key, (int32_t)ukey.size());
      if (ins.second) ukey.push_back(key);
      uid[(size_t)r] = ins.first->second;
      if (col.type == Column::INT) by_int.emplace(csv.int_at(user_col, r), ins.first->second);
    }
    return true;
  }
};

// build_user_item_dict (TrainValidTestSplit.py:121-149): rows of a subset grouped by user key,
// users in first-appearance order, ratings in subset order.
struct Groups {
  std::vector<int32_t> users;                          // uid per group
  std::vector<int64_t> start;                          // groups + 1
  std::vector<int64_t> rows;                           // CSV rows, grouped
  std::vector<int32_t> group_of;                       // uid -> group or -1

  void build(const Splitter& sp, const int64_t* subset, int64_t n) {
    group_of.assign(sp.ukey.size(), -1);
    std::vector<int64_t> count;
    for (int64_t k = 0; k < n; ++k) {
      int32_t u = sp.uid[(size_t)subset[k]];
      if (group_of[(size_t)u] < 0) { group_of[(size_t)u] = (int32_t)users.size(); users.push_back(u); count.push_back(0); }
      count[(size_t)group_of[(size_t)u]]++;
    }
    start.assign(users.size() + 1, 0);
    for (size_t g = 0; g < users.size(); ++g) start[g + 1] = start[g] + count[g];
    rows.resize((size_t)n);
    std::vector<int64_t> at(start.begin(), start.end() - 1);
    for (int64_t k = 0; k < n; ++k) rows[(size_t)at[(size_t)group_of[(size_t)sp.uid[(size_t)subset[k]]]]++] = subset[k];
  }
};

// the [item, value] pairs of one group, comma-separated (no brackets); `first` tracks the separator
void write_items(const Splitter& sp, const Groups& g, int32_t group, int value_col, Out* o, bool* first) {
  std::string& b = o->buf;
  for (int64_t k = g.start[(size_t)group]; k < g.start[(size_t)group + 1]; ++k) {
    if (!*first) b.append(", ");
    *first = false;
    b.push_back('[');
    sp.json_value(sp.item_col, g.rows[(size_t)k], &b);
    b.append(", ");
    sp.json_value(value_col, g.rows[(size_t)k], &b);
    b.push_back(']');
    o->tick();
  }
}

void write_list(const Splitter& sp, const Groups& g, int32_t group, int value_col, Out* o) {
  bool first = true;
  o->buf.push_back('[');
  write_items(sp, g, group, value_col, o, &first);
  o->buf.push_back(']');
}

// {user: [[item, value], ...]} of one subset
void write_dict(const Splitter& sp, const Groups& g, int value_col, Out* o) {
  o->buf.push_back('{');
  for (size_t k = 0; k < g.users.size(); ++k) {
    if (k) o->buf.append(", ");
    json_string(sp.ukey[(size_t)g.users[k]], &o->buf);
    o->buf.append(": ");
    write_list(sp, g, (int32_t)k, value_col, o);
  }
  o->buf.push_back('}');
}

// map_inputs_to_targets (:183-195): per target user the whole row of the input set, or null
void write_paired_inputs(const Splitter& sp, const Groups& tg, const Groups& in, int value_col, Out* o) {
  o->buf.push_back('{');
  for (size_t k = 0; k < tg.users.size(); ++k) {
    if (k) o->buf.append(", ");
    int32_t u = tg.users[k];
    json_string(sp.ukey[(size_t)u], &o->buf);
    o->buf.append(": ");
    int32_t gi = in.group_of[(size_t)u];
    if (gi < 0) o->buf.append("null"); else write_list(sp, in, gi, value_col, o);
  }
  o->buf.push_back('}');
}

// merge_timestamps (:197-211): every input-set user (input order) with the target list appended, then
// the users only the targets have
void write_merged_timestamps(const Splitter& sp, const Groups& in, const Groups& tg, Out* o) {
  o->buf.push_back('{');
  bool first_user = true;
  auto key = [&](int32_t u) {
    if (!first_user) o->buf.append(", ");
    first_user = false;
    json_string(sp.ukey[(size_t)u], &o->buf);
    o->buf.append(": ");
  };
  for (size_t k = 0; k < in.users.size(); ++k) {
    int32_t u = in.users[k];
    key(u);
    bool first = true;
    o->buf.push_back('[');
    write_items(sp, in, (int32_t)k, sp.ts_col, o, &first);
    if (tg.group_of[(size_t)u] >= 0) write_items(sp, tg, tg.group_of[(size_t)u], sp.ts_col, o, &first);
    o->buf.push_back(']');
  }
  for (size_t k = 0; k < tg.users.size(); ++k) {
    int32_t u = tg.users[k];
    if (in.group_of[(size_t)u] >= 0) continue;
    key(u);
    write_list(sp, tg, (int32_t)k, sp.ts_col, o);
  }
  o->buf.push_back('}');
}

}  // namespace

namespace {

// ratings[col].unique(): the rows where a value of column c appears for the first time, in file order. With
// `dense`, also the position of every row's value in that list (the dense id data_reader.py:24-28 would give it).
std::vector<int64_t> unique_first_rows(const ocf_csv& csv, int c, std::vector<int32_t>* dense) {
  std::vector<int64_t> firsts;
  const Column& col = csv.col[c];
  const int64_t n = csv.n_rows;
  if (dense) dense->resize((size_t)n);
  auto visit = [&](auto& seen, const auto& key, int64_t r) {
    auto ins = seen.emplace(key, (int32_t)firsts.size());
    if (ins.second) firsts.push_back(r);
    if (dense) (*dense)[(size_t)r] = ins.first->second;
  };
  if (col.type == Column::STR) {
    std::unordered_map<std::string, int32_t> seen;
    for (int64_t r = 0; r < n; ++r) visit(seen, col.s[(size_t)r], r);
  } else if (col.type == Column::INT) {
    std::unordered_map<int64_t, int32_t> seen;
    for (int64_t r = 0; r < n; ++r) visit(seen, csv.int_at(c, r), r);
  } else {
    std::unordered_map<uint64_t, int32_t> seen;
    for (int64_t r = 0; r < n; ++r) {
      double d = csv.flt_at(c, r);
      if (std::isnan(d)) d = std::nan("");              // every NaN is one value
      if (d == 0.0) d = 0.0;                           // -0.0 and 0.0 are one value
      uint64_t b; std::memcpy(&b, &d, 8);
      visit(seen, b, r);
    }
  }
  return firsts;
}

}  // namespace

extern "C" int ocf_split_write(const ocf_csv* csv, const int64_t* order, int64_t n_order, const double fractions[3],
                               const char* out_dir, int cast_user_to_int, int build_data_for_omni,
                               int include_timestamps, int save_users_and_items, int reverse_user_item_data) {
  if (!csv || !order || !fractions || !out_dir) return fail(OCF_ERR_INVALID, "ocf_split_write: bad argument");
  const int64_t n = csv->n_rows;
  if (n_order != n) return fail(OCF_ERR_INVALID, "ocf_split_write: the rating order must be a permutation of all rows");
  if (include_timestamps && csv->n_cols < 4)
    return fail(OCF_ERR_INVALID, "KeyError: 'timestamp' (include_timestamps with a 3-column schema, TrainValidTestSplit.py:138)");
  {
    std::vector<uint8_t> seen((size_t)n, 0);
    for (int64_t k = 0; k < n; ++k) {
      if (order[k] < 0 || order[k] >= n || seen[(size_t)order[k]])
        return fail(OCF_ERR_INVALID, "ocf_split_write: the rating order is not a permutation");
      seen[(size_t)order[k]] = 1;
    }
  }
  const std::string dir(out_dir);
  for (int k = 0; k < 3; ++k)
    if (!(fractions[k] >= 0.0 && fractions[k] <= 1.0)) return fail(OCF_ERR_INVALID, "ocf_split_write: a split fraction is outside [0, 1]");
  const int64_t n_tr = (int64_t)((double)n * fractions[0]);                  // :76-78 int(num_ratings * f)
  const int64_t n_va = (int64_t)((double)n * fractions[1]);
  if (n_tr + n_va > n) return fail(OCF_ERR_INVALID, "ocf_split_write: split fractions exceed 1");
  const int64_t* tr = order;
  const int64_t* va = order + n_tr;
  const int64_t* te = order + n_tr + n_va;
  const int64_t n_te = n - n_tr - n_va, n_in = n_tr + n_va;
  Splitter sp(*csv, reverse_user_item_data != 0, cast_user_to_int != 0);
  const char* suffix = include_timestamps ? "_withtimestamps" : "";

  // The output files are independent of each other: each is produced by its own thread (a task returns its
  // error message, empty = ok; the first one is reported). Only the per-row dicts wait for the user index.
  std::vector<std::thread> threads;
  std::vector<std::unique_ptr<std::string>> errors;
  auto spawn = [&](std::function<std::string()> task) {
    errors.emplace_back(new std::string());
    std::string* slot = errors.back().get();
    threads.emplace_back([task, slot] { *slot = task(); });
  };
  auto join_all = [&]() -> int {
    for (auto& t : threads) t.join();
    threads.clear();
    for (auto& e : errors)
      if (!e->empty()) return fail(OCF_ERR_INVALID, *e);
    return OCF_OK;
  };

  // convert_and_save_mml (:213-219): train+valid and test as userId,itemId,rating[,timestamp] lines
  auto save_mml = [&sp, &dir, include_timestamps](const int64_t* rows, int64_t count, const std::string& name) -> std::string {
    Out o;
    if (!o.open(dir + name)) return "cannot write " + dir + name;
    for (int64_t k = 0; k < count; ++k) {
      sp.csv_value(sp.user_col, rows[k], &o.buf); o.buf.push_back(',');
      sp.csv_value(sp.item_col, rows[k], &o.buf); o.buf.push_back(',');
      sp.csv_value(sp.rating_col, rows[k], &o.buf);
      if (include_timestamps) { o.buf.push_back(','); sp.csv_value(sp.ts_col, rows[k], &o.buf); }
      o.buf.push_back('\n');
      o.tick();
    }
    return o.close() ? "" : "write failed: " + dir + name;
  };
  spawn([&] { return save_mml(order, n_in, std::string("train_data_mml") + suffix + ".csv"); });
  spawn([&] { return save_mml(te, n_te, std::string("test_data_mml") + suffix + ".csv"); });

  Groups g_tr, g_va, g_te, g_in;
  if (build_data_for_omni) {
    if (!sp.index_users()) { join_all(); return fail(OCF_ERR_INVALID, sp.error); }
    {
      std::thread a([&] { g_tr.build(sp, tr, n_tr); }), b([&] { g_va.build(sp, va, n_va); }), c([&] { g_te.build(sp, te, n_te); });
      g_in.build(sp, order, n_in);                      // :102 a fresh dict of train+valid, in split order
      a.join(); b.join(); c.join();
    }
    auto path = [dir, suffix](const char* which) { return dir + "ratingsByUser_dicts" + suffix + "_" + which + ".json"; };
    const bool ts = include_timestamps != 0;
    spawn([&, path, ts]() -> std::string {              // train: the dict, or (ratings, timestamps) (:162-166)
      Out o;
      if (!o.open(path("train"))) return "cannot write " + path("train");
      if (ts) o.buf.push_back('[');
      write_dict(sp, g_tr, sp.rating_col, &o);
      if (ts) { o.buf.append(", "); write_dict(sp, g_tr, sp.ts_col, &o); o.buf.push_back(']'); }
      return o.close() ? "" : "write failed: " + path("train");
    });
    auto paired = [&sp, path, ts](const char* which, const Groups& tg, const Groups& in) -> std::string {   // :167-175
      Out o;
      if (!o.open(path(which))) return "cannot write " + path(which);
      o.buf.push_back('[');
      if (ts) o.buf.push_back('[');
      write_paired_inputs(sp, tg, in, sp.rating_col, &o);
      o.buf.append(", ");
      write_dict(sp, tg, sp.rating_col, &o);
      if (ts) { o.buf.append("], "); write_merged_timestamps(sp, in, tg, &o); }
      o.buf.push_back(']');
      return o.close() ? "" : "write failed: " + path(which);
    };
    spawn([&, paired] { return paired("valid", g_va, g_tr); });
    spawn([&, paired] { return paired("test", g_te, g_in); });
  }
  if (int rc = join_all()) return rc;

  if (save_users_and_items) {                           // :105-118, ids in first-appearance order of the file
    Out items, users;
    if (!items.open(dir + "unique_items_list.json") || !users.open(dir + "unique_users_list.json"))
      return fail(OCF_ERR_INVALID, "cannot write the unique lists under " + dir);
    auto unique_rows = [&](int c) { return unique_first_rows(*csv, c, nullptr); };
    // json.dump(list(ratings["itemId"].unique())): the column's own dtype
    items.buf.push_back('[');
    bool first = true;
    for (int64_t r : unique_rows(sp.item_col)) {
      if (!first) items.buf.append(", ");
      first = false;
      const Column& col = csv->col[sp.item_col];
      if (col.type == Column::STR) json_string(col.s[(size_t)r], &items.buf);
      else if (col.type == Column::INT) items.buf.append(std::to_string(csv->int_at(sp.item_col, r)));
      else py_float_repr(csv->flt_at(sp.item_col, r), &items.buf, "NaN", "Infinity");
      items.tick();
    }
    items.buf.push_back(']');
    // [str(int(x))] for movielens, [str(x)] otherwise - of the column's own values
    users.buf.push_back('[');
    first = true;
    std::string key;
    for (int64_t r : unique_rows(sp.user_col)) {
      if (!first) users.buf.append(", ");
      first = false;
      const Column& col = csv->col[sp.user_col];
      key.clear();
      if (col.type == Column::STR) {
        if (cast_user_to_int) return fail(OCF_ERR_INVALID, "ValueError: invalid literal for int(): '" + col.s[(size_t)r] + "'");
        key = col.s[(size_t)r];
      } else if (col.type == Column::INT) key = std::to_string(csv->int_at(sp.user_col, r));
      else if (cast_user_to_int) {
        const double d = csv->flt_at(sp.user_col, r);
        if (std::isnan(d) || std::isinf(d))
          return fail(OCF_ERR_INVALID, "ValueError: cannot convert float NaN/inf to integer (user id)");
        key = std::to_string((int64_t)std::trunc(d));
      } else py_float_repr(csv->flt_at(sp.user_col, r), &key, "nan", "inf");
      json_string(key, &users.buf);
      users.tick();
    }
    users.buf.push_back(']');
    if (!items.close() || !users.close()) return fail(OCF_ERR_INVALID, "write failed: unique lists under " + dir);
  }
  return OCF_OK;
}

// ---------------------------------------------------------------------------------------------
// The same split without the files: CSV -> the stores the reader builds from the splitter's output
// (what ocf_split_write + ocf_ratings_load_json give together, minus 2 x the JSON text).
// ---------------------------------------------------------------------------------------------
struct ocf_split {
  std::unique_ptr<Splitter> sp;
  Groups g[4];                        // train, valid targets, test targets, test inputs (= train + valid)
  std::vector<int32_t> item_dense;    // per CSV row: dense column of its item (first-appearance order of the file)
  int64_t n_cols = 0;
  const ocf_csv* csv = nullptr;
};

extern "C" int ocf_split_build(const ocf_csv* csv, const int64_t* order, int64_t n_order, const double fractions[3],
                               int cast_user_to_int, int reverse_user_item_data, ocf_split** out) {
  if (!csv || !order || !fractions || !out) return fail(OCF_ERR_INVALID, "ocf_split_build: bad argument");
  *out = nullptr;
  const int64_t n = csv->n_rows;
  if (n_order != n) return fail(OCF_ERR_INVALID, "ocf_split_build: the rating order must be a permutation of all rows");
  {
    std::vector<uint8_t> seen((size_t)n, 0);
    for (int64_t k = 0; k < n; ++k) {
      if (order[k] < 0 || order[k] >= n || seen[(size_t)order[k]])
        return fail(OCF_ERR_INVALID, "ocf_split_build: the rating order is not a permutation");
      seen[(size_t)order[k]] = 1;
    }
  }
  for (int k = 0; k < 3; ++k)
    if (!(fractions[k] >= 0.0 && fractions[k] <= 1.0)) return fail(OCF_ERR_INVALID, "ocf_split_build: a split fraction is outside [0, 1]");
  const int64_t n_tr = (int64_t)((double)n * fractions[0]);
  const int64_t n_va = (int64_t)((double)n * fractions[1]);
  if (n_tr + n_va > n) return fail(OCF_ERR_INVALID, "ocf_split_build: split fractions exceed 1");
  if (csv->col[2].type == Column::STR) return fail(OCF_ERR_INVALID, "ocf_split_build: the rating column is not numeric");
  auto s = std::make_unique<ocf_split>();
  s->csv = csv;
  s->sp = std::make_unique<Splitter>(*csv, reverse_user_item_data != 0, cast_user_to_int != 0);
  if (!s->sp->index_users()) return fail(OCF_ERR_INVALID, s->sp->error);
  s->n_cols = (int64_t)unique_first_rows(*csv, s->sp->item_col, &s->item_dense).size();
  s->g[0].build(*s->sp, order, n_tr);
  s->g[1].build(*s->sp, order + n_tr, n_va);
  s->g[2].build(*s->sp, order + n_tr + n_va, n - n_tr - n_va);
  s->g[3].build(*s->sp, order, n_tr + n_va);
  *out = s.release();
  return OCF_OK;
}

namespace {
// (targets, inputs) groups of a set: train has no inputs of its own
inline const Groups& split_targets(const ocf_split* s, int set) { return s->g[set]; }
inline const Groups* split_inputs(const ocf_split* s, int set) { return set == 1 ? &s->g[0] : (set == 2 ? &s->g[3] : nullptr); }
}  // namespace

extern "C" int ocf_split_info(const ocf_split* s, int64_t info[13]) {
  if (!s || !info) return fail(OCF_ERR_INVALID, "ocf_split_info: bad argument");
  info[0] = s->n_cols;
  for (int set = 0; set < 3; ++set) {
    const Groups& tg = split_targets(s, set);
    const Groups* in = split_inputs(s, set);
    int64_t key_bytes = 0, n_in = 0;
    for (int32_t u : tg.users) {
      key_bytes += (int64_t)s->sp->ukey[(size_t)u].size();
      if (in) { int32_t gi = in->group_of[(size_t)u]; if (gi >= 0) n_in += in->start[(size_t)gi + 1] - in->start[(size_t)gi]; }
    }
    info[1 + 4 * set] = (int64_t)tg.users.size();
    info[2 + 4 * set] = key_bytes;
    info[3 + 4 * set] = in ? n_in : (int64_t)tg.rows.size();       // train: its ratings are store 0
    info[4 + 4 * set] = in ? (int64_t)tg.rows.size() : 0;
  }
  return OCF_OK;
}

extern "C" int ocf_split_keys(const ocf_split* s, int set, char* bytes, int64_t* offsets) {
  if (!s || !bytes || !offsets || set < 0 || set > 2) return fail(OCF_ERR_INVALID, "ocf_split_keys: bad argument");
  const Groups& tg = split_targets(s, set);
  int64_t at = 0;
  for (size_t k = 0; k < tg.users.size(); ++k) {
    const std::string& key = s->sp->ukey[(size_t)tg.users[k]];
    offsets[k] = at;
    std::memcpy(bytes + at, key.data(), key.size());
    at += (int64_t)key.size();
  }
  offsets[tg.users.size()] = at;
  return OCF_OK;
}

extern "C" int ocf_split_csr(const ocf_split* s, int set, int part, int64_t* rowptr, int32_t* col, float* val, uint8_t* none) {
  if (!s || !rowptr || !col || !val || set < 0 || set > 2 || part < 0 || part > 1 || (set == 0 && part == 1))
    return fail(OCF_ERR_INVALID, "ocf_split_csr: bad argument");
  const Groups& tg = split_targets(s, set);
  const Groups* in = split_inputs(s, set);
  const Groups& src = (in && part == 0) ? *in : tg;
  const ocf_csv& csv = *s->csv;
  const int rc = s->sp->rating_col;
  const bool as_int = csv.col[rc].type == Column::INT;
  int64_t at = 0;
  for (size_t k = 0; k < tg.users.size(); ++k) {
    rowptr[k] = at;
    int32_t g = (in && part == 0) ? in->group_of[(size_t)tg.users[k]] : (int32_t)k;
    if (none) none[k] = g < 0 ? 1 : 0;
    if (g < 0) continue;
    for (int64_t j = src.start[(size_t)g]; j < src.start[(size_t)g + 1]; ++j) {
      const int64_t row = src.rows[(size_t)j];
      col[at] = s->item_dense[(size_t)row];
      val[at] = as_int ? (float)(double)csv.int_at(rc, row) : (float)csv.flt_at(rc, row);
      ++at;
    }
  }
  rowptr[tg.users.size()] = at;
  return OCF_OK;
}

// The column ids as the JSON list `unique_items_list.json` would hold (json.dump formatting): returns the text's
// length in *needed; copies it when cap is large enough.
extern "C" int ocf_split_columns_json(const ocf_split* s, char* buf, int64_t cap, int64_t* needed) {
  if (!s || !needed) return fail(OCF_ERR_INVALID, "ocf_split_columns_json: bad argument");
  const ocf_csv& csv = *s->csv;
  const int c = s->sp->item_col;
  const Column& col = csv.col[c];
  std::string out = "[";
  bool first = true;
  for (int64_t r : unique_first_rows(csv, c, nullptr)) {
    if (!first) out.append(", ");
    first = false;
    if (col.type == Column::STR) json_string(col.s[(size_t)r], &out);
    else if (col.type == Column::INT) out.append(std::to_string(csv.int_at(c, r)));
    else py_float_repr(csv.flt_at(c, r), &out, "NaN", "Infinity");
  }
  out.push_back(']');
  *needed = (int64_t)out.size();
  if (buf && cap >= (int64_t)out.size()) std::memcpy(buf, out.data(), out.size());
  return OCF_OK;
}

extern "C" int ocf_split_destroy(ocf_split* s) {
  delete s;
  return OCF_OK;
}
