// Host runtime (C++): tokeniser, term dictionary, posting-list (CSR) builder, SimHash/MD5 CPU oracle.
//
// This is the native replacement for the parts of SQLite FTS5 the reference leans on for indexing
// (tokenise -> inverted index, reference infomesh/index/local_store.py:103-143,198-251) and for the
// pure-Python SimHash (reference infomesh/crawler/simhash.py:43-96).  The GPU kernels consume the arrays
// produced here (csrc/search/bm25.cu, csrc/dedup/simhash.cu).
//
// Tokenisation rule (FTS5 unicode61-like, ASCII fast path): a token is a maximal run of ASCII letters/digits
// or non-ASCII bytes (so UTF-8 sequences stay inside tokens), ASCII upper-case folded to lower-case.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#define IM_API extern "C" __attribute__((visibility("default")))

namespace {

inline bool is_tok_byte(unsigned char c) {
  return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c >= 0x80;
}
inline bool is_word_byte_py(unsigned char c) {  // Python \w on ASCII: [A-Za-z0-9_]
  return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_';
}
inline unsigned char lower(unsigned char c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

struct Posting {
  int32_t doc;
  uint8_t tf;
};

struct IndexBuilder {
  std::unordered_map<std::string, int32_t> dict;
  std::vector<std::string> terms;
  std::vector<std::vector<Posting>> postings;
  std::vector<int32_t> doc_len;
  int64_t total_len = 0;
  bool frozen_vocab = false;

  int32_t term_id(const std::string& t, bool add) {
    auto it = dict.find(t);
    if (it != dict.end()) return it->second;
    if (!add) return -1;
    const int32_t id = static_cast<int32_t>(terms.size());
    dict.emplace(t, id);
    terms.push_back(t);
    postings.emplace_back();
    return id;
  }

  void tokenize(const char* text, int64_t len, std::vector<int32_t>& out, bool add) {
    std::string cur;
    for (int64_t i = 0; i <= len; ++i) {
      const unsigned char c = i < len ? static_cast<unsigned char>(text[i]) : 0;
      if (i < len && is_tok_byte(c)) {
        cur.push_back(static_cast<char>(lower(c)));
      } else if (!cur.empty()) {
        out.push_back(term_id(cur, add));
        cur.clear();
      }
    }
  }

  int32_t add_doc_terms(const int32_t* ids, int64_t n) {
    const int32_t doc = static_cast<int32_t>(doc_len.size());
    std::vector<int32_t> sorted(ids, ids + n);
    std::sort(sorted.begin(), sorted.end());
    int64_t valid = 0;
    for (size_t i = 0; i < sorted.size();) {
      size_t j = i;
      while (j < sorted.size() && sorted[j] == sorted[i]) ++j;
      const int32_t t = sorted[i];
      if (t >= 0) {
        if (static_cast<size_t>(t) >= postings.size()) postings.resize(t + 1);
        const size_t tf = j - i;
        postings[t].push_back({doc, static_cast<uint8_t>(tf > 255 ? 255 : tf)});
        valid += static_cast<int64_t>(tf);
      }
      i = j;
    }
    doc_len.push_back(static_cast<int32_t>(valid));
    total_len += valid;
    return doc;
  }
};

// ---------------------------------------------------------------- MD5 (RFC 1321) for the SimHash oracle
struct Md5 {
  static inline uint32_t rotl(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
  static void block(uint32_t st[4], const uint8_t* p) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
        0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
        0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
        0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
        0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
        0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
        0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
        0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                              14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                              4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t m[16];
    for (int i = 0; i < 16; ++i)
      m[i] = static_cast<uint32_t>(p[4 * i]) | (static_cast<uint32_t>(p[4 * i + 1]) << 8) |
             (static_cast<uint32_t>(p[4 * i + 2]) << 16) | (static_cast<uint32_t>(p[4 * i + 3]) << 24);
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
    for (int i = 0; i < 64; ++i) {
      uint32_t f;
      int g;
      if (i < 16) { f = (b & c) | (~b & d); g = i; }
      else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
      else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
      else { f = c ^ (b | ~d); g = (7 * i) & 15; }
      f = f + a + K[i] + m[g];
      a = d; d = c; c = b;
      b = b + rotl(f, S[i]);
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d;
  }
  static uint64_t first8_be(const uint8_t* data, int64_t len) {
    uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    int64_t off = 0;
    for (; off + 64 <= len; off += 64) block(st, data + off);
    uint8_t tail[128] = {0};
    const int64_t rem = len - off;
    if (rem > 0) std::memcpy(tail, data + off, static_cast<size_t>(rem));
    tail[rem] = 0x80;
    const int64_t tl = rem + 9 <= 64 ? 64 : 128;
    const uint64_t bits = static_cast<uint64_t>(len) * 8ull;
    for (int i = 0; i < 8; ++i) tail[tl - 8 + i] = static_cast<uint8_t>(bits >> (8 * i));
    block(st, tail);
    if (tl == 128) block(st, tail + 64);
    auto bswap = [](uint32_t x) {
      return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
    };
    return (static_cast<uint64_t>(bswap(st[0])) << 32) | bswap(st[1]);
  }
};

}  // namespace

// ---------------------------------------------------------------- index builder C API
IM_API void* im_ib_create() { return new IndexBuilder(); }
IM_API void im_ib_destroy(void* h) { delete static_cast<IndexBuilder*>(h); }
IM_API int im_ib_add_text(void* h, const char* text, long long len) {
  auto* b = static_cast<IndexBuilder*>(h);
  std::vector<int32_t> ids;
  b->tokenize(text, len, ids, true);
  return b->add_doc_terms(ids.data(), static_cast<int64_t>(ids.size()));
}
IM_API int im_ib_add_terms(void* h, const int* ids, long long n) {
  return static_cast<IndexBuilder*>(h)->add_doc_terms(ids, n);
}
// tokenise a query / document into term ids (no vocabulary growth unless add != 0); returns count (<= cap)
IM_API long long im_ib_tokenize(void* h, const char* text, long long len, int add, int* out, long long cap) {
  auto* b = static_cast<IndexBuilder*>(h);
  std::vector<int32_t> ids;
  b->tokenize(text, len, ids, add != 0);
  const long long n = static_cast<long long>(ids.size()) < cap ? static_cast<long long>(ids.size()) : cap;
  if (n > 0) std::memcpy(out, ids.data(), static_cast<size_t>(n) * sizeof(int32_t));   // data() may be null when empty
  return static_cast<long long>(ids.size());
}
IM_API int im_ib_lookup(void* h, const char* term, long long len) {
  auto* b = static_cast<IndexBuilder*>(h);
  std::string t(term, static_cast<size_t>(len));
  for (auto& c : t) c = static_cast<char>(lower(static_cast<unsigned char>(c)));
  return b->term_id(t, false);
}
IM_API long long im_ib_vocab(void* h) { return static_cast<long long>(static_cast<IndexBuilder*>(h)->postings.size()); }
IM_API long long im_ib_docs(void* h) { return static_cast<long long>(static_cast<IndexBuilder*>(h)->doc_len.size()); }
IM_API long long im_ib_nnz(void* h) {
  long long n = 0;
  for (auto& p : static_cast<IndexBuilder*>(h)->postings) n += static_cast<long long>(p.size());
  return n;
}
IM_API double im_ib_avg_len(void* h) {
  auto* b = static_cast<IndexBuilder*>(h);
  return b->doc_len.empty() ? 0.0 : static_cast<double>(b->total_len) / static_cast<double>(b->doc_len.size());
}
IM_API long long im_ib_term_bytes(void* h, int id, char* out, long long cap) {
  auto* b = static_cast<IndexBuilder*>(h);
  if (id < 0 || static_cast<size_t>(id) >= b->terms.size()) return -1;
  const std::string& t = b->terms[id];
  const long long n = static_cast<long long>(t.size()) < cap ? static_cast<long long>(t.size()) : cap;
  if (n > 0) std::memcpy(out, t.data(), static_cast<size_t>(n));
  return static_cast<long long>(t.size());
}
// Export CSR: off[V+1], doc[nnz], tf[nnz], doc_len[n_docs], df[V]
IM_API int im_ib_export(void* h, long long* off, int* doc, unsigned char* tf, int* doc_len, int* df) {
  auto* b = static_cast<IndexBuilder*>(h);
  long long pos = 0;
  const size_t V = b->postings.size();
  for (size_t t = 0; t < V; ++t) {
    off[t] = pos;
    for (const auto& p : b->postings[t]) {
      doc[pos] = p.doc;
      tf[pos] = p.tf;
      ++pos;
    }
    if (df) df[t] = static_cast<int>(b->postings[t].size());
  }
  off[V] = pos;
  if (!b->doc_len.empty()) std::memcpy(doc_len, b->doc_len.data(), b->doc_len.size() * sizeof(int32_t));
  return 0;
}

// ---------------------------------------------------------------- SimHash host side
// ASCII fast path of  " ".join(re.findall(r"\w+", text.lower())) : writes the normalised text and word offsets.
// Returns the number of words, or -1 if a non-ASCII byte is present (caller falls back to Python's regex),
// or -2 if a capacity is exceeded.  out must hold len + 1 bytes.
IM_API long long im_normalize_words_ascii(const char* text, long long len, char* out, long long* out_len,
                                          long long* word_start, long long* word_end, long long max_words) {
  long long n = 0, o = 0;
  bool in_word = false;
  for (long long i = 0; i < len; ++i) {
    const unsigned char c = static_cast<unsigned char>(text[i]);
    if (c >= 0x80) return -1;
    if (is_word_byte_py(c)) {
      if (!in_word) {
        if (n >= max_words) return -2;
        if (n > 0) out[o++] = ' ';
        word_start[n] = o;
        in_word = true;
      }
      out[o++] = static_cast<char>(lower(c));
    } else if (in_word) {
      word_end[n++] = o;
      in_word = false;
    }
  }
  if (in_word) word_end[n++] = o;
  *out_len = o;
  return n;
}

// CPU oracle / fallback: fingerprint of one normalised document
IM_API unsigned long long im_simhash_cpu(const char* text, const long long* word_start, const long long* word_end,
                                         long long n_words, int width) {
  if (n_words <= 0) return 0ull;
  const long long n_sh = n_words < width ? 1 : n_words - width + 1;
  int vec[64] = {0};
  for (long long i = 0; i < n_sh; ++i) {
    const long long last = n_words < width ? n_words - 1 : i + width - 1;
    const long long s = word_start[i], e = word_end[last];
    const uint64_t hsh = Md5::first8_be(reinterpret_cast<const uint8_t*>(text) + s, e - s);
    for (int b = 0; b < 64; ++b) vec[b] += ((hsh >> b) & 1ull) ? 1 : -1;
  }
  unsigned long long fp = 0;
  for (int b = 0; b < 64; ++b)
    if (vec[b] >= 0) fp |= (1ull << b);
  return fp;
}

IM_API unsigned long long im_md5_first8_be(const char* data, long long len) {
  return Md5::first8_be(reinterpret_cast<const uint8_t*>(data), len);
}

// nearest fingerprint within `threshold`: returns index or -1 (CPU fallback of hamming_scan)
IM_API long long im_hamming_find_cpu(const unsigned long long* table, long long n, unsigned long long probe,
                                     int threshold) {
  long long best = -1;
  int bd = threshold + 1;
  for (long long i = 0; i < n; ++i) {
    const int d = __builtin_popcountll(table[i] ^ probe);
    if (d < bd) {
      bd = d;
      best = i;
      if (d == 0) break;
    }
  }
  return best;
}
