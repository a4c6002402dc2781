// K1 — SimHash fingerprints (device MD5) + Hamming-distance scan for near-duplicate detection.
//
// Bit-exact with the reference's pure-Python implementation (infomesh/crawler/simhash.py:43-96):
//   words  = re.findall(r"\w+", text.lower())            (done by the host tokenizer, csrc/host/textproc.cpp)
//   shingle i = " ".join(words[i:i+3])  (one shingle of all words when fewer than 3; none -> fingerprint 0)
//   h = int.from_bytes(md5(shingle.encode("utf-8")).digest()[:8], "big")
//   bit b of the fingerprint is set iff  sum over shingles of (+1 if h bit b else -1)  >= 0
// The host hands over the normalised text (lower-cased words joined by single spaces) so that every shingle is
// one contiguous byte range [word_start[i], word_end[i+2]).  One warp per document: lanes hash shingles in
// parallel, votes are tallied with 64 ballots per 32 shingles, lane L keeps the counters of bits L and L+32.
//
// hamming_scan: every probe fingerprint against the whole fingerprint table (reference SimHashIndex.find_near,
// infomesh/crawler/simhash.py:186-205 is a linear Python scan): 128-bit loads, __popcll, per-probe packed
// atomicMin of (distance << 32 | index).
#include <cuda_bf16.h>

#include "../common/host.h"
#include "../common/ptx.cuh"

namespace im {

__device__ __constant__ uint32_t kMd5K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
__device__ __constant__ uint8_t kMd5S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                             5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                             4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                             6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

__device__ __forceinline__ void md5_block(uint32_t (&st)[4], const uint32_t (&m)[16]) {
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    uint32_t f;
    int g;
    if (i < 16) {
      f = (b & c) | (~b & d);
      g = i;
    } else if (i < 32) {
      f = (d & b) | (~d & c);
      g = (5 * i + 1) & 15;
    } else if (i < 48) {
      f = b ^ c ^ d;
      g = (3 * i + 5) & 15;
    } else {
      f = c ^ (b | ~d);
      g = (7 * i) & 15;
    }
    f = f + a + kMd5K[i] + m[g];
    a = d;
    d = c;
    c = b;
    b = b + __funnelshift_l(f, f, kMd5S[i]);
  }
  st[0] += a;
  st[1] += b;
  st[2] += c;
  st[3] += d;
}

// first 8 digest bytes interpreted big-endian
__device__ __forceinline__ uint64_t md5_first8_be(const uint8_t* __restrict__ data, int len) {
  uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  const int total = ((len + 8) / 64 + 1) * 64;  // padded length
  for (int off = 0; off < total; off += 64) {
    uint32_t m[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      uint32_t word = 0;
#pragma unroll
      for (int bidx = 0; bidx < 4; ++bidx) {
        const int p = off + w * 4 + bidx;
        uint32_t byte = 0;
        if (p < len) byte = data[p];
        else if (p == len) byte = 0x80u;
        word |= byte << (8 * bidx);
      }
      m[w] = word;
    }
    if (off + 64 == total) {
      const uint64_t bits = static_cast<uint64_t>(len) * 8ull;
      m[14] = static_cast<uint32_t>(bits);
      m[15] = static_cast<uint32_t>(bits >> 32);
    }
    md5_block(st, m);
  }
  return (static_cast<uint64_t>(__byte_perm(st[0], 0, 0x0123)) << 32) | __byte_perm(st[1], 0, 0x0123);
}

__global__ void __launch_bounds__(128)
simhash_kernel(const uint8_t* __restrict__ text, const long long* __restrict__ word_start,
               const long long* __restrict__ word_end, const long long* __restrict__ doc_word_off, int n_docs,
               int width, unsigned long long* __restrict__ out) {
  const int doc = blockIdx.x * 4 + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (doc >= n_docs) return;
  const long long w0 = doc_word_off[doc], w1 = doc_word_off[doc + 1];
  const long long nw = w1 - w0;
  if (nw <= 0) {
    if (lane == 0) out[doc] = 0ull;
    return;
  }
  const long long n_sh = nw < width ? 1 : nw - width + 1;
  int cnt_lo = 0, cnt_hi = 0;  // counters of bits `lane` and `lane + 32`
  for (long long base = 0; base < n_sh; base += 32) {
    const long long i = base + lane;
    const bool valid = i < n_sh;
    uint64_t h = 0;
    if (valid) {
      const long long first = w0 + i;
      const long long last = nw < width ? w1 - 1 : first + width - 1;
      const long long s = word_start[first], e = word_end[last];
      h = md5_first8_be(text + s, static_cast<int>(e - s));
    }
    const int n_valid = __popc(__ballot_sync(0xffffffffu, valid));
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      const int ones_lo = __popc(__ballot_sync(0xffffffffu, valid && ((h >> b) & 1ull)));
      const int ones_hi = __popc(__ballot_sync(0xffffffffu, valid && ((h >> (b + 32)) & 1ull)));
      if (static_cast<int>(lane) == b) {
        cnt_lo += 2 * ones_lo - n_valid;
        cnt_hi += 2 * ones_hi - n_valid;
      }
    }
  }
  const uint32_t lo = __ballot_sync(0xffffffffu, cnt_lo >= 0);
  const uint32_t hi = __ballot_sync(0xffffffffu, cnt_hi >= 0);
  if (lane == 0) out[doc] = (static_cast<unsigned long long>(hi) << 32) | lo;
}

// best[probe] = min over table of (hamming << 32 | index); callers pre-fill with 0xFFFFFFFFFFFFFFFF.
constexpr int kScanProbes = 256;  // probes staged in smem per launch chunk
__global__ void __launch_bounds__(256)
hamming_scan_kernel(const unsigned long long* __restrict__ table, long long n_table, long long table_index_base,
                    const unsigned long long* __restrict__ probes, int n_probes, int threshold,
                    unsigned long long* __restrict__ best, const long long* __restrict__ n_table_dev) {
  __shared__ unsigned long long sp[kScanProbes];
  if (n_table_dev != nullptr) n_table = min(n_table, *n_table_dev);   // shard fill level kept on the device (no host sync per batch)
  const int p0 = blockIdx.y * kScanProbes;
  const int np = min(kScanProbes, n_probes - p0);
  for (int i = threadIdx.x; i < np; i += blockDim.x) sp[i] = probes[p0 + i];
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x * 2;
  for (long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < n_table; i += stride) {
    unsigned long long t0, t1 = 0;
    bool has1 = i + 1 < n_table;
    if (has1) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(table + i);  // table is 16-byte aligned, i even
      t0 = v.x;
      t1 = v.y;
    } else {
      t0 = table[i];
    }
    for (int p = 0; p < np; ++p) {
      const unsigned long long pr = sp[p];
      const int d0 = __popcll(pr ^ t0);
      if (d0 <= threshold)
        atomicMin(best + p0 + p, (static_cast<unsigned long long>(d0) << 32) |
                                     static_cast<unsigned long long>((table_index_base + i) & 0xffffffffll));
      if (has1) {
        const int d1 = __popcll(pr ^ t1);
        if (d1 <= threshold)
          atomicMin(best + p0 + p, (static_cast<unsigned long long>(d1) << 32) |
                                       static_cast<unsigned long long>((table_index_base + i + 1) & 0xffffffffll));
      }
    }
  }
}

// Resolve one index-build batch without touching the host (K1, BASELINE config 5):
//   keep[j] = no rank found an already-indexed near-duplicate of passage j   (min over the gathered per-rank scan results)
//             AND no EARLIER passage i < j of this batch is within `threshold` bits (all pairs, smem-tiled popcounts).
// all_fp: [n] fingerprints of the whole batch in global order; best_parts: [P][n] packed scan results (all-ones = none).
__global__ void __launch_bounds__(256)
dedup_resolve_kernel(const unsigned long long* __restrict__ all_fp, const unsigned long long* __restrict__ best_parts, int P, int n,
                     int threshold, uint8_t* __restrict__ keep) {
  __shared__ unsigned long long tile[256];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long mine = j < n ? all_fp[j] : 0ull;
  bool dup = false;
  if (j < n)
    for (int p = 0; p < P; ++p) dup |= best_parts[static_cast<size_t>(p) * n + j] != ~0ull;
  const int j_max = min(n, (blockIdx.x + 1) * static_cast<int>(blockDim.x));   // only passages before this block's last row matter
  for (int base = 0; base < j_max; base += 256) {
    const int i = base + threadIdx.x;
    tile[threadIdx.x] = i < n ? all_fp[i] : 0ull;
    __syncthreads();
    const int lim = min(256, j - base);          // i < j
    for (int t = 0; t < lim; ++t) dup |= __popcll(mine ^ tile[t]) <= threshold;
    __syncthreads();
  }
  if (j < n) keep[j] = dup ? 0 : 1;
}

// Append this rank's surviving passages to its shard: exclusive scan of keep[slice] (one block), rows written at the
// device-side fill level, which is then advanced.  Rows that would overflow the shard are dropped and counted.
__global__ void __launch_bounds__(1024)
dedup_append_kernel(const uint8_t* __restrict__ keep, int bpr, const __nv_bfloat16* __restrict__ emb, int H,
                    const unsigned long long* __restrict__ fp, long long first_id, __nv_bfloat16* __restrict__ vectors,
                    unsigned long long* __restrict__ fingerprints, long long* __restrict__ doc_ids, long long* __restrict__ n_dev,
                    long long capacity, long long* __restrict__ counters /* {seen, dups, overflow} */) {
  __shared__ int warp_tot[32];
  __shared__ int pos_of[8192];
  __shared__ long long base_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int run = 0;   // rows kept in earlier 1024-row chunks
  for (int c0 = 0; c0 < bpr; c0 += 1024) {
    const int r = c0 + tid;
    const int k = (r < bpr && keep[r]) ? 1 : 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, k);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < wid) before += warp_tot[w];
      total += warp_tot[w];
    }
    if (r < bpr) pos_of[r] = k ? run + before + in_warp : -1;
    run += total;
    __syncthreads();
  }
  if (tid == 0) {
    base_s = *n_dev;
    const long long room = capacity - base_s;
    const long long take = run < room ? run : (room > 0 ? room : 0);
    *n_dev = base_s + take;
    counters[0] += bpr;
    counters[1] += bpr - run;
    counters[2] += run - take;
  }
  __syncthreads();
  const long long base = base_s;
  for (int r = wid; r < bpr; r += 32) {
    const int pos = pos_of[r];
    if (pos < 0 || base + pos >= capacity) continue;
    const uint4* src = reinterpret_cast<const uint4*>(emb + static_cast<size_t>(r) * H);
    uint4* dst = reinterpret_cast<uint4*>(vectors + static_cast<size_t>(base + pos) * H);
    for (int c = lane; c < H / 8; c += 32) dst[c] = src[c];
    if (lane == 0) {
      fingerprints[base + pos] = fp[r];
      doc_ids[base + pos] = first_id + r;
    }
  }
}

}  // namespace im

IM_API int im_simhash(const uint8_t* text, const long long* word_start, const long long* word_end,
                      const long long* doc_word_off, int n_docs, int width, unsigned long long* out, void* stream) {
  using namespace im;
  if (n_docs <= 0) return 0;
  simhash_kernel<<<(n_docs + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      text, word_start, word_end, doc_word_off, n_docs, width, out);
  IM_LAUNCH_OK("simhash_kernel");
  return 0;
}

static int hamming_scan_impl(const unsigned long long* table, long long n_table, long long table_index_base,
                             const unsigned long long* probes, int n_probes, int threshold, unsigned long long* best,
                             void* stream, const long long* n_table_dev) {
  using namespace im;
  if (n_table <= 0 || n_probes <= 0) return 0;
  long long blocks = (n_table / 2 + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid(static_cast<unsigned>(blocks), (n_probes + kScanProbes - 1) / kScanProbes);
  hamming_scan_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(table, n_table, table_index_base,
                                                                                probes, n_probes, threshold, best, n_table_dev);
  IM_LAUNCH_OK("hamming_scan_kernel");
  return 0;
}
IM_API int im_hamming_scan(const unsigned long long* table, long long n_table, long long table_index_base,
                           const unsigned long long* probes, int n_probes, int threshold, unsigned long long* best,
                           void* stream) {
  return hamming_scan_impl(table, n_table, table_index_base, probes, n_probes, threshold, best, stream, nullptr);
}
// Same, scanning only the first min(n_table, *n_table_dev) entries (the shard's fill level lives on the device).
IM_API int im_hamming_scan_dev(const unsigned long long* table, long long n_table, long long table_index_base,
                               const unsigned long long* probes, int n_probes, int threshold, unsigned long long* best,
                               const long long* n_table_dev, void* stream) {
  return hamming_scan_impl(table, n_table, table_index_base, probes, n_probes, threshold, best, stream, n_table_dev);
}

IM_API int im_dedup_resolve(const unsigned long long* all_fp, const unsigned long long* best_parts, int P, int n, int threshold,
                            uint8_t* keep, void* stream) {
  using namespace im;
  if (n <= 0) return 0;
  dedup_resolve_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(all_fp, best_parts, P, n, threshold, keep);
  IM_LAUNCH_OK("dedup_resolve_kernel");
  return 0;
}

IM_API int im_dedup_append(const uint8_t* keep, int bpr, const void* emb, int H, const unsigned long long* fp, long long first_id,
                           void* vectors, unsigned long long* fingerprints, long long* doc_ids, long long* n_dev, long long capacity,
                           long long* counters, void* stream) {
  using namespace im;
  if (bpr <= 0) return 0;
  if (bpr > 8192) return set_error("im_dedup_append", "at most 8192 passages per rank and batch");
  if (H % 8) return set_error("im_dedup_append", "H must be a multiple of 8");
  dedup_append_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      keep, bpr, reinterpret_cast<const __nv_bfloat16*>(emb), H, fp, first_id, reinterpret_cast<__nv_bfloat16*>(vectors), fingerprints,
      doc_ids, n_dev, capacity, counters);
  IM_LAUNCH_OK("dedup_append_kernel");
  return 0;
}
