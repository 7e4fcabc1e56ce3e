// Host side of the sparse input path: the reference's loader hands the
// model a DENSE fp32 bag-of-words matrix (Process/dataset.py:64-99 loads what
// Process/getTwittergraph.py:16-24,67-72 densified from `index:count` pairs), ~99.7 % zeros.
// Shipping it over PCIe costs 4*K bytes per node; bigcn_host_dense_to_csr makes one threaded
// pass over the host matrix and keeps the non-zero entries (8 bytes each), which is all the
// BIGCN_GEMM_SPARSE device path needs.  Pure data movement: no arithmetic happens here.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/bigcn_b200.h"

namespace {

struct Piece {
  std::vector<int32_t> col;
  std::vector<float> val;
  int64_t r0 = 0, r1 = 0;
};

// scalar scan of one row, columns ascending
inline void scan_row_scalar(const float* xr, int64_t k0, int64_t K, Piece& p) {
  for (int64_t k = k0; k < K; ++k) {
    const float v = xr[k];
    if (v != 0.0f) {
      p.col.push_back((int32_t)k);
      p.val.push_back(v);
    }
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void scan_rows_avx2(const float* x, int64_t K, Piece& p, int32_t* cnt) {
  const __m256 zero = _mm256_setzero_ps();
  for (int64_t r = p.r0; r < p.r1; ++r) {
    const float* xr = x + r * K;
    const size_t before = p.col.size();
    int64_t k = 0;
    // 32 floats (one 128 B line) per step: OR the four compare masks, look closer only on a hit
    for (; k + 32 <= K; k += 32) {
      _mm_prefetch(reinterpret_cast<const char*>(xr + k + 512), _MM_HINT_NTA);   // 2 KB ahead, no cache pollution
      _mm_prefetch(reinterpret_cast<const char*>(xr + k + 528), _MM_HINT_NTA);
      const __m256 a = _mm256_loadu_ps(xr + k), b = _mm256_loadu_ps(xr + k + 8);
      const __m256 c = _mm256_loadu_ps(xr + k + 16), d = _mm256_loadu_ps(xr + k + 24);
      const __m256 any = _mm256_or_ps(_mm256_or_ps(_mm256_cmp_ps(a, zero, _CMP_NEQ_UQ), _mm256_cmp_ps(b, zero, _CMP_NEQ_UQ)),
                                      _mm256_or_ps(_mm256_cmp_ps(c, zero, _CMP_NEQ_UQ), _mm256_cmp_ps(d, zero, _CMP_NEQ_UQ)));
      if (_mm256_movemask_ps(any)) scan_row_scalar(xr, k, k + 32, p);
    }
    if (k < K) scan_row_scalar(xr, k, K, p);
    cnt[r] = (int32_t)(p.col.size() - before);
  }
}
#endif

#if defined(__x86_64__)
// AVX-512: 64 floats (two 128 B lines) per step; the compare yields the non-zero mask directly and the
// hits leave through compress-stores (columns ascending), no scalar rescan
__attribute__((target("avx512f"))) void scan_rows_avx512(const float* x, int64_t K, Piece& p, int32_t* cnt) {
  const __m512 zero = _mm512_setzero_ps();
  const __m512i lane = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
  size_t n = p.col.size();
  auto room = [&](size_t extra) {          // compress-stores write in place: keep 64 free slots ahead
    if (p.col.size() < n + extra) {
      const size_t want = std::max(p.col.size() * 2, n + extra + 4096);
      p.col.resize(want);
      p.val.resize(want);
    }
  };
  for (int64_t r = p.r0; r < p.r1; ++r) {
    const float* xr = x + r * K;
    const size_t before = n;
    int64_t k = 0;
    for (; k + 64 <= K; k += 64) {
      _mm_prefetch(reinterpret_cast<const char*>(xr + k + 768), _MM_HINT_NTA);   // 3 KB ahead
      _mm_prefetch(reinterpret_cast<const char*>(xr + k + 784), _MM_HINT_NTA);
      _mm_prefetch(reinterpret_cast<const char*>(xr + k + 800), _MM_HINT_NTA);
      _mm_prefetch(reinterpret_cast<const char*>(xr + k + 816), _MM_HINT_NTA);
      const __m512 a = _mm512_loadu_ps(xr + k), b = _mm512_loadu_ps(xr + k + 16);
      const __m512 c = _mm512_loadu_ps(xr + k + 32), d = _mm512_loadu_ps(xr + k + 48);
      const __mmask16 ma = _mm512_cmp_ps_mask(a, zero, _CMP_NEQ_UQ), mb = _mm512_cmp_ps_mask(b, zero, _CMP_NEQ_UQ);
      const __mmask16 mc = _mm512_cmp_ps_mask(c, zero, _CMP_NEQ_UQ), md = _mm512_cmp_ps_mask(d, zero, _CMP_NEQ_UQ);
      if ((ma | mb | mc | md) == 0) continue;
      room(64);
      const __m512 v[4] = {a, b, c, d};
      const __mmask16 m[4] = {ma, mb, mc, md};
      for (int q = 0; q < 4; ++q) {
        if (!m[q]) continue;
        const __m512i cols = _mm512_add_epi32(lane, _mm512_set1_epi32((int)(k + 16 * q)));
        _mm512_mask_compressstoreu_epi32(p.col.data() + n, m[q], cols);
        _mm512_mask_compressstoreu_ps(p.val.data() + n, m[q], v[q]);
        n += (size_t)__builtin_popcount((unsigned)m[q]);
      }
    }
    for (; k < K; ++k) {
      const float vv = xr[k];
      if (vv != 0.0f) {
        room(1);
        p.col[n] = (int32_t)k;
        p.val[n] = vv;
        ++n;
      }
    }
    cnt[r] = (int32_t)(n - before);
  }
  p.col.resize(n);
  p.val.resize(n);
}
#endif

void scan_rows_plain(const float* x, int64_t K, Piece& p, int32_t* cnt) {
  for (int64_t r = p.r0; r < p.r1; ++r) {
    const size_t before = p.col.size();
    scan_row_scalar(x + r * K, 0, K, p);
    cnt[r] = (int32_t)(p.col.size() - before);
  }
}

}  // namespace

extern "C" int64_t bigcn_host_dense_to_csr(const float* x, int64_t N, int64_t K, int32_t* ptr, int32_t* col,
                                           float* val, int64_t cap, int32_t n_threads) {
  if (N < 0 || K <= 0 || !ptr || (N > 0 && !x)) return -1;
  int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if ((int64_t)T > N) T = (int)std::max<int64_t>(N, 1);
#if defined(__x86_64__)
  const bool avx2 = __builtin_cpu_supports("avx2");
  const bool avx512 = __builtin_cpu_supports("avx512f");
#else
  const bool avx2 = false;
#endif
  std::vector<Piece> pieces(T);
  // ptr[1..N] holds the per-row counts until the prefix pass
  int32_t* cnt = ptr + 1;
  auto work = [&](int t) {
    Piece& p = pieces[t];
    p.r0 = N * t / T;
    p.r1 = N * (t + 1) / T;
    const size_t guess = (size_t)(p.r1 - p.r0) * 24;
    p.col.reserve(guess);
    p.val.reserve(guess);
#if defined(__x86_64__)
    if (avx512) {
      scan_rows_avx512(x, K, p, cnt);
      return;
    }
    if (avx2) {
      scan_rows_avx2(x, K, p, cnt);
      return;
    }
#endif
    scan_rows_plain(x, K, p, cnt);
  };
  std::vector<std::thread> th;
  th.reserve(T);
  for (int t = 1; t < T; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& h : th) h.join();
  int64_t total = 0;
  std::vector<int64_t> base(T);
  for (int t = 0; t < T; ++t) {
    base[t] = total;
    total += (int64_t)pieces[t].col.size();
  }
  if (total > cap || total > 2147483647ll) return -total;
  ptr[0] = 0;
  // row offsets (serial prefix over N ints: ~10 us per 10^5 rows) and the threaded copy-out
  int64_t run = 0;
  for (int64_t r = 0; r < N; ++r) {
    run += cnt[r];
    ptr[r + 1] = (int32_t)run;
  }
  auto copy_out = [&](int t) {
    const Piece& p = pieces[t];
    if (p.col.empty()) return;
    memcpy(col + base[t], p.col.data(), p.col.size() * sizeof(int32_t));
    memcpy(val + base[t], p.val.data(), p.val.size() * sizeof(float));
  };
  th.clear();
  for (int t = 1; t < T; ++t) th.emplace_back(copy_out, t);
  copy_out(0);
  for (auto& h : th) h.join();
  return total;
}


// Host-memory read bandwidth as the compaction threads see it (a STREAM-style probe: `n_threads` threads each sum their
// slice of the buffer with 64-bit loads, `reps` passes; returns GB/s).  bench.py reports it beside the end-to-end
// numbers fed from dense pinned host matrices: that route cannot run faster than PCIe plus this number allow.
extern "C" double bigcn_host_read_gbs(const void* buf, int64_t bytes, int32_t n_threads, int32_t reps) {
  if (!buf || bytes < 8 || reps < 1) return 0.0;
  int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  const int64_t words = bytes / 8;
  std::vector<uint64_t> sink(T, 0);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) {
    th.emplace_back([&, t] {
      const uint64_t* p = reinterpret_cast<const uint64_t*>(buf);
      const int64_t lo = words * t / T, hi = words * (t + 1) / T;
      uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int r = 0; r < reps; ++r) {
        int64_t i = lo;
        for (; i + 4 <= hi; i += 4) {
          a0 += p[i]; a1 += p[i + 1]; a2 += p[i + 2]; a3 += p[i + 3];
        }
        for (; i < hi; ++i) a0 += p[i];
      }
      sink[t] = a0 + a1 + a2 + a3;
    });
  }
  for (auto& x : th) x.join();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  volatile uint64_t keep = 0;
  for (int t = 0; t < T; ++t) keep += sink[t];
  (void)keep;
  return sec > 0 ? (double)bytes * reps / sec / 1e9 : 0.0;
}
