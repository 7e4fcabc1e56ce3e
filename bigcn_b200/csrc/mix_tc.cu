// The 64 x 64 contractions of conv2.lin and of its backward on the tensor cores (tcgen05 + TMEM).
//
//   forward  (BiGCN_Twitter.py:51-56): Z  = A1 W2a^T + dropout(relu(x_root[batch])) W2b^T     A1 = dropout(relu(H1))
//   backward                         : G1 = (T2 W2a) * mask * [H1 > 0],  db1 partial sums
//
// Both are [N, 64] x [64, 64] products: tiny per row, but at one half-warp per row (k_prop1_mix, k_bwd_mix) the 256
// FFMA instructions per row are what those kernels issue most.  Here a CTA owns 128 rows: TMA brings the A tile
// (A1 or T2, L2 resident) and the 64 x 64 weight tile (hi / lo split) into 128B-swizzled shared memory, all eight
// warps split the A tile into hi + lo in place (fp32-class split TF32: A_hi W_hi + A_hi W_lo + A_lo W_hi + A_lo W_lo), one thread
// issues the 32 tcgen05.mma (M = 128, N = 64, K = 8; all four hi / lo products) into a 64-column TMEM accumulator, and the epilogue -- one
// THREAD per row and 32-column half, straight out of TMEM -- adds what is row-specific:
//   forward : the root part.  Per node, one Philox draw per non-zero root column of its tree decides which columns
//             survive; the surviving rows of W2b^T (L1 / L2 resident, the same for every node of a tree) are added.
//             Also hands the keep decisions (one 64-bit word per node) to the backward.
//   backward: the relu / dropout mask taken from A1 (> 0 exactly where the forward kept a positive H1), the store of
//             G1 and the tile's column sums for db1.
// k_prop1_act is the sweep that feeds the forward product: conv1 propagate + bias + relu + dropout (k_prop1_mix
// without its matvec and root part).
#include "gather.cuh"
#include "tc.cuh"

namespace bigcn {

constexpr int HT_M = 128;                         // rows per CTA
constexpr int HT_A_BYTES = HT_M * 64 * 4;         // 32 KB: two k-blocks of [128 rows][128 B]
constexpr int HT_B_BYTES = 64 * 64 * 4;           // 16 KB: two k-blocks of [64 rows][128 B]
constexpr int HT_SMEM = 2 * HT_A_BYTES + 2 * HT_B_BYTES + 1024;

struct H64Dir {
  float* out;                    // Z (forward) / G1 (backward)  [N][64]
  // forward epilogue
  const float* w2bT;             // [K][64]
  const float* P;                // [B][64] eval-mode root projection
  DropSpec drop;
  unsigned long long* keep;      // [N] or NULL
  // backward epilogue
  const float* a1;               // [N][64]
  float* part;                   // [tiles][64] column sums of G1
};
struct H64Args {
  H64Dir d[2];
  int64_t N, K, node_id_base;
  const int64_t* batch;
  const int32_t* rnz_cnt;
  const int32_t* rnz_col;
  const float* rnz_val;
};

enum { EPI_MIX = 0, EPI_BWD = 1 };

template <int EPI>
__global__ void __launch_bounds__(256, 2)
k_h64_tc(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
         const __grid_constant__ CUtensorMap map_bh0, const __grid_constant__ CUtensorMap map_bh1,
         const __grid_constant__ CUtensorMap map_bl0, const __grid_constant__ CUtensorMap map_bl1, const H64Args p) {
  extern __shared__ __align__(1024) uint8_t ht_smem[];
  __shared__ __align__(8) uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[4][H];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.y;
  const CUtensorMap* map_a = dir == 0 ? &map_a0 : &map_a1;
  const CUtensorMap* map_bh = dir == 0 ? &map_bh0 : &map_bh1;
  const CUtensorMap* map_bl = dir == 0 ? &map_bl0 : &map_bl1;
  const uint32_t smem0 = (smem_u32(ht_smem) + 1023u) & ~1023u;
  const uint32_t sA = smem0, sAlo = smem0 + HT_A_BYTES, sBh = smem0 + 2 * HT_A_BYTES, sBl = sBh + HT_B_BYTES;
  const int64_t m0 = (int64_t)blockIdx.x * HT_M;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(map_a);
    tma_prefetch_desc(map_bh);
    tma_prefetch_desc(map_bl);
    mbar_init(smem_u32(&bar_full), 1);
    mbar_init(smem_u32(&bar_mma), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0 && lane == 0) {   // TMA: the A tile and the weight tiles, two k-blocks of 32 columns each
    const uint32_t full = smem_u32(&bar_full);
    mbar_expect_tx(full, (uint32_t)(HT_A_BYTES + 2 * HT_B_BYTES));
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_2d(sA + kb * (HT_A_BYTES / 2), map_a, full, kb * 32, (int)m0);
      tma_load_2d(sBh + kb * (HT_B_BYTES / 2), map_bh, full, kb * 32, 0);
      tma_load_2d(sBl + kb * (HT_B_BYTES / 2), map_bl, full, kb * 32, 0);
    }
  }
  mbar_wait(smem_u32(&bar_full), 0);
  // every thread: its share of the A tile -> hi (in place) | lo (sibling buffer, same swizzle)
  for (int off = threadIdx.x * 16; off < HT_A_BYTES; off += 256 * 16) {
    uint32_t a, b, c, d;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sA + off));
    const uint32_t ha = a & 0xFFFFE000u, hb = b & 0xFFFFE000u, hc = c & 0xFFFFE000u, hd = d & 0xFFFFE000u;
    const uint32_t la = tf32_lo_bits(a, ha), lb = tf32_lo_bits(b, hb), lc = tf32_lo_bits(c, hc), ld = tf32_lo_bits(d, hd);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sA + off), "r"(ha), "r"(hb), "r"(hc), "r"(hd) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(sAlo + off), "r"(la), "r"(lb), "r"(lc), "r"(ld) : "memory");
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (warp == 1 && lane == 0) {   // MMA issuer
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(HT_M, 64);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t koff = k * 32;   // 8 fp32 per K step inside the 128 B atom
        const uint64_t da = make_desc_k_sw128(sA + kb * (HT_A_BYTES / 2) + koff);
        const uint64_t dl = make_desc_k_sw128(sAlo + kb * (HT_A_BYTES / 2) + koff);
        const uint64_t bh = make_desc_k_sw128(sBh + kb * (HT_B_BYTES / 2) + koff);
        const uint64_t bl = make_desc_k_sw128(sBl + kb * (HT_B_BYTES / 2) + koff);
        tc_mma_tf32(tmem_base, da, bh, idesc, (kb | k) ? 1u : 0u);
        tc_mma_tf32(tmem_base, da, bl, idesc, 1u);
        tc_mma_tf32(tmem_base, dl, bh, idesc, 1u);
        tc_mma_tf32(tmem_base, dl, bl, idesc, 1u);   // the 2^-20 term: free at this size, and db1 sums thousands of rows
      }
    }
    tc_commit(smem_u32(&bar_mma));
  }
  mbar_wait(smem_u32(&bar_mma), 0);
  tc_fence_after();

  // ---- epilogue: thread = (row, 32-column half) -------------------------------------------------------------
  const int q = warp & 3, c0 = (warp >> 2) * 32;
  const int64_t i = m0 + q * 32 + lane;
  const H64Dir& dd = p.d[dir];
  uint32_t r[32];
  tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
  tc_wait_ld();
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(r[j]);
  const bool live = i < p.N;
  if (EPI == EPI_MIX) {
    if (live) {
      const int64_t b = p.batch[i];
      if (dd.drop.on) {
        const int n = p.rnz_cnt[b];
        const int64_t node = p.node_id_base + i;
        const int32_t* rc = p.rnz_col + b * p.K;
        const float* rv = p.rnz_val + b * p.K;
        unsigned long long kept = 0ull;   // bit t = root slot t survived for this node (slots 0..63)
        for (int t = 0; t < n; ++t) {
          const int k = rc[t];
          const uint32_t c = (uint32_t)(H + k);
          const Philox4 ph = drop_block(dd.drop, node, c >> 2);
          const bool keep = philox_elem(ph, c & 3) >= dd.drop.thresh;
          if (keep && t < 64) kept |= 1ull << t;
          const float s = keep ? __fmul_rn(rv[t], dd.drop.scale) : 0.f;
          const float4* wr = reinterpret_cast<const float4*>(dd.w2bT + (int64_t)k * H + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = __ldg(wr + j);
            acc[4 * j + 0] = fmaf(s, w.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(s, w.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(s, w.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(s, w.w, acc[4 * j + 3]);
          }
        }
        if (c0 == 0 && dd.keep != nullptr) dd.keep[i] = kept;
      } else {
        const float4* pr = reinterpret_cast<const float4*>(dd.P + b * H + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w = pr[j];
          acc[4 * j + 0] += w.x; acc[4 * j + 1] += w.y; acc[4 * j + 2] += w.z; acc[4 * j + 3] += w.w;
        }
      }
      float4* dst = reinterpret_cast<float4*>(dd.out + i * H + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    }
  } else {
    const float sc = dd.drop.on ? dd.drop.scale : 1.f;   // x 1.0f is exact
    if (live) {
      const float4* ar = reinterpret_cast<const float4*>(dd.a1 + i * H + c0);
      float4* dst = reinterpret_cast<float4*>(dd.out + i * H + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 h = ar[j];
        acc[4 * j + 0] = h.x > 0.f ? __fmul_rn(acc[4 * j + 0], sc) : 0.f;
        acc[4 * j + 1] = h.y > 0.f ? __fmul_rn(acc[4 * j + 1], sc) : 0.f;
        acc[4 * j + 2] = h.z > 0.f ? __fmul_rn(acc[4 * j + 2], sc) : 0.f;
        acc[4 * j + 3] = h.w > 0.f ? __fmul_rn(acc[4 * j + 3], sc) : 0.f;
        dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    }
    // column sums of the tile: butterfly over the 32 rows of the warp, then the four row quarters in order
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
      if (lane == 0) red[q][c0 + j] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (EPI == EPI_BWD && threadIdx.x < H)
    dd.part[(int64_t)blockIdx.x * H + threadIdx.x] =
        ((red[0][threadIdx.x] + red[1][threadIdx.x]) + red[2][threadIdx.x]) + red[3][threadIdx.x];
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base));
  }
}

// hi / lo split of the conv2 weights' hidden block W2a = W2[:, :64] ([o][k], row pitch ld) and of its transpose:
// out[d] = { W2a_hi, W2a_lo, W2a^T_hi, W2a^T_lo }, 4096 floats each
struct W2aSplitArgs {
  const float* w2[2];
  float* out[2];
  int64_t ld;
};
__global__ void __launch_bounds__(256) k_w2a_split(W2aSplitArgs a) {
  const int d = blockIdx.y;
  const int idx = blockIdx.x * 256 + threadIdx.x;   // o * 64 + k
  if (idx >= H * H) return;
  const int o = idx >> 6, k = idx & 63;
  const float v = a.w2[d][(int64_t)o * a.ld + k];
  const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  const float l = __uint_as_float((__float_as_uint(v - h) + 0x1000u) & 0xFFFFE000u);
  float* out = a.out[d];
  out[idx] = h;
  out[H * H + idx] = l;
  out[2 * H * H + k * H + o] = h;
  out[3 * H * H + k * H + o] = l;
}

// ---- conv1 propagate + bias + relu + dropout: H1, A1 (the sweep half of the forward mix) -----------------------
struct PostAct {
  static constexpr bool kPairs = false;
  static constexpr bool kInline = true;
  float* h1;
  float* a1;
  DropSpec drop;
  int64_t node_id_base;
  float4 b1;
  __device__ __forceinline__ void operator()(int i, float4 v, int sub, unsigned, float*) const {
    v.x = __fadd_rn(v.x, b1.x);
    v.y = __fadd_rn(v.y, b1.y);
    v.z = __fadd_rn(v.z, b1.z);
    v.w = __fadd_rn(v.w, b1.w);
    st4(h1 + (int64_t)i * H + 4 * sub, v);
    float4 av = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    if (drop.on) {
      const Philox4 r = drop_block(drop, node_id_base + i, (uint32_t)sub);
      av.x = r.x >= drop.thresh ? __fmul_rn(av.x, drop.scale) : 0.f;
      av.y = r.y >= drop.thresh ? __fmul_rn(av.y, drop.scale) : 0.f;
      av.z = r.z >= drop.thresh ? __fmul_rn(av.z, drop.scale) : 0.f;
      av.w = r.w >= drop.thresh ? __fmul_rn(av.w, drop.scale) : 0.f;
    }
    st4(a1 + (int64_t)i * H + 4 * sub, av);
  }
};
constexpr int ACT_R = 8, ACT_Q = 8;
__global__ void __launch_bounds__(256, 3) k_prop1_act(MixArgs a) {
  extern __shared__ __align__(128) float sweep_smem[];
  const MixDir p = a.d[blockIdx.y];
  const int sub = threadIdx.x & 15;
  PostAct post{p.h1, p.a1, p.drop, a.node_id_base, ld4(p.b1 + 4 * sub)};
  csr_sweep<ACT_R, ACT_Q, true>(Csr{p.ptr, p.idx, p.lng, p.E}, WtGcn{p.dis}, (int)a.N, a.cb, sweep_smem,
                                ValRow{p.xw, a.ldxw}, post);
}

bool mix_tc_available() { return encode_fn() != nullptr; }
size_t mix_tc_scratch_floats() { return (size_t)2 * 4 * H * H; }

static void mix_tc_attrs() {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_prop1_act, cudaFuncAttributeMaxDynamicSharedMemorySize, SweepSmem<ACT_R, ACT_Q>::kBytes);
    cudaFuncSetAttribute(k_h64_tc<EPI_MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM);
    cudaFuncSetAttribute(k_h64_tc<EPI_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM);
    attr = true;
  }
}

// scratch (mix_tc_scratch_floats() floats) <- hi / lo of W2a and W2a^T of the active directions: once per forward
int mix_tc_split_weights(const float* const* w2, int ndir, int64_t ldw2, float* scratch, cudaStream_t st) {
  mix_tc_attrs();
  W2aSplitArgs sp{};
  sp.ld = ldw2;
  for (int q = 0; q < ndir; ++q) {
    sp.w2[q] = w2[q];
    sp.out[q] = scratch + (size_t)q * 4 * H * H;
  }
  k_w2a_split<<<dim3(16, ndir), 256, 0, st>>>(sp);
  BIGCN_CHECK_LAUNCH("k_w2a_split");
  return 0;
}

// scratch: as left by mix_tc_split_weights
int mix_tc_forward(const MixArgs& a0, int ndir, float* scratch, cudaStream_t st) {
  if (a0.N == 0) return 0;
  BIGCN_CHECK_ARG(encode_fn() != nullptr, "mix_tc: cuTensorMapEncodeTiled is unavailable in this driver");
  mix_tc_attrs();
  struct { float* out[2]; } sp;
  for (int q = 0; q < 2; ++q) sp.out[q] = scratch + (size_t)q * 4 * H * H;
  MixArgs a = a0;
  const int max_ctas = num_sms() * 3;
  a.cb = sweep_cb(a.N, ACT_R, max_ctas);
  k_prop1_act<<<dim3(sweep_grid(a.N, ACT_R, a.cb, max_ctas), ndir), 256, SweepSmem<ACT_R, ACT_Q>::kBytes, st>>>(a);
  BIGCN_CHECK_LAUNCH("k_prop1_act");
  CUtensorMap ma[2], mh[2], ml[2];
  H64Args h{};
  h.N = a.N; h.K = a.K; h.node_id_base = a.node_id_base; h.batch = a.batch;
  h.rnz_cnt = a.rnz_cnt; h.rnz_col = a.rnz_col; h.rnz_val = a.rnz_val;
  for (int q = 0; q < ndir; ++q) {
    const MixDir& m = a.d[q];
    if (int rc = make_map(&ma[q], m.a1, a.N, H, H, HT_M)) return rc;
    if (int rc = make_map(&mh[q], sp.out[q], H, H, H, H)) return rc;
    if (int rc = make_map(&ml[q], sp.out[q] + H * H, H, H, H, H)) return rc;
    h.d[q].out = m.z; h.d[q].w2bT = m.w2bT; h.d[q].P = m.P; h.d[q].drop = m.drop; h.d[q].keep = m.keep;
  }
  if (ndir == 1) { ma[1] = ma[0]; mh[1] = mh[0]; ml[1] = ml[0]; }
  k_h64_tc<EPI_MIX><<<dim3((unsigned)ceil_div(a.N, HT_M), ndir), 256, HT_SMEM, st>>>(ma[0], ma[1], mh[0], mh[1], ml[0], ml[1], h);
  BIGCN_CHECK_LAUNCH("k_h64_tc<mix>");
  return 0;
}

// G1 = (T2 W2a) * mask; part[tile][64] = column sums per 128-row tile (BM_ROWS).  scratch: as left by mix_tc_split_weights
int mix_tc_backward(const BwdMixArgs& a, int ndir, const float* scratch, cudaStream_t st) {
  if (a.N == 0) return 0;
  mix_tc_attrs();
  static_assert(BM_ROWS == HT_M, "db1 partials: one per 128-row tile");
  CUtensorMap ma[2], mh[2], ml[2];
  H64Args h{};
  h.N = a.N;
  for (int q = 0; q < ndir; ++q) {
    const BwdMixDir& m = a.d[q];
    const float* s = scratch + (size_t)q * 4 * H * H;
    if (int rc = make_map(&ma[q], m.t2, a.N, H, H, HT_M)) return rc;
    if (int rc = make_map(&mh[q], s + 2 * H * H, H, H, H, H)) return rc;   // B[n = k][kdim = o] = W2a^T
    if (int rc = make_map(&ml[q], s + 3 * H * H, H, H, H, H)) return rc;
    h.d[q].out = m.g1; h.d[q].a1 = m.h1; h.d[q].part = m.part; h.d[q].drop = m.drop;
  }
  if (ndir == 1) { ma[1] = ma[0]; mh[1] = mh[0]; ml[1] = ml[0]; }
  k_h64_tc<EPI_BWD><<<dim3((unsigned)ceil_div(a.N, HT_M), ndir), 256, HT_SMEM, st>>>(ma[0], ma[1], mh[0], mh[1], ml[0], ml[1], h);
  BIGCN_CHECK_LAUNCH("k_h64_tc<bwd>");
  return 0;
}

}  // namespace bigcn
