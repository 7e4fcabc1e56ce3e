// Head (fc + log_softmax), nll loss and Adam: the small fp32 pieces around the path.
// Replaces BiGCN_Twitter.py:129-130 (fc, log_softmax), :184 (F.nll_loss) and the
// torch.optim.Adam step configured at :146-153.
#include "kernels.cuh"

namespace bigcn {

constexpr int FEAT = 4 * H;  // 256 = (out_feats + hid_feats) * 2

// warp per tree; lane c holds logit c (C <= 32)
__global__ void __launch_bounds__(256) k_head_fwd(const float* __restrict__ feat, int64_t B, int C,
                                                  const float* __restrict__ W,
                                                  const float* __restrict__ bias,
                                                  float* __restrict__ logp) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float fv[FEAT / 32];
#pragma unroll
  for (int j = 0; j < FEAT / 32; ++j) fv[j] = feat[b * FEAT + j * 32 + lane];
  float mine = -INFINITY;
  for (int c = 0; c < C; ++c) {
    float p = 0.f;
#pragma unroll
    for (int j = 0; j < FEAT / 32; ++j) p = fmaf(fv[j], W[c * FEAT + j * 32 + lane], p);
    p = warp_sum(p) + bias[c];
    if (lane == c) mine = p;
  }
  float m = mine;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
  const float ex = lane < C ? expf(mine - m) : 0.f;
  const float se = warp_sum(ex);
  if (lane < C) logp[b * C + lane] = (mine - m) - logf(se);
}

// grad_feat[b][f] = sum_c dl[b][c] W[c][f],  dl = g - exp(logp) * sum_c g  (log_softmax backward)
__global__ void __launch_bounds__(256) k_head_bwd_feat(const float* __restrict__ g,
                                                       const float* __restrict__ logp, int64_t B,
                                                       int C, const float* __restrict__ W,
                                                       float* __restrict__ gfeat,
                                                       float* __restrict__ dl_out) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const float gv = lane < C ? g[b * C + lane] : 0.f;
  const float sg = warp_sum(gv);
  const float dl = lane < C ? gv - expf(logp[b * C + lane]) * sg : 0.f;
  if (lane < C) dl_out[b * C + lane] = dl;
  if (gfeat == nullptr) return;
  float out[FEAT / 32];
#pragma unroll
  for (int j = 0; j < FEAT / 32; ++j) out[j] = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = __shfl_sync(FULL_MASK, dl, c);
#pragma unroll
    for (int j = 0; j < FEAT / 32; ++j) out[j] = fmaf(d, W[c * FEAT + j * 32 + lane], out[j]);
  }
#pragma unroll
  for (int j = 0; j < FEAT / 32; ++j) gfeat[b * FEAT + j * 32 + lane] = out[j];
}

// One tree's training head, by one warp (lane c holds class c): fv = the lane's eight feature columns
// (j * 32 + lane); returns the lane's eight columns of grad_feat in out.
//   logp = log_softmax(feat W^T + b);  g = -[c == y] / B_global;  dl = g - exp(logp) * sum(g)
//   grad_feat = dl W;  lossvec[b] = -logp[b][y_b]   (summed in tree order by k_loss_sum)
__device__ __forceinline__ void head_train_tree(const float (&fv)[4 * H / 32], int lane, int64_t b, const int64_t* __restrict__ y,
                                                int C, float inv_bg, const float* __restrict__ W, const float* __restrict__ bias,
                                                float* __restrict__ logp, float* __restrict__ dl_out, float* __restrict__ lossvec,
                                                float (&out)[4 * H / 32]) {
  constexpr int NJ = 4 * H / 32;
  float mine = -INFINITY;
  for (int c = 0; c < C; ++c) {
    float p = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) p = fmaf(fv[j], W[c * (4 * H) + j * 32 + lane], p);
    p = warp_sum(p) + bias[c];
    if (lane == c) mine = p;
  }
  float m = mine;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
  const float ex = lane < C ? expf(mine - m) : 0.f;
  const float se = warp_sum(ex);
  const float lp = (mine - m) - logf(se);
  if (lane < C) logp[b * C + lane] = lp;
  const int64_t t = y[b];
  const bool hit = t >= 0 && t < C;
  const float gv = (lane < C && hit && lane == (int)t) ? -inv_bg : 0.f;
  const float sg = warp_sum(gv);
  const float dl = lane < C ? gv - expf(lp) * sg : 0.f;
  if (lane < C) dl_out[b * C + lane] = dl;
  const float mylp = __shfl_sync(FULL_MASK, lp, hit ? (int)t : 0);
  if (lane == 0) lossvec[b] = hit ? -mylp : 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) out[j] = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = __shfl_sync(FULL_MASK, dl, c);
#pragma unroll
    for (int j = 0; j < NJ; ++j) out[j] = fmaf(d, W[c * (4 * H) + j * 32 + lane], out[j]);
  }
}

// Training head in one launch: fc + log_softmax (:129-130), F.nll_loss's gradient (:184) and the
// log_softmax / fc backward towards the features.  Warp per tree (head_train_tree).
__global__ void __launch_bounds__(256) k_head_train(const float* __restrict__ feat, const int64_t* __restrict__ y,
                                                    int64_t B, int C, float inv_bg, const float* __restrict__ W,
                                                    const float* __restrict__ bias, float* __restrict__ logp,
                                                    float* __restrict__ dl_out, float* __restrict__ gfeat,
                                                    float* __restrict__ lossvec) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float fv[FEAT / 32], out[FEAT / 32];
#pragma unroll
  for (int j = 0; j < FEAT / 32; ++j) fv[j] = feat[b * FEAT + j * 32 + lane];
  head_train_tree(fv, lane, b, y, C, inv_bg, W, bias, logp, dl_out, lossvec, out);
#pragma unroll
  for (int j = 0; j < FEAT / 32; ++j) gfeat[b * FEAT + j * 32 + lane] = out[j];
}
// The step's tail in ONE launch (was three: k_readout_final -> k_head_train -> k_gscale, each a few
// microseconds of work behind a launch gap on the critical path).  CTA = 4 trees:
//   256 threads: (tree, column) pairs finish the readout of both directions (readout_final_tree), feat -> shared;
//   4 warps    : one per tree, the head exactly as k_head_train (fc, log_softmax, nll gradient, grad_feat -> shared);
//   256 threads: gs = grad_feat / n_b per direction and the db2 partial of the CTA's four trees (k_gscale's order).
__global__ void __launch_bounds__(256) k_train_tail(TailArgs a) {
  __shared__ float s_feat[4][FEAT];
  __shared__ float s_g[4][FEAT];
  __shared__ float red[2][4][H];
  const int tl = threadIdx.x >> 6, f = threadIdx.x & 63;
  const int64_t b = (int64_t)blockIdx.x * 4 + tl;
  const bool valid = b < a.ro.B;
  float cnt[2] = {0.f, 0.f};
  if (valid) {
    float mean[2], root[2];
    readout_final_tree(a.ro, b, f, mean, root, cnt);
    for (int d = 0; d < a.ro.ndir; ++d) {
      s_feat[tl][a.ro.feat_base[d] + f] = mean[d];
      s_feat[tl][a.ro.feat_base[d] + H + f] = root[d];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w < 4) {                      // warp w: the head of tree 4 * blockIdx.x + w
    const int64_t bw = (int64_t)blockIdx.x * 4 + w;
    if (bw < a.ro.B) {
      float fv[FEAT / 32], out[FEAT / 32];
#pragma unroll
      for (int j = 0; j < FEAT / 32; ++j) fv[j] = s_feat[w][j * 32 + lane];
      head_train_tree(fv, lane, bw, a.y, a.C, a.inv_bg, a.W, a.bias, a.logp, a.dl, a.lossvec, out);
#pragma unroll
      for (int j = 0; j < FEAT / 32; ++j) {
        s_g[w][j * 32 + lane] = out[j];
        a.gfeat[bw * FEAT + j * 32 + lane] = out[j];
      }
    }
  }
  __syncthreads();
  for (int d = 0; d < a.ro.ndir; ++d) {
    float acc = 0.f;
    if (valid) {
      const int n = a.gs.node_ptr[b + 1] - a.gs.node_ptr[b];
      const float gv = __fdiv_rn(s_g[tl][a.gs.feat_base[d] + f], (float)(n > 0 ? n : 1));
      a.gs.gs[d][b * H + f] = gv;
      acc = fmaf(gv, cnt[d], acc);
    }
    red[d][tl][f] = acc;
  }
  __syncthreads();
  if (tl == 0)
    for (int d = 0; d < a.ro.ndir; ++d)
      a.gs.part[d][(int64_t)blockIdx.x * H + f] = ((red[d][0][f] + red[d][1][f]) + red[d][2][f]) + red[d][3][f];
}

// loss = (1/Bg) * sum_b lossvec[b], same strided order as k_nll
__global__ void __launch_bounds__(256) k_loss_sum(const float* __restrict__ lossvec, int64_t B, float inv_bg,
                                                  float* __restrict__ loss) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += 256) acc += lossvec[b];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] * inv_bg;
}

// Per-class confusion counts of a batch, accumulated on the device: replaces
// `_, pred = out.max(dim=-1)` + the Python loops of tools/evaluate.py:3-31 (evaluation4class) /
// :93-108 (evaluationclass), without the per-batch `.item()` syncs of BiGCN_Twitter.py:188-191.
//   counts[c] = {TP, FN, FP, TN} for class c;  totals = {trees, correct, sum of -logp[y]*1e6 (fixed point)}
// Integer atomics only: the sums do not depend on the order of arrival.
__global__ void __launch_bounds__(256) k_eval_counts(const float* __restrict__ logp, const int64_t* __restrict__ y,
                                                     int64_t B, int C, unsigned long long* counts,
                                                     unsigned long long* totals) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int pred = 0;
  float best = logp[b * C];
  for (int c = 1; c < C; ++c) {   // first maximum, as torch.max
    const float v = logp[b * C + c];
    if (v > best) {
      best = v;
      pred = c;
    }
  }
  const int64_t act = y[b];
  for (int c = 0; c < C; ++c) {
    const int slot = (act == c ? (pred == c ? 0 : 1) : (pred == c ? 2 : 3));
    atomicAdd(&counts[c * 4 + slot], 1ull);
  }
  atomicAdd(&totals[0], 1ull);
  if (act == pred) atomicAdd(&totals[1], 1ull);
  if (act >= 0 && act < C) {
    // fixed-point loss term; a non-finite or absurd log-prob (NaN parameters) saturates instead of llrintf's UB
    const float t = -logp[b * C + act] * 1e6f;
    const float tc = t == t ? fminf(fmaxf(t, 0.f), 1e15f) : 1e15f;
    atomicAdd(&totals[2], (unsigned long long)llrintf(tc));
  }
}

// dW[c][f] = sum_b dl[b][c] feat[b][f] (+ bias column f == 256): thread per (c, f), a chunk
// of HB_TREES trees per blockIdx.y, trees in order; chunks are summed in order by k_head_bwd_red.
constexpr int HB_TREES = 32;
__global__ void k_head_bwd_w(const float* __restrict__ dl, const float* __restrict__ feat, int64_t B,
                             int C, float* __restrict__ part) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * (FEAT + 1)) return;
  const int c = idx / (FEAT + 1), f = idx % (FEAT + 1);
  const int64_t b0 = (int64_t)blockIdx.y * HB_TREES, b1 = min(B, b0 + HB_TREES);
  float acc = 0.f;
  for (int64_t b = b0; b < b1; ++b) acc = fmaf(dl[b * C + c], f < FEAT ? feat[b * FEAT + f] : 1.f, acc);
  part[(size_t)blockIdx.y * C * (FEAT + 1) + idx] = acc;
}
__global__ void k_head_bwd_red(const float* __restrict__ part, int nchunk, int C,
                               float* __restrict__ dW, float* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * (FEAT + 1)) return;
  const int c = idx / (FEAT + 1), f = idx % (FEAT + 1);
  float acc = 0.f;
  for (int q = 0; q < nchunk; ++q) acc += part[(size_t)q * C * (FEAT + 1) + idx];
  if (f < FEAT) dW[c * FEAT + f] = acc;
  else db[c] = acc;
}

// loss = -(1/Bg) sum_b logp[b][y_b]; grad[b][c] = -(1/Bg) [c == y_b]   (single CTA, trees in order)
__global__ void __launch_bounds__(256) k_nll(const float* __restrict__ logp,
                                             const int64_t* __restrict__ y, int64_t B, int C,
                                             float inv_bg, float* loss, float* grad) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += 256) {
    const int64_t t = y[b];
    if (t >= 0 && t < C) acc -= logp[b * C + t];
    if (grad)
      for (int c = 0; c < C; ++c) grad[b * C + c] = (c == t) ? -inv_bg : 0.f;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss) *loss = red[0] * inv_bg;
}

// torch.optim.Adam, single-tensor form (coupled L2 weight decay, no amsgrad)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, int64_t n, const int64_t* __restrict__ seg_end,
                       const float* __restrict__ seg_lr, int n_seg, double beta1d, double beta2d,
                       float eps, float wd, float gscale, int64_t* step_count) {
  __shared__ float s_bc[2];
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  // the first quad of every thread is on its way while thread 0 works out the bias corrections
  float4 pv0 = make_float4(0.f, 0.f, 0.f, 0.f), gv0 = pv0, mv0 = pv0, vv0 = pv0;
  if (tid < n4) {
    pv0 = reinterpret_cast<float4*>(p)[tid];
    gv0 = reinterpret_cast<const float4*>(g)[tid];
    mv0 = reinterpret_cast<float4*>(m)[tid];
    vv0 = reinterpret_cast<float4*>(v)[tid];
  }
  if (threadIdx.x == 0) {   // the double-precision pow() runs once per block, not per thread
    const double step = (double)(*(volatile int64_t*)step_count + 1);
    s_bc[0] = (float)(1.0 - pow(beta1d, step));
    s_bc[1] = (float)sqrt(1.0 - pow(beta2d, step));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2s = s_bc[1];
  const float beta2 = (float)beta2d;
  const float omb1 = (float)(1.0 - beta1d), omb2 = (float)(1.0 - beta2d);
  auto lr_of = [&](int64_t i) {
    float lr = seg_lr[n_seg - 1];
    for (int s = 0; s < n_seg; ++s)
      if (i < seg_end[s]) {
        lr = seg_lr[s];
        break;
      }
    return lr;
  };
  auto upd = [&](float pv, float gv, float& mi, float& vi, float lr) {
    const float gr = fmaf(wd, pv, gv * gscale);
    mi = mi + (gr - mi) * omb1;
    vi = vi * beta2 + omb2 * (gr * gr);
    const float denom = sqrtf(vi) / bc2s + eps;
    return pv - (lr / bc1) * (mi / denom);
  };
  for (int64_t q = tid; q < n4; q += nth) {
    float4 pv = pv0, gv = gv0, mv = mv0, vv = vv0;
    if (q != tid) {
      pv = reinterpret_cast<float4*>(p)[q];
      gv = reinterpret_cast<const float4*>(g)[q];
      mv = reinterpret_cast<float4*>(m)[q];
      vv = reinterpret_cast<float4*>(v)[q];
    }
    // lr groups may change inside a quad only if a segment end is not a multiple of 4
    const float lr0 = lr_of(4 * q + 0), lr3 = lr_of(4 * q + 3);
    const bool same = lr0 == lr3;
    pv.x = upd(pv.x, gv.x, mv.x, vv.x, lr0);
    pv.y = upd(pv.y, gv.y, mv.y, vv.y, same ? lr0 : lr_of(4 * q + 1));
    pv.z = upd(pv.z, gv.z, mv.z, vv.z, same ? lr0 : lr_of(4 * q + 2));
    pv.w = upd(pv.w, gv.w, mv.w, vv.w, lr3);
    reinterpret_cast<float4*>(p)[q] = pv;
    reinterpret_cast<float4*>(m)[q] = mv;
    reinterpret_cast<float4*>(v)[q] = vv;
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nth) {
    float mi = m[i], vi = v[i];
    p[i] = upd(p[i], g[i], mi, vi, lr_of(i));
    m[i] = mi;
    v[i] = vi;
  }
  // the last block to finish advances the step (every block read it at its start) and resets
  // the arrival counter kept in step_count[1]
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* arrive = reinterpret_cast<unsigned long long*>(step_count + 1);
    __threadfence();
    if (atomicAdd(arrive, 1ull) == (unsigned long long)gridDim.x - 1) {
      *arrive = 0ull;
      *step_count += 1;
      step_count[2] += 1;   // calls counter: the dropout seed offset of the next step (opts.seed_dev)
    }
  }
}
__global__ void k_step_inc(int64_t* step_count) {
  *step_count += 1;
  step_count[2] += 1;
}

// ---- data-parallel optimiser step over peer memory (NVLink 5 / NVSwitch) ---------------------
// Replaces "all-reduce the flat gradient, then Adam on every rank" (the reference trains on one
// GPU; SURVEY.md 8e adds the tree-sharded data parallelism).  Every rank's flat gradient and
// parameter buffers are mapped into every process (symmetric memory).  Rank r owns the slice
// [lo, hi) of the flat vector: it loads that slice of ALL ranks' gradients through peer
// addresses, adds them in rank order (the same sum on any world size ordering, bit-identical
// parameters on all ranks), runs Adam on its shard of the optimiser state and stores the new
// parameter values straight into every rank's parameter buffer -- reduce-scatter, optimiser and
// all-gather in one kernel, 2/world of the all-reduce traffic per link direction, 1/world of the
// Adam work.  The caller brackets it with cross-rank barriers.
constexpr int DP_MAX_WORLD = 16;
struct DpArgs {
  const float* grads[DP_MAX_WORLD];
  float* params[DP_MAX_WORLD];
  int world, rank;
  float* m;
  float* v;
  int64_t lo, hi;    // multiples of 4 (except hi == n)
  const int64_t* seg_end;
  const float* seg_lr;
  int n_seg;
  double beta1, beta2;
  float eps, wd, gscale;
  int64_t* step_count;
  // in-kernel cross-rank barriers (or all NULL: the caller brackets the launch with its own barriers).  sig[q] is
  // rank q's signal block in symmetric memory, 2 * DP_MAX_WORLD uint64: [r] = "rank r's gradient is complete",
  // [DP_MAX_WORLD + r] = "rank r has read every gradient slice it needs and written its parameter slice everywhere";
  // values are epochs (step_count[2] + 1: never rewound), so nothing is ever reset.
  unsigned long long* sig[DP_MAX_WORLD];
  // PUSH form (with sig): stage[q] is rank q's staging buffer in symmetric memory, world x chunk floats; rank r writes
  // its gradient slice q into stage[q][r * chunk ...] with plain peer STORES (one way, bandwidth-bound) before the
  // first barrier, and every rank then reduces its own slice from LOCAL memory -- no peer loads (round trips) at all.
  float* stage[DP_MAX_WORLD];
  int64_t chunk, n;
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256) k_dp_reduce_adam(DpArgs a) {
  __shared__ float s_bc[2];
  const bool fused_sync = a.sig[0] != nullptr;
  const unsigned long long epoch = (unsigned long long)(*(volatile int64_t*)(a.step_count + 2)) + 1ull;
  // timestamps of the last launch (globaltimer, ns) in the local signal block [32..35]: kernel start, gradients of all
  // ranks complete, own slice done, every rank done -- where a data-parallel step waits (tools/dp_phases.py)
  unsigned long long* dbg = fused_sync ? a.sig[a.rank] + 2 * DP_MAX_WORLD : nullptr;
  auto stamp = [&](int k) {
    if (dbg != nullptr) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[k] = t;
    }
  };
  if (blockIdx.x == 0 && threadIdx.x == 0) stamp(0);
  if (threadIdx.x == 0) {
    const double step = (double)(*(volatile int64_t*)a.step_count + 1);
    s_bc[0] = (float)(1.0 - pow(a.beta1, step));
    s_bc[1] = (float)sqrt(1.0 - pow(a.beta2, step));
  }
  const bool push = fused_sync && a.stage[0] != nullptr;
  __shared__ int s_signal;
  if (fused_sync) {
    if (push) {
      // phase 0: my gradient slice q -> rank q's staging row [rank] (peer stores), all CTAs, grid-stride
      const int64_t tid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth0 = (int64_t)gridDim.x * blockDim.x;
      for (int q = 0; q < a.world; ++q) {
        if (q == a.rank) continue;
        const int64_t qlo = min(a.chunk * q, a.n), qhi = min(qlo + a.chunk, a.n);
        const float4* src = reinterpret_cast<const float4*>(a.grads[a.rank] + qlo);
        float4* dst = reinterpret_cast<float4*>(a.stage[q] + (int64_t)a.rank * a.chunk);
        const int64_t n4q = (qhi - qlo) / 4;
        for (int64_t t = tid0; t < n4q; t += nth0) dst[t] = src[t];
        for (int64_t t = 4 * n4q + tid0; t < qhi - qlo; t += nth0)
          a.stage[q][(int64_t)a.rank * a.chunk + t] = a.grads[a.rank][qlo + t];
      }
      __syncthreads();
      if (threadIdx.x == 0) {   // the last CTA to finish pushing tells every rank (step_count[3]: its arrival counter)
        __threadfence_system();
        unsigned long long* arrive0 = reinterpret_cast<unsigned long long*>(a.step_count + 3);
        s_signal = atomicAdd(arrive0, 1ull) == (unsigned long long)gridDim.x - 1;
        if (s_signal) *arrive0 = 0ull;
      }
      __syncthreads();
    } else if (threadIdx.x == 0) {
      s_signal = blockIdx.x == 0;   // pull form: the gradient was complete before this kernel started
    }
    if (!push) __syncthreads();
    // barrier 1: this rank's gradient is complete (and, push form, delivered): tell every rank, then wait until
    // every rank has told us.  One CTA signals, every CTA waits on LOCAL flags.
    if (s_signal && threadIdx.x < a.world) {
      __threadfence_system();
      st_release_sys(a.sig[threadIdx.x] + a.rank, epoch);
    }
    if (threadIdx.x < a.world) {
      const unsigned long long* f = a.sig[a.rank] + threadIdx.x;
      while (ld_acquire_sys(f) < epoch) __nanosleep(32);
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) stamp(1);
  const float bc1 = s_bc[0], bc2s = s_bc[1];
  const float beta2 = (float)a.beta2;
  const float omb1 = (float)(1.0 - a.beta1), omb2 = (float)(1.0 - a.beta2);
  auto lr_of = [&](int64_t i) {
    float lr = a.seg_lr[a.n_seg - 1];
    for (int s = 0; s < a.n_seg; ++s)
      if (i < a.seg_end[s]) {
        lr = a.seg_lr[s];
        break;
      }
    return lr;
  };
  auto upd = [&](float pv, float gv, float& mi, float& vi, float lr) {
    const float gr = fmaf(a.wd, pv, gv * a.gscale);
    mi = mi + (gr - mi) * omb1;
    vi = vi * beta2 + omb2 * (gr * gr);
    const float denom = sqrtf(vi) / bc2s + a.eps;
    return pv - (lr / bc1) * (mi / denom);
  };
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = (a.hi - a.lo) / 4;
  for (int64_t q = tid; q < n4; q += nth) {
    const int64_t i = a.lo + 4 * q;
    float4 gq[DP_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r)
      if (r < a.world)   // push form: every slice sits in local memory (rank r's in my staging row r); pull form: peer loads
        gq[r] = (push && r != a.rank) ? __ldcg(reinterpret_cast<const float4*>(a.stage[a.rank] + (int64_t)r * a.chunk + (i - a.lo)))
                                      : *reinterpret_cast<const float4*>(a.grads[r] + i);
    float4 g = gq[0];
#pragma unroll
    for (int r = 1; r < DP_MAX_WORLD; ++r)
      if (r < a.world) {
        g.x += gq[r].x; g.y += gq[r].y; g.z += gq[r].z; g.w += gq[r].w;
      }
    float4 pv = *reinterpret_cast<const float4*>(a.params[a.rank] + i);
    float4 mv = *reinterpret_cast<float4*>(a.m + i);
    float4 vv = *reinterpret_cast<float4*>(a.v + i);
    const float lr0 = lr_of(i + 0), lr3 = lr_of(i + 3);
    const bool same = lr0 == lr3;
    pv.x = upd(pv.x, g.x, mv.x, vv.x, lr0);
    pv.y = upd(pv.y, g.y, mv.y, vv.y, same ? lr0 : lr_of(i + 1));
    pv.z = upd(pv.z, g.z, mv.z, vv.z, same ? lr0 : lr_of(i + 2));
    pv.w = upd(pv.w, g.w, mv.w, vv.w, lr3);
    *reinterpret_cast<float4*>(a.m + i) = mv;
    *reinterpret_cast<float4*>(a.v + i) = vv;
#pragma unroll
    for (int r = 0; r < DP_MAX_WORLD; ++r)
      if (r < a.world) *reinterpret_cast<float4*>(a.params[r] + i) = pv;          // peer stores
  }
  for (int64_t i = a.lo + 4 * n4 + tid; i < a.hi; i += nth) {   // tail of the last rank's slice
    float g = 0.f;
    for (int r = 0; r < a.world; ++r)
      g += (push && r != a.rank) ? __ldcg(a.stage[a.rank] + (int64_t)r * a.chunk + (i - a.lo)) : a.grads[r][i];
    float mi = a.m[i], vi = a.v[i];
    const float pn = upd(a.params[a.rank][i], g, mi, vi, lr_of(i));
    a.m[i] = mi;
    a.v[i] = vi;
    for (int r = 0; r < a.world; ++r) a.params[r][i] = pn;
  }
  __syncthreads();
  __shared__ int s_last;
  if (threadIdx.x == 0) {   // last block: advance the step, reset the arrival counter (step_count[1])
    unsigned long long* arrive = reinterpret_cast<unsigned long long*>(a.step_count + 1);
    __threadfence_system();
    s_last = atomicAdd(arrive, 1ull) == (unsigned long long)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) stamp(2);
  if (fused_sync) {
    // barrier 2: every block of this rank is done (peer loads of the gradients, peer stores of the parameters,
    // fenced above): tell every rank, and leave only when every rank has said the same -- at kernel exit all
    // parameters are in place everywhere and every gradient buffer is free again.
    if (threadIdx.x < a.world) {
      st_release_sys(a.sig[threadIdx.x] + DP_MAX_WORLD + a.rank, epoch);
      const unsigned long long* f = a.sig[a.rank] + DP_MAX_WORLD + threadIdx.x;
      while (ld_acquire_sys(f) < epoch) __nanosleep(32);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    stamp(3);
    unsigned long long* arrive = reinterpret_cast<unsigned long long*>(a.step_count + 1);
    *arrive = 0ull;
    *a.step_count += 1;
    a.step_count[2] += 1;
  }
}

}  // namespace bigcn

using namespace bigcn;

extern "C" int bigcn_head_forward(const float* feat, int64_t B, int64_t C, const float* fc_w,
                                  const float* fc_b, float* logp, bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(C >= 1 && C <= 32, "head_forward: C must be in [1,32]");
  if (B == 0) return 0;
  k_head_fwd<<<(int)ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(feat, B, (int)C, fc_w, fc_b, logp);
  BIGCN_CHECK_LAUNCH("k_head_fwd");
  return 0;
}

extern "C" size_t bigcn_head_backward_scratch_floats(int64_t B, int64_t C) {
  const int64_t nchunk = B > 0 ? ceil_div(B, HB_TREES) : 1;
  return (size_t)(B * C + nchunk * C * (FEAT + 1));
}

extern "C" int bigcn_head_backward(const float* grad_logp, const float* logp, const float* feat,
                                   int64_t B, int64_t C, const float* fc_w, float* grad_feat,
                                   float* d_fc_w, float* d_fc_b, float* scratch,
                                   size_t scratch_floats, bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(C >= 1 && C <= 32, "head_backward: C must be in [1,32]");
  BIGCN_CHECK_ARG(scratch && scratch_floats >= bigcn_head_backward_scratch_floats(B, C),
                  "head_backward: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* dl = scratch;
  float* part = scratch + B * C;
  const int nchunk = B > 0 ? (int)ceil_div(B, HB_TREES) : 1;
  if (B > 0) {
    k_head_bwd_feat<<<(int)ceil_div(B, 8), 256, 0, st>>>(grad_logp, logp, B, (int)C, fc_w, grad_feat, dl);
    BIGCN_CHECK_LAUNCH("k_head_bwd_feat");
  }
  if (d_fc_w) {
    const int tot = (int)C * (FEAT + 1);
    k_head_bwd_w<<<dim3((tot + 127) / 128, nchunk), 128, 0, st>>>(dl, feat, B, (int)C, part);
    BIGCN_CHECK_LAUNCH("k_head_bwd_w");
    k_head_bwd_red<<<(tot + 127) / 128, 128, 0, st>>>(part, nchunk, (int)C, d_fc_w, d_fc_b);
    BIGCN_CHECK_LAUNCH("k_head_bwd_red");
  }
  return 0;
}

static_assert(GS_TREES == 4, "k_train_tail writes one db2 partial per four trees, as k_gscale does");
int bigcn::train_tail_launch(const TailArgs& a, cudaStream_t st) {
  if (a.ro.B == 0) return 0;
  k_train_tail<<<(int)ceil_div(a.ro.B, 4), 256, 0, st>>>(a);
  BIGCN_CHECK_LAUNCH("k_train_tail");
  return 0;
}

extern "C" size_t bigcn_head_train_scratch_floats(int64_t B, int64_t C) {
  return bigcn_head_backward_scratch_floats(B, C) + (size_t)(B > 0 ? B : 1);
}

// fc + log_softmax + nll_loss and their backward in one call (the training step's head):
// one launch on the caller's stream for what the features' backward waits for (grad_feat); the
// fc gradients and the loss scalar are formed on the side stream and joined by the next
// bigcn_features_backward / bigcn_adam_step / bigcn_dp_reduce_adam on this stream.
extern "C" int bigcn_head_train(const float* feat, const int64_t* y, int64_t B, int64_t C, int64_t B_global,
                                const float* fc_w, const float* fc_b, float* logp, float* loss, float* grad_feat,
                                float* d_fc_w, float* d_fc_b, float* scratch, size_t scratch_floats,
                                bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(C >= 1 && C <= 32, "head_train: C must be in [1,32]");
  BIGCN_CHECK_ARG(B_global > 0, "head_train: B_global must be positive");
  BIGCN_CHECK_ARG(feat && y && fc_w && fc_b && logp && loss && grad_feat && d_fc_w && d_fc_b, "head_train: NULL argument");
  BIGCN_CHECK_ARG(scratch && scratch_floats >= bigcn_head_train_scratch_floats(B, C), "head_train: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* dl = scratch;
  float* part = scratch + B * C;
  const int nchunk = B > 0 ? (int)ceil_div(B, HB_TREES) : 1;
  float* lossvec = part + (size_t)nchunk * C * (FEAT + 1);
  const float inv_bg = 1.0f / (float)B_global;
  if (B > 0) {
    k_head_train<<<(int)ceil_div(B, 8), 256, 0, st>>>(feat, y, B, (int)C, inv_bg, fc_w, fc_b, logp, dl, grad_feat,
                                                      lossvec);
    BIGCN_CHECK_LAUNCH("k_head_train");
  }
  cudaStream_t ss = side_fork(st);
  const int tot = (int)C * (FEAT + 1);
  k_head_bwd_w<<<dim3((tot + 127) / 128, nchunk), 128, 0, ss>>>(dl, feat, B, (int)C, part);
  BIGCN_CHECK_LAUNCH("k_head_bwd_w");
  k_head_bwd_red<<<(tot + 127) / 128, 128, 0, ss>>>(part, nchunk, (int)C, d_fc_w, d_fc_b);
  BIGCN_CHECK_LAUNCH("k_head_bwd_red");
  k_loss_sum<<<1, 256, 0, ss>>>(lossvec, B, inv_bg, loss);
  BIGCN_CHECK_LAUNCH("k_loss_sum");
  return 0;
}

// the fused tail: `ta` arrives with the readout / gscale halves filled in by api.cu (they live in the features workspace)
int bigcn::train_tail_run(TailArgs ta, const float* feat, const int64_t* y, int64_t B, int64_t C, int64_t B_global,
                          const float* fc_w, const float* fc_b, float* logp, float* loss, float* grad_feat,
                          float* d_fc_w, float* d_fc_b, float* scratch, size_t scratch_floats, cudaStream_t st) {
  BIGCN_CHECK_ARG(C >= 1 && C <= 32, "train_tail: C must be in [1,32]");
  BIGCN_CHECK_ARG(B_global > 0, "train_tail: B_global must be positive");
  BIGCN_CHECK_ARG(feat && y && fc_w && fc_b && logp && loss && grad_feat && d_fc_w && d_fc_b, "train_tail: NULL argument");
  BIGCN_CHECK_ARG(scratch && scratch_floats >= bigcn_head_train_scratch_floats(B, C), "train_tail: scratch too small");
  float* dl = scratch;
  float* part = scratch + B * C;
  const int nchunk = B > 0 ? (int)ceil_div(B, HB_TREES) : 1;
  float* lossvec = part + (size_t)nchunk * C * (FEAT + 1);
  const float inv_bg = 1.0f / (float)B_global;
  ta.y = y; ta.C = (int)C; ta.inv_bg = inv_bg; ta.W = fc_w; ta.bias = fc_b; ta.logp = logp; ta.dl = dl;
  ta.gfeat = grad_feat; ta.lossvec = lossvec;
  if (int rc = train_tail_launch(ta, st)) return rc;
  cudaStream_t ss = side_fork(st);
  const int tot = (int)C * (FEAT + 1);
  k_head_bwd_w<<<dim3((tot + 127) / 128, nchunk), 128, 0, ss>>>(dl, feat, B, (int)C, part);
  BIGCN_CHECK_LAUNCH("k_head_bwd_w");
  k_head_bwd_red<<<(tot + 127) / 128, 128, 0, ss>>>(part, nchunk, (int)C, d_fc_w, d_fc_b);
  BIGCN_CHECK_LAUNCH("k_head_bwd_red");
  k_loss_sum<<<1, 256, 0, ss>>>(lossvec, B, inv_bg, loss);
  BIGCN_CHECK_LAUNCH("k_loss_sum");
  return 0;
}

extern "C" int bigcn_eval_counts(const float* logp, const int64_t* y, int64_t B, int64_t C, int64_t* counts,
                                 int64_t* totals, bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(C >= 1 && C <= 32 && counts && totals, "eval_counts: bad arguments");
  if (B == 0) return 0;
  k_eval_counts<<<(int)ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(
      logp, y, B, (int)C, reinterpret_cast<unsigned long long*>(counts), reinterpret_cast<unsigned long long*>(totals));
  BIGCN_CHECK_LAUNCH("k_eval_counts");
  return 0;
}

extern "C" int bigcn_nll_loss(const float* logp, const int64_t* y, int64_t B, int64_t C,
                              int64_t B_global, float* loss, float* grad_logp,
                              bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(B_global > 0, "nll_loss: B_global must be positive");
  k_nll<<<1, 256, 0, (cudaStream_t)stream>>>(logp, y, B, (int)C, 1.0f / (float)B_global, loss, grad_logp);
  BIGCN_CHECK_LAUNCH("k_nll");
  return 0;
}

extern "C" int bigcn_dp_slice(int64_t n, int32_t world, int32_t rank, int64_t* lo, int64_t* hi) {
  BIGCN_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && lo && hi, "dp_slice: bad arguments");
  const int64_t chunk = ceil_div(ceil_div(n, 4), world) * 4;
  *lo = chunk * rank < n ? chunk * rank : n;
  *hi = *lo + chunk < n ? *lo + chunk : n;
  return 0;
}

extern "C" int64_t bigcn_dp_stage_chunk(int64_t n, int32_t world) {
  return world >= 1 ? ceil_div(ceil_div(n, 4), world) * 4 : 0;
}

extern "C" int bigcn_dp_reduce_adam(const float* const* grads, float* const* params, int32_t world, int32_t rank,
                                    float* exp_avg, float* exp_avg_sq, int64_t n, const int64_t* seg_end,
                                    const float* seg_lr, int32_t n_seg, double beta1, double beta2, double eps,
                                    double weight_decay, double grad_scale, int64_t* step_count,
                                    void* const* signals, float* const* stage, bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(grads && params && world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world,
                  "dp_reduce_adam: world must be 1..%d", DP_MAX_WORLD);
  BIGCN_CHECK_ARG(n_seg >= 1, "dp_reduce_adam: need at least one lr segment");
  cudaStream_t st = (cudaStream_t)stream;
  side_join(st);   // fc gradients / loss of bigcn_head_train
  DpArgs a{};
  for (int r = 0; r < world; ++r) {
    BIGCN_CHECK_ARG(grads[r] && params[r], "dp_reduce_adam: NULL peer buffer");
    BIGCN_CHECK_ARG(((reinterpret_cast<uintptr_t>(grads[r]) | reinterpret_cast<uintptr_t>(params[r])) & 15) == 0,
                    "dp_reduce_adam: peer buffers must be 16 B aligned");
    a.grads[r] = grads[r];
    a.params[r] = params[r];
    a.sig[r] = signals ? reinterpret_cast<unsigned long long*>(signals[r]) : nullptr;
    BIGCN_CHECK_ARG(!signals || signals[r], "dp_reduce_adam: NULL signal block");
    a.stage[r] = (signals && stage) ? stage[r] : nullptr;
    BIGCN_CHECK_ARG(!(signals && stage) || (stage[r] && (reinterpret_cast<uintptr_t>(stage[r]) & 15) == 0),
                    "dp_reduce_adam: NULL / unaligned staging buffer");
  }
  a.n = n;
  a.chunk = ceil_div(ceil_div(n, 4), world) * 4;   // = the slice length of bigcn_dp_slice
  a.world = world; a.rank = rank; a.m = exp_avg; a.v = exp_avg_sq;
  bigcn_dp_slice(n, world, rank, &a.lo, &a.hi);
  a.seg_end = seg_end; a.seg_lr = seg_lr; a.n_seg = n_seg;
  a.beta1 = beta1; a.beta2 = beta2; a.eps = (float)eps; a.wd = (float)weight_decay; a.gscale = (float)grad_scale;
  a.step_count = step_count;
  if (a.hi > a.lo || signals) {   // with in-kernel barriers even an empty slice takes part
    int blocks = (int)ceil_div((a.hi - a.lo + 3) / 4, 256);
    // Every CTA must be able to be resident at once: in the push form the rank signals only when ALL of its CTAs have
    // delivered their slices, while the CTAs that are done spin on the peers' flags and keep their slots.  (Kernels of
    // other streams may hold slots for a while; they finish on their own and the remaining CTAs then start.)
    static int per_sm = 0;
    if (per_sm == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dp_reduce_adam, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    const int cap = num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_dp_reduce_adam<<<blocks, 256, 0, st>>>(a);
    BIGCN_CHECK_LAUNCH("k_dp_reduce_adam");
  } else {
    k_step_inc<<<1, 1, 0, st>>>(step_count);
    BIGCN_CHECK_LAUNCH("k_step_inc");
  }
  return 0;
}

extern "C" int bigcn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               int64_t n, const int64_t* seg_end, const float* seg_lr, int32_t n_seg,
                               double beta1, double beta2, double eps, double weight_decay,
                               double grad_scale, int64_t* step_count, bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(n_seg >= 1, "adam_step: need at least one lr segment");
  cudaStream_t st = (cudaStream_t)stream;
  side_join(st);   // fc gradients / loss of bigcn_head_train
  if (n > 0) {
    int blocks = (int)ceil_div(n, 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    k_adam<<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, seg_end, seg_lr, n_seg, beta1,
                                   beta2, (float)eps, (float)weight_decay, (float)grad_scale, step_count);
    BIGCN_CHECK_LAUNCH("k_adam");
  } else {
    k_step_inc<<<1, 1, 0, st>>>(step_count);
    BIGCN_CHECK_LAUNCH("k_step_inc");
  }
  return 0;
}
