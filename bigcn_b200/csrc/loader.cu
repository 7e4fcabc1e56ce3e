// On-device batch assembly + DropEdge: replaces, for a dataset kept
// resident in HBM, what BiGraphDataset.__getitem__ (Process/dataset.py:64-99) and PyG's
// Batch.from_data_list collate (model/Twitter/BiGCN_Twitter.py:168) do on the host per batch:
//   * DropEdge: keep int(e * (1 - rate)) positions of a tree's edge list, uniformly without
//     replacement, order preserved, drawn independently for the TD list and for the BU list
//     (dataset.py:68-90; random.sample -> the k smallest of per-position Philox keys);
//   * collate: concatenate x, offset edge_index / BU_edge_index / rootindex by the cumulative
//     node count, batch[i] = tree of node i, y.
// The packed dataset keeps x as CSR (the `index:count` pairs getTwittergraph.py:16-24 reads,
// never densified), so the assembled batch feeds the BIGCN_GEMM_SPARSE path directly.
// Offsets per tree are computed by the caller on the host from the (host-known) tree sizes:
// nothing here synchronises.
#include "kernels.cuh"

namespace bigcn {

struct AsmArgs {
  // packed dataset
  const int64_t* node_ptr;   // [T+1]
  const int64_t* edge_ptr;   // [T+1]
  const int32_t* edge_src;   // [E_all] local ids: row 0 of the tree's edgeindex (parent)
  const int32_t* edge_dst;   // [E_all] row 1 (child)
  const int64_t* x_ptr;      // [N_all+1]
  const int32_t* x_col;
  const float* x_val;
  const int32_t* root_local; // [T]
  const int64_t* y_all;      // [T]
  // this batch
  const int64_t* tree_id;    // [B]
  const int64_t* node_off;   // [B+1]
  const int64_t* td_off;     // [B+1] kept TD edges
  const int64_t* bu_off;     // [B+1]
  const int64_t* nnz_off;    // [B+1]
  int64_t B, E_td, E_bu;
  uint32_t k0, k1;           // Philox key (seed)
  // outputs
  int64_t* edge_index;       // [2][E_td]
  int64_t* bu_edge_index;    // [2][E_bu]
  int64_t* batch;            // [N]
  int64_t* rootindex;        // [B]
  int64_t* y;                // [B]
  int32_t* ox_ptr;           // [N+1]
  int32_t* ox_col;
  float* ox_val;
};

// nodes, features, labels: CTA per tree
__global__ void __launch_bounds__(256) k_asm_nodes(AsmArgs a) {
  const int64_t b = blockIdx.x;
  const int64_t t = a.tree_id[b];
  const int64_t n0 = a.node_ptr[t], n = a.node_ptr[t + 1] - n0;
  const int64_t off = a.node_off[b];
  const int64_t z0 = a.x_ptr[n0], nz = a.x_ptr[n0 + n] - z0;
  const int64_t zoff = a.nnz_off[b];
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    a.batch[off + i] = b;
    a.ox_ptr[off + i] = (int32_t)(zoff + (a.x_ptr[n0 + i] - z0));
  }
  for (int64_t p = threadIdx.x; p < nz; p += blockDim.x) {
    a.ox_col[zoff + p] = a.x_col[z0 + p];
    a.ox_val[zoff + p] = a.x_val[z0 + p];
  }
  if (threadIdx.x == 0) {
    a.rootindex[b] = off + a.root_local[t];
    a.y[b] = a.y_all[t];
    if (b == a.B - 1) a.ox_ptr[off + n] = (int32_t)(zoff + nz);
  }
}

// edges.  DropEdge = a uniformly random k-subset of the e positions, order preserved
// (random.sample(range(e), k) followed by sorted(), dataset.py:72-75,84-87).
// CTA per (tree, direction): every position draws a 32-bit Philox key; a 32-step bisection finds
// the k-th smallest key v; positions with key < v are kept, plus the first ones with key == v
// until k are kept (ties, ~e^2 / 2^33 likely, resolved by position); an ordered block scan
// compacts the kept edges.  Every k-subset is equally likely (up to those ties).
constexpr int DE_MAX = 8192;   // edges per tree handled in shared memory (32 KB of keys)

__device__ __forceinline__ uint32_t edge_key(const AsmArgs& a, int64_t t, int dir, int64_t i) {
  const Philox4 r = philox4x32_10((uint32_t)(i >> 2), (uint32_t)(t & 0xffffffffll), (uint32_t)((uint64_t)t >> 32),
                                  0x44450000u + (uint32_t)dir, a.k0, a.k1);
  return philox_elem(r, (int)(i & 3));
}

__device__ __forceinline__ int block_sum128(int v, int* red) {   // 128 threads, all get the total
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(128) k_asm_edges(AsmArgs a) {
  __shared__ uint32_t key[DE_MAX];
  __shared__ int red[4];
  __shared__ int wbase[4];
  const int64_t b = blockIdx.x >> 1;
  const int dir = (int)(blockIdx.x & 1);          // 0: TD list [row; col], 1: BU list [col; row]
  const int64_t t = a.tree_id[b];
  const int64_t e0 = a.edge_ptr[t];
  const int e = (int)(a.edge_ptr[t + 1] - e0);
  const int64_t* offs = dir ? a.bu_off : a.td_off;
  const int64_t o0 = offs[b];
  const int k = (int)(offs[b + 1] - o0);
  const int64_t noff = a.node_off[b];
  int64_t* out = dir ? a.bu_edge_index : a.edge_index;
  const int64_t E = dir ? a.E_bu : a.E_td;
  auto emit = [&](int i, int64_t pos) {
    const int64_t s = a.edge_src[e0 + i] + noff, d = a.edge_dst[e0 + i] + noff;
    out[o0 + pos] = dir ? d : s;
    out[E + o0 + pos] = dir ? s : d;
  };
  if (k >= e) {   // rate 0 (or nothing to drop): plain offset copy
    for (int i = threadIdx.x; i < e; i += 128) emit(i, i);
    return;
  }
  if (e > DE_MAX) {   // very large tree: Knuth's selection sampling (TAOCP 3.4.2 S) by one thread
    if (threadIdx.x == 0) {
      int kept = 0;
      for (int i = 0; i < e && kept < k; ++i) {
        const uint32_t u = edge_key(a, t, dir, i);
        // u / 2^32 < (k - kept) / (e - i)
        if ((unsigned long long)u * (unsigned long long)(e - i) < ((unsigned long long)(k - kept) << 32)) emit(i, kept++);
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < e; i += 128) key[i] = edge_key(a, t, dir, i);
  __syncthreads();
  // smallest v with #{key <= v} >= k
  uint32_t lo = 0u, hi = 0xffffffffu;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    int c = 0;
    for (int i = threadIdx.x; i < e; i += 128) c += key[i] <= mid;
    if (block_sum128(c, red) >= k) hi = mid;
    else lo = mid + 1;
  }
  const uint32_t v = lo;
  int below = 0;
  for (int i = threadIdx.x; i < e; i += 128) below += key[i] < v;
  const int ties_kept = k - block_sum128(below, red);     // how many of the key == v positions are kept
  // ordered compaction, 128 positions per round
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int base = 0, ties_seen = 0;
  for (int i0 = 0; i0 < e; i0 += 128) {
    const int i = i0 + threadIdx.x;
    const bool in = i < e;
    const bool lt = in && key[i] < v, eq = in && key[i] == v;
    // position among the ties (block-ordered)
    const unsigned mq = __ballot_sync(FULL_MASK, eq);
    __syncthreads();
    if (lane == 0) wbase[w] = __popc(mq);
    __syncthreads();
    int tie_rank = ties_seen + __popc(mq & ((1u << lane) - 1u));
    for (int q = 0; q < w; ++q) tie_rank += wbase[q];
    const int ties_round = wbase[0] + wbase[1] + wbase[2] + wbase[3];
    const bool keep = lt || (eq && tie_rank < ties_kept);
    const unsigned mk = __ballot_sync(FULL_MASK, keep);
    __syncthreads();
    if (lane == 0) wbase[w] = __popc(mk);
    __syncthreads();
    int pos = base + __popc(mk & ((1u << lane) - 1u));
    for (int q = 0; q < w; ++q) pos += wbase[q];
    if (keep) emit(i, pos);
    base += wbase[0] + wbase[1] + wbase[2] + wbase[3];
    ties_seen += ties_round;
  }
}

}  // namespace bigcn

using namespace bigcn;

extern "C" int bigcn_assemble_batch(const int64_t* node_ptr, const int64_t* edge_ptr, const int32_t* edge_src,
                                    const int32_t* edge_dst, const int64_t* x_ptr, const int32_t* x_col,
                                    const float* x_val, const int32_t* root_local, const int64_t* y_all,
                                    const int64_t* tree_id, const int64_t* node_off, const int64_t* td_off,
                                    const int64_t* bu_off, const int64_t* nnz_off, int64_t B, int64_t E_td,
                                    int64_t E_bu, uint64_t seed, int64_t* edge_index, int64_t* bu_edge_index,
                                    int64_t* batch, int64_t* rootindex, int64_t* y, int32_t* ox_ptr, int32_t* ox_col,
                                    float* ox_val, bigcn_stream_t stream) {
  BIGCN_CHECK_ARG(B >= 0 && E_td >= 0 && E_bu >= 0, "assemble_batch: bad sizes");
  if (B == 0) return 0;
  BIGCN_CHECK_ARG(node_ptr && edge_ptr && x_ptr && root_local && y_all && tree_id && node_off && td_off && bu_off &&
                      nnz_off && batch && rootindex && y && ox_ptr,
                  "assemble_batch: NULL argument");
  AsmArgs a{};
  a.node_ptr = node_ptr; a.edge_ptr = edge_ptr; a.edge_src = edge_src; a.edge_dst = edge_dst;
  a.x_ptr = x_ptr; a.x_col = x_col; a.x_val = x_val; a.root_local = root_local; a.y_all = y_all;
  a.tree_id = tree_id; a.node_off = node_off; a.td_off = td_off; a.bu_off = bu_off; a.nnz_off = nnz_off;
  a.B = B; a.E_td = E_td; a.E_bu = E_bu;
  a.k0 = (uint32_t)(seed & 0xffffffffull); a.k1 = (uint32_t)(seed >> 32);
  a.edge_index = edge_index; a.bu_edge_index = bu_edge_index; a.batch = batch; a.rootindex = rootindex; a.y = y;
  a.ox_ptr = ox_ptr; a.ox_col = ox_col; a.ox_val = ox_val;
  cudaStream_t st = (cudaStream_t)stream;
  k_asm_nodes<<<(unsigned)B, 256, 0, st>>>(a);
  BIGCN_CHECK_LAUNCH("k_asm_nodes");
  k_asm_edges<<<(unsigned)(2 * B), 128, 0, st>>>(a);
  BIGCN_CHECK_LAUNCH("k_asm_edges");
  return 0;
}
