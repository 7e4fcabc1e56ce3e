// Integer graph prep: replaces gcn_norm / add_remaining_self_loops [torch_geometric]
// reached from BiGCN_Twitter.py:42,56,92,105 and the Python max(data.batch) of :47.
//
// For each direction it builds, from the int64 COO list exactly as PyG hands it over:
//   CSR by target (in_ptr/in_idx)   -> A-hat      (forward aggregate at col)
//   CSR by source (out_ptr/out_idx) -> A-hat^T    (backward)
//   deg (+1 self-loop), dis = 1/sqrt(deg) (IEEE div+sqrt == torch CPU pow(-0.5)), rowsum
// and node_ptr[B+1] from the sorted batch vector.  Self-loops in the input are dropped
// (add_remaining_self_loops re-adds one unit loop per node).  Entries of a CSR row keep
// edge-list order (stable LSD radix sort, no float atomics anywhere), so every output is
// bit-reproducible and equals oracle.gcn_oracle.graph_prep.
//
// All of this is HBM/latency-bound integer work: 32-bit keys, 8-bit digits,
// ceil(bits(N)/8) passes, the four sorts (2 directions x {by target, by source})
// batched in blockIdx.y so a prep is a fixed, short launch sequence.
#include "gather.cuh"

namespace bigcn {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;  // 2048 keys per block
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048

struct PrepDir {
  const int64_t* ei;
  int64_t E;
  bigcn_graph_t g;
  int32_t* in_eid;   // optional (edge-weighted conv): edge id of every CSR entry; the sort then carries the
  int32_t* out_eid;  // edge id as its payload and the last pass looks the endpoint up
};

struct PrepArgs {
  PrepDir d[2];
  int32_t ndir;
  int32_t deg_by;
  int64_t N, B;
  const int64_t* batch;
  int32_t* node_ptr;
  int32_t* flags;
  int32_t* cnt;      // [2*ndir][N]   sort s = 2*dir + orient (0: by target, 1: by source)
  int32_t* keys[2];  // [2*ndir][Emax] ping-pong
  int32_t* vals[2];
  int32_t* hist;     // [2*ndir][256][nblk]
  int32_t* bsum;     // [2*ndir][nscanblk]
  int64_t Emax;
  int32_t nblk;      // radix blocks (over Emax)
  int32_t nscanblk;
};

// ---- pass 0: validate, count degrees, emit (key,val) pairs, node_ptr ---------------
__global__ void k_prep_count(PrepArgs a) {
  const int dir = blockIdx.y;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const PrepDir& pd = a.d[dir];
  const int32_t sentinel = (int32_t)a.N;
  int32_t* cnt_in = a.cnt + (int64_t)(2 * dir) * a.N;
  int32_t* cnt_out = a.cnt + (int64_t)(2 * dir + 1) * a.N;
  int32_t* k_in = a.keys[0] + (int64_t)(2 * dir) * a.Emax;
  int32_t* k_out = a.keys[0] + (int64_t)(2 * dir + 1) * a.Emax;
  int32_t* v_in = a.vals[0] + (int64_t)(2 * dir) * a.Emax;
  int32_t* v_out = a.vals[0] + (int64_t)(2 * dir + 1) * a.Emax;
  for (int64_t e = t0; e < pd.E; e += stride) {
    const int64_t r = pd.ei[e], c = pd.ei[pd.E + e];
    const bool in_range = r >= 0 && r < a.N && c >= 0 && c < a.N;
    if (!in_range) atomicOr(a.flags, BIGCN_FLAG_EDGE_RANGE);
    const bool valid = in_range && r != c;
    if (valid) {
      atomicAdd(cnt_in + c, 1);
      atomicAdd(cnt_out + r, 1);
    }
    const bool carry = pd.in_eid != nullptr;
    k_in[e] = valid ? (int32_t)c : sentinel;
    v_in[e] = carry ? (int32_t)e : (int32_t)r;
    k_out[e] = valid ? (int32_t)r : sentinel;
    v_out[e] = carry ? (int32_t)e : (int32_t)c;
  }
  // hub-row lists of this direction: counters and arrival flags start at zero
  for (int o = 0; o < 2; ++o) {
    int32_t* lng = o ? pd.g.out_long : pd.g.in_long;
    if (lng == nullptr) continue;
    const int64_t hdr = long_hdr_ints(pd.E);
    for (int64_t i = t0; i < hdr; i += stride) lng[i] = 0;
  }
  if (dir == 0 && a.node_ptr != nullptr) {
    for (int64_t i = t0; i <= a.N; i += stride) {
      const int64_t prev = i > 0 ? a.batch[i - 1] : -1;
      const int64_t cur = i < a.N ? a.batch[i] : a.B;
      if (i < a.N && (cur < 0 || cur >= a.B)) atomicOr(a.flags, BIGCN_FLAG_BATCH_ORDER);
      if (cur < prev) atomicOr(a.flags, BIGCN_FLAG_BATCH_ORDER);
      int64_t lo = prev + 1 < 0 ? 0 : prev + 1;
      int64_t hi = cur > a.B ? a.B : cur;
      for (int64_t b = lo; b <= hi; ++b) a.node_ptr[b] = (int32_t)i;
    }
  }
}

// ---- exclusive scan of the 2*ndir count arrays -> ptr arrays --------------------------
__device__ __forceinline__ int32_t* ptr_array(const PrepArgs& a, int s) {
  const bigcn_graph_t& g = a.d[s >> 1].g;
  return (s & 1) ? g.out_ptr : g.in_ptr;
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  // 256 threads; returns exclusive prefix of v, *total = block sum
  __shared__ int warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect warp_tot reuse across calls
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int pre = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < SCAN_THREADS / 32; ++i) {
    int t = warp_tot[i];
    if (i < w) pre += t;
    tot += t;
  }
  *total = tot;
  return pre + inc - v;
}

__global__ void k_scan_blocksum(PrepArgs a) {
  const int s = blockIdx.y;
  const int32_t* cnt = a.cnt + (int64_t)s * a.N;
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
  int sum = 0;
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; ++it) {
    const int64_t i = base + it * SCAN_THREADS + threadIdx.x;
    if (i < a.N) sum += cnt[i];
  }
  int tot;
  block_exclusive_scan(sum, &tot);
  if (threadIdx.x == 0) a.bsum[(int64_t)s * a.nscanblk + blockIdx.x] = tot;
}

__global__ void k_scan_apply(PrepArgs a) {
  const int s = blockIdx.y;
  const int32_t* cnt = a.cnt + (int64_t)s * a.N;
  int32_t* ptr = ptr_array(a, s);
  // prefix over the earlier blocks' sums
  int part = 0;
  for (int b = threadIdx.x; b < (int)blockIdx.x; b += SCAN_THREADS)
    part += a.bsum[(int64_t)s * a.nscanblk + b];
  int block_prefix;
  block_exclusive_scan(part, &block_prefix);
  // each thread owns SCAN_ITEMS consecutive elements
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int tsum = 0;
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; ++it) {
    const int64_t i = base + it;
    v[it] = i < a.N ? cnt[i] : 0;
    tsum += v[it];
  }
  int tot;
  int run = block_prefix + block_exclusive_scan(tsum, &tot);
#pragma unroll
  for (int it = 0; it < SCAN_ITEMS; ++it) {
    const int64_t i = base + it;
    if (i < a.N) ptr[i] = run;
    run += v[it];
    if (i == a.N - 1) ptr[a.N] = run;
  }
  if (a.N == 0 && blockIdx.x == 0 && threadIdx.x == 0) ptr[0] = 0;
}

// ---- stable LSD radix sort, 8-bit digits ------------------------------------------------
__global__ void k_rs_hist(PrepArgs a, int src, int shift) {
  const int s = blockIdx.y;
  const int64_t E = a.d[s >> 1].E;
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  if (base < E) {
    const int32_t* keys = a.keys[src] + (int64_t)s * a.Emax;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
      const int64_t e = base + r * RS_THREADS + threadIdx.x;
      if (e < E) atomicAdd(&h[(keys[e] >> shift) & 255], 1);
    }
  }
  __syncthreads();
  a.hist[((int64_t)s * 256 + threadIdx.x) * a.nblk + blockIdx.x] = h[threadIdx.x];
}

__global__ void k_rs_scan(PrepArgs a) {
  // one block per sort; thread d owns digit d
  const int s = blockIdx.x;
  int32_t* row = a.hist + ((int64_t)s * 256 + threadIdx.x) * a.nblk;
  int tot = 0;
  for (int b = 0; b < a.nblk; ++b) tot += row[b];
  int dummy;
  int run = block_exclusive_scan(tot, &dummy);
  for (int b = 0; b < a.nblk; ++b) {
    int t = row[b];
    row[b] = run;
    run += t;
  }
}

__global__ void k_rs_scatter(PrepArgs a, int src, int shift, int last) {
  const int s = blockIdx.y;
  const int64_t E = a.d[s >> 1].E;
  const int64_t base = (int64_t)blockIdx.x * RS_TILE;
  if (base >= E) return;
  const int32_t* keys = a.keys[src] + (int64_t)s * a.Emax;
  const int32_t* vals = a.vals[src] + (int64_t)s * a.Emax;
  int32_t* keys_o = a.keys[src ^ 1] + (int64_t)s * a.Emax;
  int32_t* vals_o = a.vals[src ^ 1] + (int64_t)s * a.Emax;
  int32_t* eid_o = nullptr;
  const int64_t* endpoint = nullptr;
  if (last) {
    const PrepDir& pd = a.d[s >> 1];
    vals_o = (s & 1) ? pd.g.out_idx : pd.g.in_idx;
    if (pd.in_eid != nullptr) {
      eid_o = (s & 1) ? pd.out_eid : pd.in_eid;
      endpoint = pd.ei + ((s & 1) ? pd.E : 0);   // by-source CSR lists targets, by-target CSR lists sources
    }
  }
  __shared__ int dbase[256];
  __shared__ int wcnt[RS_THREADS / 32][257];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  dbase[threadIdx.x] = a.hist[((int64_t)s * 256 + threadIdx.x) * a.nblk + blockIdx.x];
  for (int r = 0; r < RS_ROUNDS; ++r) {
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 257; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t e = base + r * RS_THREADS + threadIdx.x;
    const bool valid = e < E;
    const int32_t key = valid ? keys[e] : 0;
    const int32_t val = valid ? vals[e] : 0;
    const int digit = valid ? ((key >> shift) & 255) : 256;
    const unsigned peers = __match_any_sync(FULL_MASK, digit);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    if (rank == 0) wcnt[w][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      int pre = 0;
      for (int i = 0; i < w; ++i) pre += wcnt[i][digit];
      const int pos = dbase[digit] + pre + rank;
      if (!last) keys_o[pos] = key;
      if (eid_o != nullptr) {
        eid_o[pos] = val;
        vals_o[pos] = (int32_t)endpoint[val];
      } else {
        vals_o[pos] = val;
      }
    }
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int i = 0; i < RS_THREADS / 32; ++i) tot += wcnt[i][threadIdx.x];
    dbase[threadIdx.x] += tot;
    __syncthreads();
  }
}

// ---- finalize: deg, dis, (rowsum) --------------------------------------------------------
__global__ void k_prep_deg(PrepArgs a) {
  const int dir = blockIdx.y;
  const int32_t* c = a.cnt + (int64_t)(2 * dir + (a.deg_by == BIGCN_DEG_BY_SOURCE ? 1 : 0)) * a.N;
  const bigcn_graph_t& g = a.d[dir].g;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int d = c[i] + 1;
    g.deg[i] = d;
    // torch CPU deg.pow(-0.5) == 1.0f / sqrtf(deg), IEEE sqrt and divide (not rsqrtf)
    g.dis[i] = __fdiv_rn(1.0f, __fsqrt_rn((float)d));
  }
}

// hub rows of the four CSRs (blockIdx.y = 2*dir + orient): rows with more than LONG_ROW entries
// get a slot and a run of (row, chunk) items; slot / item numbering depends on the arrival
// order of the atomics, the sums formed from them do not.
__global__ void k_prep_long(PrepArgs a) {
  const int s = blockIdx.y;
  const PrepDir& pd = a.d[s >> 1];
  int32_t* lng = (s & 1) ? pd.g.out_long : pd.g.in_long;
  if (lng == nullptr) return;
  const int32_t* ptr = (s & 1) ? pd.g.out_ptr : pd.g.in_ptr;
  const LongView L = long_view(lng, pd.E);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int len = ptr[i + 1] - ptr[i];
    if (len <= LONG_ROW) continue;
    int lch, nch;
    long_chunking(len, lch, nch);
    const int slot = atomicAdd(&L.cnt[0], 1);
    const int t0 = atomicAdd(&L.cnt[1], nch);
    L.row[slot] = (int32_t)i;
    L.item0[slot] = t0;
    for (int c = 0; c < nch; ++c) L.item_slot[t0 + c] = slot;
  }
}

__global__ void k_prep_rowsum(PrepArgs a) {
  // warp per row, COO' order: in-edges sequentially, then the self-loop
  const int dir = blockIdx.y;
  const bigcn_graph_t& g = a.d[dir].g;
  if (g.rowsum == nullptr) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarp = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < a.N; i += nwarp) {
    const int s = g.in_ptr[i], e = g.in_ptr[i + 1];
    const float di = g.dis[i];
    float acc = 0.f;
    for (int b = s; b < e; b += 32) {
      const int n = min(32, e - b);
      float p = 0.f;
      if (lane < n) p = __fmul_rn(g.dis[g.in_idx[b + lane]], di);
      for (int l = 0; l < n; ++l) acc = __fadd_rn(acc, __shfl_sync(FULL_MASK, p, l));
    }
    acc = __fadd_rn(acc, __fmul_rn(di, di));
    if (lane == 0) g.rowsum[i] = acc;
  }
}

static int radix_passes(int64_t N) {
  int bits = 1;
  while (((int64_t)1 << bits) <= N) ++bits;  // keys range over [0, N] (N = sentinel)
  return (bits + 7) / 8;
}

struct PrepLayout {
  size_t cnt, keys0, keys1, vals0, vals1, hist, bsum, total;
};
static PrepLayout prep_layout(int64_t N, int64_t Emax, int ndir) {
  PrepLayout L;
  const int ns = 2 * ndir;
  const int64_t nblk = ceil_div(Emax > 0 ? Emax : 1, RS_TILE);
  const int64_t nsb = ceil_div(N > 0 ? N : 1, SCAN_TILE);
  Carver d(nullptr, 0);
  d.take<int32_t>((size_t)ns * N); L.cnt = d.off - (size_t)ns * N * 4;
  d.take<int32_t>((size_t)ns * Emax); L.keys0 = d.off - (size_t)ns * Emax * 4;
  d.take<int32_t>((size_t)ns * Emax); L.keys1 = d.off - (size_t)ns * Emax * 4;
  d.take<int32_t>((size_t)ns * Emax); L.vals0 = d.off - (size_t)ns * Emax * 4;
  d.take<int32_t>((size_t)ns * Emax); L.vals1 = d.off - (size_t)ns * Emax * 4;
  d.take<int32_t>((size_t)ns * 256 * nblk); L.hist = d.off - (size_t)ns * 256 * nblk * 4;
  d.take<int32_t>((size_t)ns * nsb); L.bsum = d.off - (size_t)ns * nsb * 4;
  L.total = align_up(d.off, 256);
  return L;
}

int graph_prep_impl(int32_t n_dirs, const int64_t* const* edge_index, const int64_t* E, int64_t N,
                    const int64_t* batch, int64_t B, int32_t deg_by, const bigcn_graph_t* graphs,
                    int32_t* node_ptr, int32_t* flags, void* workspace, size_t workspace_bytes,
                    cudaStream_t st, int32_t* const* eids) {
  BIGCN_CHECK_ARG(n_dirs == 1 || n_dirs == 2, "graph_prep: n_dirs must be 1 or 2");
  BIGCN_CHECK_ARG(N >= 0 && N < (1ll << 31) - 1, "graph_prep: N out of int32 range");
  int64_t Emax = 0;
  for (int d = 0; d < n_dirs; ++d) {
    BIGCN_CHECK_ARG(E[d] >= 0 && E[d] < (1ll << 31) - 1, "graph_prep: E out of int32 range");
    if (E[d] > Emax) Emax = E[d];
  }
  BIGCN_CHECK_ARG(flags != nullptr, "graph_prep: flags is NULL");
  BIGCN_CHECK_ARG((node_ptr == nullptr) || (batch != nullptr && B >= 0), "graph_prep: node_ptr needs batch");
  const PrepLayout L = prep_layout(N, Emax, n_dirs);
  BIGCN_CHECK_ARG(workspace_bytes >= L.total, "graph_prep: workspace too small (%zu < %zu)",
                  workspace_bytes, L.total);
  char* ws = reinterpret_cast<char*>(workspace);
  PrepArgs a{};
  a.ndir = n_dirs;
  a.deg_by = deg_by;
  a.N = N;
  a.B = B;
  a.batch = batch;
  a.node_ptr = node_ptr;
  a.flags = flags;
  for (int d = 0; d < n_dirs; ++d) {
    a.d[d].ei = edge_index[d];
    a.d[d].E = E[d];
    a.d[d].g = graphs[d];
    a.d[d].in_eid = eids ? eids[2 * d] : nullptr;
    a.d[d].out_eid = eids ? eids[2 * d + 1] : nullptr;
  }
  a.cnt = reinterpret_cast<int32_t*>(ws + L.cnt);
  a.keys[0] = reinterpret_cast<int32_t*>(ws + L.keys0);
  a.keys[1] = reinterpret_cast<int32_t*>(ws + L.keys1);
  a.vals[0] = reinterpret_cast<int32_t*>(ws + L.vals0);
  a.vals[1] = reinterpret_cast<int32_t*>(ws + L.vals1);
  a.hist = reinterpret_cast<int32_t*>(ws + L.hist);
  a.bsum = reinterpret_cast<int32_t*>(ws + L.bsum);
  a.Emax = Emax;
  a.nblk = (int32_t)ceil_div(Emax > 0 ? Emax : 1, RS_TILE);
  a.nscanblk = (int32_t)ceil_div(N > 0 ? N : 1, SCAN_TILE);
  const int ns = 2 * n_dirs;

  if (N > 0) cudaMemsetAsync(a.cnt, 0, (size_t)ns * N * sizeof(int32_t), st);
  {
    const int64_t work = (Emax > N + 1 ? Emax : N + 1);
    int blocks = (int)ceil_div(work, 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_prep_count<<<dim3(blocks, n_dirs), 256, 0, st>>>(a);
    BIGCN_CHECK_LAUNCH("k_prep_count");
  }
  k_scan_blocksum<<<dim3(a.nscanblk, ns), SCAN_THREADS, 0, st>>>(a);
  BIGCN_CHECK_LAUNCH("k_scan_blocksum");
  k_scan_apply<<<dim3(a.nscanblk, ns), SCAN_THREADS, 0, st>>>(a);
  BIGCN_CHECK_LAUNCH("k_scan_apply");
  if (Emax > 0) {
    const int passes = radix_passes(N);
    for (int p = 0; p < passes; ++p) {
      const int src = p & 1;
      k_rs_hist<<<dim3(a.nblk, ns), RS_THREADS, 0, st>>>(a, src, 8 * p);
      BIGCN_CHECK_LAUNCH("k_rs_hist");
      k_rs_scan<<<ns, 256, 0, st>>>(a);
      BIGCN_CHECK_LAUNCH("k_rs_scan");
      k_rs_scatter<<<dim3(a.nblk, ns), RS_THREADS, 0, st>>>(a, src, 8 * p, p == passes - 1);
      BIGCN_CHECK_LAUNCH("k_rs_scatter");
    }
  }
  if (N > 0) {
    int blocks = (int)ceil_div(N, 256);
    const int cap = num_sms() * 8;
    if (blocks > cap) blocks = cap;
    k_prep_deg<<<dim3(blocks, n_dirs), 256, 0, st>>>(a);
    BIGCN_CHECK_LAUNCH("k_prep_deg");
    bool want_long = false;
    for (int d = 0; d < n_dirs; ++d) want_long |= graphs[d].in_long != nullptr || graphs[d].out_long != nullptr;
    if (want_long && Emax > LONG_ROW) {
      k_prep_long<<<dim3(blocks, ns), 256, 0, st>>>(a);
      BIGCN_CHECK_LAUNCH("k_prep_long");
    }
    bool want_rowsum = false;
    for (int d = 0; d < n_dirs; ++d) want_rowsum |= graphs[d].rowsum != nullptr;
    if (want_rowsum) {
      int rb = (int)ceil_div(N, 8);
      if (rb > cap) rb = cap;
      k_prep_rowsum<<<dim3(rb, n_dirs), 256, 0, st>>>(a);
      BIGCN_CHECK_LAUNCH("k_prep_rowsum");
    }
  }
  return 0;
}

size_t long_ws_ints(int64_t E) { return (size_t)long_ws_ints_impl(E > 0 ? E : 0); }

size_t graph_prep_ws_bytes(int64_t N, int64_t Emax, int ndir) {
  return prep_layout(N, Emax, ndir).total;
}

}  // namespace bigcn

extern "C" size_t bigcn_graph_prep_workspace_bytes(int64_t N, int64_t E_max, int32_t n_dirs) {
  return bigcn::graph_prep_ws_bytes(N, E_max, n_dirs);
}

extern "C" size_t bigcn_long_ws_ints(int64_t E) { return bigcn::long_ws_ints(E); }

extern "C" int bigcn_graph_prep(int32_t n_dirs, const int64_t* const* edge_index, const int64_t* E,
                                int64_t N, const int64_t* batch, int64_t B, int32_t deg_by,
                                const bigcn_graph_t* graphs, int32_t* node_ptr, int32_t* flags,
                                void* workspace, size_t workspace_bytes, bigcn_stream_t stream) {
  return bigcn::graph_prep_impl(n_dirs, edge_index, E, N, batch, B, deg_by, graphs, node_ptr, flags,
                                workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}
