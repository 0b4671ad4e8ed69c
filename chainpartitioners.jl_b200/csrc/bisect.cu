// bisect.cu -- kernel family "probe": BisectCostBottleneckSplitter / LazyBisectCostBottleneckSplitter
// (BisectCostBottleneckSplitter.jl:6-63, LazyBisectCostBottleneckSplitter.jl:8-70,140-258,260-388).
//
// The reference bisects the cost value c = (c_lo + c_hi) / 2 and runs one greedy feasibility probe
// per threshold: part k extends as far right as c(spl[k], j', k) <= c allows.  The streaming
// ("lazy") and the random-access form return the same split vectors for monotone costs, so both
// map onto one engine here:
//   * the next `depth` levels of the bisection tree (2^depth - 1 thresholds, each computed with the
//     reference's own double expression) are probed CONCURRENTLY, one CTA per threshold, and the
//     tree is then walked by feasibility -- the threshold sequence, and therefore the returned
//     spl_hi, is identical to the sequential loop by construction;
//   * inside a probe, each part boundary is found by a multi-way search over a thread-block CLUSTER
//     (8 CTAs x 128 threads = 1024 candidates j' per round, every thread evaluates the oracle at one
//     candidate; counts are exchanged through distributed shared memory, one cluster barrier per
//     round) instead of the reference's one-query-at-a-time binary search.  Spreading a threshold
//     over 8 SMs matters: the rank descents are one random 32-byte sector per level and candidate,
//     and a single SM's L1 serialises them (ncu, profiles/r01).  Previous probes' windows
//     (spl_lo/spl_hi, BisectCost...:29-33,46,55,58) are used as search hints only; a miss falls
//     back to the full range, so the result never depends on them.
#include <algorithm>
#include <cmath>
#include <cooperative_groups.h>
#include "engine.cuh"

namespace cg = cooperative_groups;

namespace cpb {

static constexpr int BS_THREADS = 128;
static constexpr int BS_CLUSTER = 8;
static constexpr int BS_WIDTH = BS_THREADS * BS_CLUSTER;  // candidates per round
static constexpr int BS_MAX_DEPTH = 4;

struct BisectState {
  double c_lo, c_hi;
  int done;
  int probes;
  int rounds;
  int _pad;
};

// largest x in [a, b] with c(j, x) <= c, or a-1 if c(j, a) > c.  Monotone predicate.  Cluster-wide:
// every CTA of the cluster calls it with the same arguments and gets the same answer.
template <class T>
__device__ __forceinline__ i64 wide_search(const DevOracle& o, cg::cluster_group& cluster, int (*s_cnt)[BS_CLUSTER], int& phase,
                                           u32 j, i64 a, i64 b, double c) {
  const unsigned crank = cluster.block_rank();
  while (true) {
    const i64 S = b - a + 1;
    if (S <= 0) return a - 1;
    const i64 stride = (S + BS_WIDTH - 1) / BS_WIDTH;
    const i64 x = a + (i64)(crank * BS_THREADS + threadIdx.x) * stride;
    bool ok = false;
    if (x <= b) ok = cost_leq(dev_cost<T>(o, j, (u32)x), c);
    const int mine = __syncthreads_count(ok);
    if (threadIdx.x < BS_CLUSTER) {  // push my count into slot [crank] of every CTA of the cluster
      int* remote = cluster.map_shared_rank(&s_cnt[phase][crank], threadIdx.x);
      *remote = mine;
    }
    cluster.sync();
    int ct = 0;
#pragma unroll
    for (int p = 0; p < BS_CLUSTER; ++p) ct += s_cnt[phase][p];
    phase ^= 1;  // double-buffered: a fast CTA may already be writing the next round's counts
    if (ct == 0) return a - 1;
    const i64 base = a + (i64)(ct - 1) * stride;
    if (stride == 1) return base;
    a = base + 1;
    b = min(b, base + stride - 1);
  }
}

template <class T>
__global__ void __cluster_dims__(BS_CLUSTER, 1, 1) __launch_bounds__(BS_THREADS)
    k_bisect_round(const __grid_constant__ DevOracle o, int K, double eps1, const BisectState* __restrict__ st,
                   const int* __restrict__ hint_lo, const int* __restrict__ hint_hi, int* __restrict__ node_spl,
                   int* __restrict__ node_res, double* __restrict__ node_c) {
  __shared__ double s_c;
  __shared__ int s_valid;
  __shared__ int s_cnt[2][BS_CLUSTER];
  cg::cluster_group cluster = cg::this_cluster();
  const int node = blockIdx.x / BS_CLUSTER;
  const bool writer = cluster.block_rank() == 0 && threadIdx.x == 0;
  if (threadIdx.x == 0) {
    int valid = st->done ? 0 : 1;
    double lo = st->c_lo, hi = st->c_hi;
    // path from the root to this node in heap order: left child = "ancestor was feasible"
    int path[BS_MAX_DEPTH];
    int len = 0;
    for (int i = node; i > 0; i = (i - 1) >> 1) path[len++] = (i & 1);  // 1 = left child
    for (int t = len - 1; t >= 0 && valid; --t) {
      if (!(lo * eps1 < hi)) { valid = 0; break; }
      const double c = (lo + hi) / 2;
      if (path[t]) hi = c; else lo = c;
    }
    if (valid && !(lo * eps1 < hi)) valid = 0;
    s_valid = valid;
    s_c = (lo + hi) / 2;
  }
  __syncthreads();
  if (!s_valid) {  // uniform over the cluster: every CTA derives it from the same state
    if (writer) node_res[node] = 0;
    return;
  }
  const double c = s_c;
  const i64 n1 = (i64)o.n + 1;
  int* spl = node_spl + (size_t)node * (K + 2);  // 1-based, spl[1..K+1]
  if (writer) { spl[1] = 1; spl[K + 1] = (int)n1; }
  int phase = 0;
  i64 j = 1;
  bool broke = false;
  for (int k = 1; k <= K - 1; ++k) {
    i64 a = max(j, (i64)hint_lo[k + 1]);
    i64 b = min((i64)hint_hi[k + 1], n1);
    if (b < a) { a = j; b = n1; }
    i64 r = wide_search<T>(o, cluster, s_cnt, phase, (u32)j, a, b, c);
    if (r == a - 1 && a > j) r = wide_search<T>(o, cluster, s_cnt, phase, (u32)j, j, a - 1, c);
    else if (r == b && b < n1) r = wide_search<T>(o, cluster, s_cnt, phase, (u32)j, b + 1, n1, c);
    if (r < j) {  // even the empty part exceeds c (BisectCost...:47-51)
      broke = true;
      if (writer)
        for (int t = k + 1; t <= K; ++t) spl[t] = (int)j;
      break;
    }
    if (writer) spl[k + 1] = (int)r;
    j = r;
  }
  if (writer) {
    bool feas = false;
    if (!broke) feas = cost_leq(dev_cost<T>(o, (u32)j, (u32)n1), c);
    node_c[node] = c;
    node_res[node] = feas ? 2 : 1;
  }
  cluster.sync();  // no CTA may exit while peers can still write into its shared memory
}

// walks the probed subtree by feasibility (BisectCost...:53-59)
__global__ void k_bisect_advance(int K, int P, double eps1, BisectState* st, int* __restrict__ hint_lo, int* __restrict__ hint_hi,
                                 int* __restrict__ best, const int* __restrict__ node_spl, const int* __restrict__ node_res,
                                 const double* __restrict__ node_c) {
  __shared__ int s_node, s_res;
  if (st->done) return;
  int node = 0;
  while (node < P) {
    const int res = node_res[node];
    if (res == 0) {
      if (threadIdx.x == 0) st->done = 1;
      return;
    }
    const int* spl = node_spl + (size_t)node * (K + 2);
    if (res == 2) {
      for (int t = 1 + threadIdx.x; t <= K + 1; t += blockDim.x) { best[t] = spl[t]; hint_hi[t] = spl[t]; }
      if (threadIdx.x == 0) { st->c_hi = node_c[node]; st->probes += 1; }
      node = 2 * node + 1;
    } else {
      for (int t = 1 + threadIdx.x; t <= K + 1; t += blockDim.x) hint_lo[t] = spl[t];
      if (threadIdx.x == 0) { st->c_lo = node_c[node]; st->probes += 1; }
      node = 2 * node + 2;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    st->rounds += 1;
    if (!(st->c_lo * eps1 < st->c_hi)) st->done = 1;
  }
  (void)s_node; (void)s_res;
}

// pulls a read-only buffer into L2 ahead of the latency-bound probes (one prefetch per 128-byte line)
__global__ void k_l2_prefetch(const char* __restrict__ p, size_t bytes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 128;
  for (size_t off = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128; off < bytes; off += stride)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

__global__ void k_bisect_init(int K, int n1, int* hint_lo, int* hint_hi, int* best) {
  for (int t = 1 + threadIdx.x; t <= K + 1; t += blockDim.x) {
    hint_lo[t] = (t == K + 1) ? n1 : 1;
    hint_hi[t] = (t == 1) ? 1 : n1;
    best[t] = (t == 1) ? 1 : n1;
  }
}

static int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s ? std::atoi(s) : dflt;
}

void solve_bisect(Oracle& f, bool lazy, double eps, i64 K, int64_t* h_spl_out) {
  CPB_REQUIRE(K >= 1, "K must be >= 1");
  CPB_REQUIRE(f.dev.kind != CPB_MODEL_BLOCK, "bisection needs a random-access oracle");
  const Matrix& A = *f.A;
  CPB_REQUIRE(A.n + 1 < ((i64)1 << 31) && K + 2 < ((i64)1 << 30), "problem too large for 32-bit split points");
  oracle_ensure_ranks(f);
  double bnd[2];
  oracle_bound(f, K, bnd);  // (c_lo, c_hi) ./ 1
  if (lazy) {
    // LazyBisect...:55-57 / :233-235: c_lo = max(c_lo, f(empty part k)) -- alpha for the affine models
    double a0 = f.mdl.coef[0];
    if (f.mdl.kind == CPB_MODEL_COLBLOCK) a0 = f.mdl.alpha_col[0];
    bnd[0] = std::max(bnd[0], a0);
  }
  ProfScope prof("probe");
  // the rank descents are dependent random sector reads: make them L2 hits when the index fits
  for (RankStruct* rs : {f.net.get(), f.dianet.get(), f.selfnet.get(), f.selfpin.get()})
    if (rs && rs->wm.bytes() > 0 && rs->wm.bytes() <= ((size_t)64 << 20) && env_int("CPB_L2_PREFETCH", 1)) {
      const size_t bytes = rs->wm.bytes();
      const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((bytes / 128 + 255) / 256, (size_t)ctx().sm_count * 8));
      CPB_LAUNCH(k_l2_prefetch, grid, 256, 0, (const char*)rs->wm.blocks.get(), bytes);
    }
  int depth = std::min(std::max(env_int("CPB_BISECT_DEPTH", 4), 1), BS_MAX_DEPTH);
  const int P = (1 << depth) - 1;
  const double eps1 = 1 + eps;
  DBuf<BisectState> st(1);
  DBuf<int> hint_lo(K + 2), hint_hi(K + 2), best(K + 2), node_spl((size_t)P * (K + 2)), node_res(P);
  DBuf<double> node_c(P);
  BisectState h{};
  h.c_lo = bnd[0];
  h.c_hi = bnd[1];
  h.done = !(h.c_lo * eps1 < h.c_hi);
  CPB_CUDA(cudaMemcpyAsync(st.get(), &h, sizeof(h), cudaMemcpyHostToDevice, ctx().stream));
  CPB_LAUNCH(k_bisect_init, 1, 256, 0, (int)K, (int)(A.n + 1), hint_lo.get(), hint_hi.get(), best.get());
  const int batch = std::max(1, 6 / depth);
  for (int guard = 0; guard < 4096 && !h.done; ++guard) {
    for (int r = 0; r < batch; ++r) {
      if (f.dev.is_float)
        CPB_LAUNCH(k_bisect_round<double>, P * BS_CLUSTER, BS_THREADS, 0, f.dev, (int)K, eps1, st.get(), hint_lo.get(), hint_hi.get(), node_spl.get(), node_res.get(), node_c.get());
      else
        CPB_LAUNCH(k_bisect_round<i64>, P * BS_CLUSTER, BS_THREADS, 0, f.dev, (int)K, eps1, st.get(), hint_lo.get(), hint_hi.get(), node_spl.get(), node_res.get(), node_c.get());
      CPB_LAUNCH(k_bisect_advance, 1, 256, 0, (int)K, P, eps1, st.get(), hint_lo.get(), hint_hi.get(), best.get(), node_spl.get(), node_res.get(), node_c.get());
    }
    CPB_CUDA(cudaMemcpyAsync(&h, st.get(), sizeof(h), cudaMemcpyDeviceToHost, ctx().stream));
    CPB_CUDA(cudaStreamSynchronize(ctx().stream));
  }
  CPB_REQUIRE(h.done, "bisection did not terminate (eps too small for Float64?)");
  std::vector<int> hb(K + 2);
  CPB_CUDA(cudaMemcpyAsync(hb.data(), best.get(), (K + 2) * sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
  CPB_CUDA(cudaStreamSynchronize(ctx().stream));
  for (i64 k = 1; k <= K + 1; ++k) h_spl_out[k - 1] = hb[k];
}

}  // namespace cpb
