// prefix.cu -- the reference's exported 2-D prefix structures on the device (SparsePrefixMatrices.jl):
//   DominanceCount / BinaryDominanceCount / SparseStepwiseDominanceCount   :396-821    C[i, j] = #{nonzeros (r, c): r <= i-1, c <= j-1}
//   DominanceSum / SparseStepwiseDominanceSum                              :1-392      S[i, j] = sum of their values
//   RookCount / BinaryRookCount / RookSum                                  :825-1273   the same over the N points (idx[q], q) of a permutation
//
// One structure serves all of them (the reference's hints and b / H / b' arguments only choose between CPU layouts; the
// counts are what must match): the wavelet matrix of wavelet.cu over the row of every nonzero in column order, and -- for the
// sums -- one exclusive prefix-sum array of the values per bit level, in the element order of that level.  The zeros of a
// level land in stable order at the front of the next level, so "the values of the 0-bit elements of [s, e)" is a difference
// of two entries of the next level's prefix sums; a query is the usual L-step descent with two rank lookups per level.
// Sums are 64-bit integers with wrap-around (the arithmetic of the reference's UInt / Int tests, test_SparsePrefixMatrices.jl:
// 43-69); Float64 sums depend on the summation order and are not offered.
#include <algorithm>
#include "engine.cuh"
#include "primitives.cuh"

namespace cpb {

// ------------------------------------------------------------------------------ 64-bit exclusive scan
static constexpr int S64_THREADS = 256;
static constexpr int S64_ITEMS = 8;
static constexpr int S64_TILE = S64_THREADS * S64_ITEMS;

__device__ __forceinline__ u64 block_excl_scan64(u64 x, u64* smem, u64& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  u64 inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u64 y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) smem[w] = inc;
  __syncthreads();
  u64 wbase = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < S64_THREADS / 32; ++k) {
    const u64 s = smem[k];
    if (k < w) wbase += s;
    tot += s;
  }
  __syncthreads();
  total = tot;
  return wbase + inc - x;
}

__global__ void __launch_bounds__(S64_THREADS) k_scan64_reduce(const u64* __restrict__ in, u64* __restrict__ sums, size_t n) {
  __shared__ u64 sm[S64_THREADS / 32];
  const size_t b0 = (size_t)blockIdx.x * S64_TILE;
  const size_t b1 = min(n, b0 + (size_t)S64_TILE);
  u64 s = 0;
  for (size_t i = b0 + threadIdx.x; i < b1; i += S64_THREADS) s += in[i];
  u64 tot;
  block_excl_scan64(s, sm, tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// out[i] = carry + in[0] + ... + in[i-1] for the tile's range; the thread after the last element also writes out[n] (the
// total) when it belongs to this tile.  in and out may not alias (out is one longer).
__global__ void __launch_bounds__(S64_THREADS) k_scan64_apply(const u64* __restrict__ in, u64* __restrict__ out, const u64* __restrict__ tile_off, size_t n) {
  __shared__ u64 sm[S64_THREADS / 32];
  const size_t b0 = (size_t)blockIdx.x * S64_TILE;
  const size_t i0 = b0 + (size_t)threadIdx.x * S64_ITEMS;
  u64 v[S64_ITEMS];
  u64 tsum = 0;
#pragma unroll
  for (int k = 0; k < S64_ITEMS; ++k) {
    v[k] = (i0 + k < n) ? in[i0 + k] : 0ull;
    tsum += v[k];
  }
  u64 tot;
  u64 ex = block_excl_scan64(tsum, sm, tot) + (tile_off ? tile_off[blockIdx.x] : 0ull);
#pragma unroll
  for (int k = 0; k < S64_ITEMS; ++k) {
    if (i0 + k <= n) out[i0 + k] = ex;  // index n receives the grand total
    ex += v[k];
  }
}

// out[0..n] (n + 1 entries) = exclusive prefix sums of in[0..n), wrap-around 64-bit
static void exclusive_scan_u64(const u64* in, u64* out, size_t n) {
  const size_t tiles = n / S64_TILE + 1;  // the tile that holds index n exists even when n is a multiple of the tile
  if (tiles == 1) {
    CPB_LAUNCH(k_scan64_apply, 1, S64_THREADS, 0, in, out, (const u64*)nullptr, n);
    return;
  }
  DBuf<u64> sums(tiles), offs(tiles + 1);
  CPB_LAUNCH(k_scan64_reduce, (unsigned)tiles, S64_THREADS, 0, in, sums.get(), n);
  exclusive_scan_u64(sums.get(), offs.get(), tiles);
  CPB_LAUNCH(k_scan64_apply, (unsigned)tiles, S64_THREADS, 0, in, out, offs.get(), n);
}

// ------------------------------------------------------------------------------ weights through the levels
// nxt[position of p in level l+1] = cur[p]: the rank blocks of level l give both the bit of p and the zeros before it
__global__ void __launch_bounds__(256) k_wm_carry_values(DevWM w, int l, const u64* __restrict__ cur, u64* __restrict__ nxt) {
  const u32 p = blockIdx.x * 256u + threadIdx.x;
  if (p >= w.npts) return;
  const u32 b = p / WM_BLOCK, off = p - b * WM_BLOCK;
  const u32 word = __ldg(w.blocks + ((size_t)l * w.nblk + b) * 8 + 1 + (off >> 5));
  const u32 z0 = wm_rank0(w, l, p);
  const u32 dst = ((word >> (off & 31)) & 1u) ? __ldg(w.z + l) + (p - z0) : z0;
  nxt[dst] = cur[p];
}

struct DevPrefix {
  DevWM wm;
  const u32* pos;   // [n+1] nonzeros before column j (0-based j); nullptr: one point per column (rook form)
  const u64* wsum;  // [(L+1)][N+1] or nullptr
  i64 m, n;
};

// C[i, j] and S[i, j] for 1 <= i <= m+1, 1 <= j <= n+1 (SparsePrefixMatrices.jl:187-254, 537-604, 660-689, 957-1021, 1137-1206, 1246-1273)
__global__ void __launch_bounds__(256) k_prefix_query(DevPrefix d, i64 Q, const i64* __restrict__ qi, const i64* __restrict__ qj, i64* __restrict__ count_out,
                                                      i64* __restrict__ sum_out) {
  const i64 t = (i64)blockIdx.x * 256 + threadIdx.x;
  if (t >= Q) return;
  const u32 v = (u32)(qi[t] - 1);  // rows (0-based) below v are dominated
  const u32 j0 = (u32)(qj[t] - 1);
  const u32 e = d.pos ? __ldg(d.pos + j0) : j0;
  const DevWM& w = d.wm;
  const bool all = (w.L < 32 && (v >> w.L) != 0) || e == 0;
  if (count_out) count_out[t] = all ? (i64)e : (i64)wm_rank_lt(w, e, v);
  if (sum_out) {
    const size_t stride = (size_t)w.npts + 1;
    u64 acc = 0;
    if (all) {
      acc = d.wsum[e];
    } else {
      u32 s = 0, ee = e;
      for (int l = 0; l < w.L; ++l) {
        const u32 e0 = wm_rank0(w, l, ee);
        const u32 s0 = s ? wm_rank0(w, l, s) : 0u;
        if ((v >> (w.L - 1 - l)) & 1u) {
          const u64* W = d.wsum + (size_t)(l + 1) * stride;
          acc += W[e0] - W[s0];
          const u32 z = __ldg(w.z + l);
          s = z + (s - s0);
          ee = z + (ee - e0);
        } else {
          s = s0;
          ee = e0;
        }
      }
    }
    sum_out[t] = (i64)acc;
  }
}

std::unique_ptr<PrefixMatrix> prefix_build(i64 m, i64 n, i64 N, const i64* d_pos, const i64* d_idx, const i64* d_val) {
  CPB_REQUIRE(m >= 0 && n >= 0 && N >= 0, "negative dimension");
  CPB_REQUIRE(N + n + 1 < ((i64)1 << 31) && m < ((i64)1 << 31) - 2, "too many points for the 32-bit device index");
  CPB_REQUIRE(d_pos || n == N, "the rook form needs one point per column (n == N)");
  auto P = std::make_unique<PrefixMatrix>();
  P->m = m; P->n = n; P->N = N;
  DBuf<u32> flags(1);
  flags.zero();
  if (d_pos) {
    P->pos.alloc((size_t)n + 1);
    narrow_minus1(d_pos, P->pos.get(), (size_t)n + 1, 1, N + 1, flags.get());
    check_monotone(P->pos.get(), (size_t)n + 1, flags.get());
  }
  DBuf<u32> keys((size_t)std::max<i64>(N, 1)), scratch((size_t)std::max<i64>(N, 1));
  narrow_minus1(d_idx, keys.get(), (size_t)N, 1, m, flags.get());
  u32 hf = 0;
  i64 ends[2] = {1, N + 1};
  CPB_CUDA(cudaMemcpyAsync(&hf, flags.get(), sizeof(u32), cudaMemcpyDeviceToHost, ctx().stream));
  if (d_pos) {
    CPB_CUDA(cudaMemcpyAsync(&ends[0], d_pos, sizeof(i64), cudaMemcpyDeviceToHost, ctx().stream));
    CPB_CUDA(cudaMemcpyAsync(&ends[1], d_pos + n, sizeof(i64), cudaMemcpyDeviceToHost, ctx().stream));
  }
  CPB_CUDA(cudaStreamSynchronize(ctx().stream));
  CPB_REQUIRE((hf & 1u) == 0, "pos/idx entry out of range (expected 1-based indices)");
  CPB_REQUIRE((hf & 2u) == 0, "pos is not non-decreasing");
  CPB_REQUIRE(ends[0] == 1 && ends[1] == N + 1, "pos[1] must be 1 and pos[n+1] must be N+1");
  P->wm.build(keys.get(), scratch.get(), (size_t)N, (u64)std::max<i64>(m - 1, 0));
  if (d_val) {
    const int L = P->wm.L;
    const size_t stride = (size_t)N + 1;
    P->wsum.alloc((size_t)(L + 1) * stride);
    ProfScope prof("build_dominance_sums", (double)(L + 1) * (double)N * 32.0);
    DBuf<u64> a((size_t)std::max<i64>(N, 1)), b((size_t)std::max<i64>(N, 1));
    const u64* cur = (const u64*)d_val;
    u64* bufs[2] = {a.get(), b.get()};
    u64* const W = (u64*)P->wsum.get();
    exclusive_scan_u64(cur, W, (size_t)N);
    for (int l = 0; l < L; ++l) {
      u64* nxt = bufs[l & 1];
      if (N > 0) CPB_LAUNCH(k_wm_carry_values, (unsigned)((N + 255) / 256), 256, 0, P->wm.dev(), l, cur, nxt);
      exclusive_scan_u64(nxt, W + (size_t)(l + 1) * stride, (size_t)N);
      cur = nxt;
    }
  }
  return P;
}

void prefix_query(const PrefixMatrix& P, i64 Q, const i64* d_i, const i64* d_j, i64* d_count, i64* d_sum) {
  if (Q == 0) return;
  CPB_REQUIRE(!d_sum || P.wsum.get(), "this prefix matrix was built without values");
  DevPrefix d{P.wm.dev(), P.pos.get(), (const u64*)P.wsum.get(), P.m, P.n};
  ProfScope prof("prefix_query", (double)Q * (24.0 + (double)P.wm.L * 64.0));
  CPB_LAUNCH(k_prefix_query, (unsigned)((Q + 255) / 256), 256, 0, d, Q, d_i, d_j, d_count, d_sum);
}

}  // namespace cpb
