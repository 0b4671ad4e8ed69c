// wavelet.cuh -- the 2-D dominance index of the engine: a wavelet matrix over the "link" value of
// each point, points ordered by their x coordinate.
//
// It replaces all three dominance structures of the reference (SparsePrefixMatrices.jl:396-821:
// b-ary DominanceCount, BinaryDominanceCount, SparseStepwiseDominanceCount) -- the reference's
// hints only choose between CPU data structures; the counts are what must match.
//
// Layout in HBM (one allocation):   level l (0 = most significant bit), block b:
//     blocks[(l * nblk + b) * 8 + 0]      = number of 0-bits of level l before element b*224
//     blocks[(l * nblk + b) * 8 + 1..7]   = the 224 bits of elements b*224 .. b*224+223
// i.e. one 32-byte sector answers one rank query.  z[l] = total 0-bits of level l.
#pragma once
#include "common.cuh"

namespace cpb {

static constexpr int WM_BLOCK = 224;       // elements per 32-byte rank block
#ifndef CPB_WM_WARPS
#define CPB_WM_WARPS 8
#endif
static constexpr int WM_TILE_BLOCKS = CPB_WM_WARPS;  // rank blocks per build tile (one per warp of the build CTA)
static constexpr int WM_TILE = WM_BLOCK * WM_TILE_BLOCKS;  // 1792 elements per build CTA

struct DevWM {
  const u32* blocks;  // [L][nblk][8]
  const u32* z;       // [L]
  const u32* S0;      // [2^L]  value-only half of the rank descent (see wm_rank_lt)
  u32 nblk;
  u32 npts;
  int L;
};

// A rank structure = wavelet matrix + the prefix array over x:  P[x] = #{points with x_p < x},
// x in 1..n+1 (1-based column numbers).
struct DevRank {
  DevWM wm;
  const u32* P;  // indexable for 1 <= x <= n+1
};

#ifdef __CUDACC__
// zeros among the first p elements of level l
__device__ __forceinline__ u32 wm_rank0(const DevWM& w, int l, u32 p) {
  const u32 b = p / WM_BLOCK;
  const u32 off = p - b * WM_BLOCK;
  const u32* blk = w.blocks + ((size_t)l * w.nblk + b) * 8;
  u32 hdr, wd[7];
  // one 256-bit load = exactly one 32-byte sector = one L1 wavefront per lane (LDG.E.256 on sm_100a)
  asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=r"(hdr), "=r"(wd[0]), "=r"(wd[1]), "=r"(wd[2]), "=r"(wd[3]), "=r"(wd[4]), "=r"(wd[5]), "=r"(wd[6])
      : "l"(blk));
  u32 ones = 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int nb = (int)off - 32 * k;
    const u32 mask = nb >= 32 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
    ones += __popc(wd[k] & mask);
  }
  return hdr + off - ones;
}

// sum over the 1-bits of v of the zeros left of the descending position, starting from e
__device__ __forceinline__ u32 wm_descend(const DevWM& w, u32 e, u32 v) {
  u32 cnt = 0;
  for (int l = 0; l < w.L; ++l) {
    const u32 e0 = wm_rank0(w, l, e);
    if ((v >> (w.L - 1 - l)) & 1u) {
      cnt += e0;
      e = __ldg(w.z + l) + (e - e0);
    } else {
      e = e0;
    }
  }
  return cnt;
}

// #{ p < e : val_p < v }  for 0 <= e <= npts, any v >= 0.
// The textbook descent tracks the interval [s, e) with s starting at 0; s depends on v alone, so its
// contribution S0[v] = wm_descend(0, v) is tabulated at build time and a query walks ONE chain of L
// dependent 32-byte loads instead of two.
__device__ __forceinline__ u32 wm_rank_lt(const DevWM& w, u32 e, u32 v) {
  if (w.L < 32 && (v >> w.L) != 0) return e;  // v beyond the value range: everything is smaller
  const u32 s0 = __ldg(w.S0 + v);
  return wm_descend(w, e, v) - s0;
}

// #{ points : x_p < jp, val_p >= j }   (the reference's lnk(n+2-j, j'), SparseColorArrays.jl:121-125)
__device__ __forceinline__ u32 rank_count_ge(const DevRank& r, u32 j, u32 jp) {
  const u32 e = __ldg(r.P + jp);
  return e - wm_rank_lt(r.wm, e, j);
}
#endif

// Host-side owner of a wavelet matrix.
struct WaveletMatrix {
  DBuf<u32> blocks;
  DBuf<u32> z;
  DBuf<u32> S0;
  u32 nblk = 0, npts = 0;
  int L = 0;
  // Builds from vals[0..n) (values <= max_value).  `vals` is consumed (used as a ping-pong buffer).
  void build(u32* vals, u32* scratch, size_t n, u64 max_value);
  DevWM dev() const { return DevWM{blocks.get(), z.get(), S0.get(), nblk, npts, L}; }
  size_t bytes() const { return (size_t)L * nblk * 32; }  // rank blocks only
};

}  // namespace cpb
