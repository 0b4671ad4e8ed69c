// wavelet.cu -- construction of the wavelet-matrix dominance index (kernel "build_dominance").
//
// One launch per bit level: every CTA (8 warps) owns a tile of 1792 consecutive elements (8 rank blocks, one per warp),
// emits the level's bit-plane + running zero counts (one 32-byte block per 224 elements) and
// stably partitions its elements into the next level's order.  The per-tile zero counts a level
// needs are accumulated by the PREVIOUS level's launch (warp-aggregated atomics on the
// destination tile), so each level reads the keys once and writes them once:
//     algorithmic bytes per level = n * (4 + 4) + n / 7 * (32 / 32) ...  
#include <algorithm>
#include "primitives.cuh"
#include "wavelet.cuh"

namespace cpb {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int WM_THREADS = 32 * WM_TILE_BLOCKS;  // one 224-element rank block (7 rounds of 32 lanes) per warp
static constexpr int WM_ROUNDS = 7;

// zeros of bit `bit` per tile (first level only)
__global__ void __launch_bounds__(256) k_wm_count(const u32* __restrict__ cur, u32 n, int bit, u32* __restrict__ tile_zeros) {
  __shared__ u32 sm[8];
  const size_t base = (size_t)blockIdx.x * WM_TILE;
  u32 c = 0;
  for (int i = threadIdx.x; i < WM_TILE; i += 256) {
    const size_t idx = base + i;
    if (idx < n) c += ((cur[idx] >> bit) & 1u) ^ 1u;
  }
  c = __reduce_add_sync(FULL, c);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0;
    for (int k = 0; k < 8; ++k) t += sm[k];
    tile_zeros[blockIdx.x] = t;
  }
}

// tile_off: exclusive scan of per-tile zero counts, tile_off[tiles] = z (total zeros of the level)
__global__ void __launch_bounds__(WM_THREADS, 2048 / WM_THREADS) k_wm_level(const u32* __restrict__ cur, u32* __restrict__ nxt, u32 n, int bit, int next_bit,
                                                            const u32* __restrict__ tile_off, u32 tiles, u32* __restrict__ next_tile_zeros,
                                                            u32* __restrict__ blocks, u32 nblk, u32* __restrict__ z_out) {
  __shared__ u32 s_wz[32];
  __shared__ u32 s_next[4];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const u32 tile = blockIdx.x;
  const size_t tbase = (size_t)tile * WM_TILE;
  const u32 wbase = w * WM_BLOCK;  // element offset of this warp's rank block inside the tile
  if (tid < 4) s_next[tid] = 0;

  // ---- full tiles (all but the last): the level is issue-bound (ncu: 88 % issue slots busy at 40 % of dram peak), so
  //      this path keeps the per-element instruction count down -- no bounds predicates, the zero masks are the
  //      complements of the one masks, and the next level's per-tile zero counts come from warp-uniform popcounts (a
  //      warp's zeros / ones land in one contiguous destination range; only a warp whose range straddles a
  //      destination-tile boundary looks at individual lanes).
  if (tbase + WM_TILE <= (size_t)n) {
    const u32* src = cur + tbase + wbase + lane;
    u32 v[WM_ROUNDS], ones[WM_ROUNDS];
#pragma unroll
    for (int r = 0; r < WM_ROUNDS; ++r) v[r] = src[r * 32];
    const u32 Z0 = tile_off[tile];
    const u32 ztot = tile_off[tiles];
    u32 wo = 0;
#pragma unroll
    for (int r = 0; r < WM_ROUNDS; ++r) {
      ones[r] = __ballot_sync(FULL, (v[r] >> bit) & 1u);
      wo += __popc(ones[r]);
    }
    const u32 wz = WM_BLOCK - wo;
    if (lane == 0) s_wz[w] = wz;
    __syncthreads();
    const u32 zbefore = __reduce_add_sync(FULL, lane < w ? s_wz[lane] : 0u);  // zeros of the tile before this warp
    if (tile == 0 && tid == 0) *z_out = ztot;
    {  // rank block: header (zeros before the block) + 7 bit words, one 32-byte store
      u32 word = Z0 + zbefore;
#pragma unroll
      for (int k = 0; k < WM_ROUNDS; ++k)
        if (lane == k + 1) word = ones[k];
      if (lane < 8) blocks[((size_t)tile * WM_TILE_BLOCKS + w) * 8 + lane] = word;
    }
    const u32 O0 = ztot + (u32)(tbase - Z0);
    const unsigned lt = (1u << lane) - 1u;
    u32 zrun = zbefore;
    u32* const dz = nxt + Z0;
    u32* const dones = nxt + O0;
    const u32 in_tile0 = wbase + lane;
#pragma unroll
    for (int r = 0; r < WM_ROUNDS; ++r) {
      const u32 zm = ~ones[r];
      const u32 zb = zrun + __popc(zm & lt);  // zeros of the tile before this element
      if ((v[r] >> bit) & 1u) dones[in_tile0 + r * 32 - zb] = v[r]; else dz[zb] = v[r];  // (ones before me = index - zeros before me)
      zrun += __popc(zm);
    }
    if (next_bit >= 0) {
      const u32 bndZ = (Z0 / WM_TILE + 1) * WM_TILE;
      const u32 bndO = (O0 / WM_TILE + 1) * WM_TILE;
      u32 wz_nz = 0, wo_nz = 0;
#pragma unroll
      for (int r = 0; r < WM_ROUNDS; ++r) {
        const u32 nzr = __ballot_sync(FULL, ((v[r] >> next_bit) & 1u) == 0u);
        wz_nz += __popc(~ones[r] & nzr);
        wo_nz += __popc(ones[r] & nzr);
      }
      // how many of this warp's zeros (ones) with a 0 next bit land before the destination-tile boundary
      auto before_boundary = [&](bool zero_class, u32 start, u32 count, u32 bnd, u32 total_nz) -> u32 {
        if (start + count <= bnd) return total_nz;
        if (start >= bnd) return 0u;
        const u32 k = bnd - start;  // elements of the class landing before the boundary
        u32 run = 0, before = 0;
#pragma unroll
        for (int r = 0; r < WM_ROUNDS; ++r) {
          const u32 cm = zero_class ? ~ones[r] : ones[r];
          const u32 cr = __popc(cm);
          if (run < k) {  // (warp-uniform) this round still has elements before the boundary
            const u32 nzr = __ballot_sync(FULL, ((v[r] >> next_bit) & 1u) == 0u);
            if (run + cr <= k) {
              before += __popc(cm & nzr);
            } else {
              const bool mine = ((cm >> lane) & 1u) && (u32)__popc(cm & lt) < k - run;
              before += __popc(__ballot_sync(FULL, mine) & nzr);
            }
          }
          run += cr;
        }
        return before;
      };
      const u32 a0 = before_boundary(true, Z0 + zbefore, wz, bndZ, wz_nz);
      const u32 a2 = before_boundary(false, O0 + (wbase - zbefore), wo, bndO, wo_nz);
      if (lane == 0) {
        if (a0) atomicAdd(&s_next[0], a0);
        if (wz_nz - a0) atomicAdd(&s_next[1], wz_nz - a0);
        if (a2) atomicAdd(&s_next[2], a2);
        if (wo_nz - a2) atomicAdd(&s_next[3], wo_nz - a2);
      }
      __syncthreads();
      if (tid < 4 && s_next[tid]) {
        const u32 t = (tid < 2 ? bndZ : bndO) / WM_TILE - 1 + (tid & 1);
        atomicAdd(&next_tile_zeros[t], s_next[tid]);
      }
    }
    return;
  }

  u32 v[WM_ROUNDS], ones[WM_ROUNDS], zmask[WM_ROUNDS];
  u32 wz = 0;
#pragma unroll
  for (int r = 0; r < WM_ROUNDS; ++r) {
    const size_t idx = tbase + wbase + r * 32 + lane;
    const bool ok = idx < n;
    v[r] = ok ? cur[idx] : 0u;
    ones[r] = __ballot_sync(FULL, ok && ((v[r] >> bit) & 1u));
    zmask[r] = __ballot_sync(FULL, ok) & ~ones[r];
    wz += __popc(zmask[r]);
  }
  if (lane == 0) s_wz[w] = wz;
  __syncthreads();
  const u32 zbefore = __reduce_add_sync(FULL, lane < w ? s_wz[lane] : 0u);  // zeros of the tile before this warp
  const u32 tz = 0;
  const u32 Z0 = tile_off[tile];
  const u32 ztot = tile_off[tiles];
  if (tile == 0 && tid == 0) *z_out = ztot;

  // this warp's rank block: header (zeros before the block) + 7 bit words, one 32-byte store
  {
    const u32 blk = tile * WM_TILE_BLOCKS + w;
    u32 word = Z0 + zbefore;
#pragma unroll
    for (int k = 0; k < WM_ROUNDS; ++k)
      if (lane == k + 1) word = ones[k];
    if (lane < 8 && blk < nblk) blocks[(size_t)blk * 8 + lane] = word;
  }

  // stable partition into the next level's order.  Zeros of this tile land in [Z0, Z0 + tz), ones in
  // [O0, O0 + ...): each range straddles at most one boundary between destination tiles, so the
  // next level's per-tile zero counts need only four counters per CTA.
  const u32 O0 = ztot + (u32)(tbase - Z0);
  const u32 bndZ = (Z0 / WM_TILE + 1) * WM_TILE;
  const u32 bndO = (O0 / WM_TILE + 1) * WM_TILE;
  const unsigned lt = (1u << lane) - 1u;
  u32 zrun = zbefore;
  u32 c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma unroll
  for (int r = 0; r < WM_ROUNDS; ++r) {
    const u32 in_tile = wbase + r * 32 + lane;
    const bool ok = tbase + in_tile < n;
    const bool is_zero = (zmask[r] >> lane) & 1u;
    const u32 zb = zrun + __popc(zmask[r] & lt);
    const u32 dst = is_zero ? Z0 + zb : O0 + (in_tile - zb);
    if (ok) {
      nxt[dst] = v[r];
      if (next_bit >= 0 && ((v[r] >> next_bit) & 1u) == 0u) {
        if (is_zero) { if (dst < bndZ) ++c0; else ++c1; }
        else { if (dst < bndO) ++c2; else ++c3; }
      }
    }
    zrun += __popc(zmask[r]);
  }
  if (next_bit >= 0) {
    c0 = __reduce_add_sync(FULL, c0);
    c1 = __reduce_add_sync(FULL, c1);
    c2 = __reduce_add_sync(FULL, c2);
    c3 = __reduce_add_sync(FULL, c3);
    if (lane == 0) {
      if (c0) atomicAdd(&s_next[0], c0);
      if (c1) atomicAdd(&s_next[1], c1);
      if (c2) atomicAdd(&s_next[2], c2);
      if (c3) atomicAdd(&s_next[3], c3);
    }
    __syncthreads();
    if (tid < 4 && s_next[tid]) {
      const u32 t = (tid < 2 ? bndZ : bndO) / WM_TILE - 1 + (tid & 1);
      atomicAdd(&next_tile_zeros[t], s_next[tid]);
    }
  }
  (void)tz;
}

// S0[v] = wm_descend(0, v) for every value of the L-bit range
__global__ void k_wm_s0(DevWM w, u32* __restrict__ S0) {
  const size_t total = (size_t)1 << w.L;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) S0[v] = wm_descend(w, 0u, (u32)v);
}

void WaveletMatrix::build(u32* vals, u32* scratch, size_t n, u64 max_value) {
  CPB_REQUIRE(n < ((size_t)1 << 31), "too many points for the 32-bit device index");
  L = bits_for(max_value);
  npts = (u32)n;
  nblk = (u32)(n / WM_BLOCK + 1);
  const u32 tiles = (nblk + WM_TILE_BLOCKS - 1) / WM_TILE_BLOCKS;
  blocks.alloc((size_t)L * nblk * 8);
  z.alloc(L);
  ProfScope prof("build_dominance", (double)L * ((double)n * 8.0 + (double)nblk * 32.0));
  DBuf<u32> tz[2];
  tz[0].alloc(tiles + 1);
  tz[1].alloc(tiles + 1);
  DBuf<u32> toff(tiles + 1);
  tz[0].zero();
  if (n > 0) CPB_LAUNCH(k_wm_count, tiles, 256, 0, vals, (u32)n, L - 1, tz[0].get());
  u32* cur = vals;
  u32* nxt = scratch;
  for (int l = 0; l < L; ++l) {
    const int bit = L - 1 - l;
    const int next_bit = (l + 1 < L) ? bit - 1 : -1;
    DBuf<u32>& tzc = tz[l & 1];
    DBuf<u32>& tzn = tz[(l + 1) & 1];
    exclusive_scan_u32(tzc.get(), toff.get(), tiles + 1);  // last input entry is 0 -> toff[tiles] = total
    if (next_bit >= 0) tzn.zero();
    {
      ProfScope pk("k_wm_level", (double)n * 8.0 + (double)nblk * 32.0);
      CPB_LAUNCH(k_wm_level, tiles, WM_THREADS, 0, cur, nxt, (u32)n, bit, next_bit, toff.get(), tiles, tzn.get(),
                 blocks.get() + (size_t)l * nblk * 8, nblk, z.get() + l);
    }
    u32* t = cur; cur = nxt; nxt = t;
  }
  CPB_REQUIRE(L <= 30, "value range too wide");
  S0.alloc((size_t)1 << L);
  {
    const size_t total = (size_t)1 << L;
    const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)ctx().sm_count * 16));
    CPB_LAUNCH(k_wm_s0, grid, 256, 0, dev(), S0.get());
  }
}

}  // namespace cpb
