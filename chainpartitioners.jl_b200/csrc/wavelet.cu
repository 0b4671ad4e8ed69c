// wavelet.cu -- construction of the wavelet-matrix dominance index (kernel "build_dominance").
//
// One launch per bit level: every CTA owns a tile of 7168 consecutive elements (32 rank blocks),
// emits the level's bit-plane + running zero counts (one 32-byte block per 224 elements) and
// stably partitions its elements into the next level's order.  The per-tile zero counts a level
// needs are accumulated by the PREVIOUS level's launch (warp-aggregated atomics on the
// destination tile), so each level reads the keys once and writes them once:
//     algorithmic bytes per level = n * (4 + 4) + n / 7 * (32 / 32) ...  
#include "primitives.cuh"
#include "wavelet.cuh"

namespace cpb {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr int WM_THREADS = 256;
static constexpr int WM_ROUNDS = WM_TILE / WM_THREADS;  // 28 rounds of 32 lanes per warp (4 blocks x 7 words)

// zeros of bit `bit` per tile (first level only)
__global__ void __launch_bounds__(WM_THREADS) k_wm_count(const u32* __restrict__ cur, u32 n, int bit, u32* __restrict__ tile_zeros) {
  __shared__ u32 sm[8];
  const size_t base = (size_t)blockIdx.x * WM_TILE;
  u32 c = 0;
  for (int i = threadIdx.x; i < WM_TILE; i += WM_THREADS) {
    const size_t idx = base + i;
    if (idx < n) c += ((cur[idx] >> bit) & 1u) ^ 1u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(FULL, c, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0;
    for (int k = 0; k < 8; ++k) t += sm[k];
    tile_zeros[blockIdx.x] = t;
  }
}

// tile_off: exclusive scan of per-tile zero counts, tile_off[tiles] = z (total zeros of the level)
__global__ void __launch_bounds__(WM_THREADS) k_wm_level(const u32* __restrict__ cur, u32* __restrict__ nxt, u32 n, int bit, int next_bit,
                                                         const u32* __restrict__ tile_off, u32 tiles, u32* __restrict__ next_tile_zeros,
                                                         u32* __restrict__ blocks, u32 nblk, u32* __restrict__ z_out) {
  __shared__ u32 s_wz[8];
  __shared__ u32 s_words[8][32];  // per warp: 4 blocks x 8 words, written out coalesced
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const u32 tile = blockIdx.x;
  const size_t tbase = (size_t)tile * WM_TILE;
  const u32 wbase = w * (WM_ROUNDS * 32);  // element offset of this warp inside the tile

  u32 v[WM_ROUNDS];
  u32 zmask[WM_ROUNDS];
  u32 wz = 0;
#pragma unroll
  for (int r = 0; r < WM_ROUNDS; ++r) {
    const size_t idx = tbase + wbase + r * 32 + lane;
    const bool ok = idx < n;
    v[r] = ok ? cur[idx] : 0u;
    const u32 b = (v[r] >> bit) & 1u;
    const unsigned ones = __ballot_sync(FULL, ok && b);
    const unsigned valid = __ballot_sync(FULL, ok);
    zmask[r] = valid & ~ones;
    if (lane == 0) s_words[w][(r / 7) * 8 + 1 + (r % 7)] = ones;
    wz += __popc(zmask[r]);
  }
  if (lane == 0) s_wz[w] = wz;
  __syncthreads();
  u32 zbefore = 0;  // zeros of the tile before this warp
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k < w) zbefore += s_wz[k];
  const u32 Z0 = tile_off[tile];
  const u32 ztot = tile_off[tiles];
  if (tile == 0 && tid == 0) *z_out = ztot;

  // rank-block headers: zeros before each of this warp's 4 blocks
  {
    u32 run = Z0 + zbefore;
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      if (lane == 0) s_words[w][bb * 8] = run;
#pragma unroll
      for (int k = 0; k < 7; ++k) run += __popc(zmask[bb * 7 + k]);
    }
  }
  __syncwarp();
  {
    const u32 blk = tile * WM_TILE_BLOCKS + w * 4 + (lane >> 3);
    if (blk < nblk) blocks[(size_t)blk * 8 + (lane & 7)] = s_words[w][lane];
  }

  // stable partition into the next level's order
  const unsigned lt = (1u << lane) - 1u;
  u32 zrun = zbefore;
#pragma unroll
  for (int r = 0; r < WM_ROUNDS; ++r) {
    const u32 in_tile = wbase + r * 32 + lane;
    const size_t idx = tbase + in_tile;
    const bool ok = idx < n;
    const bool is_zero = (zmask[r] >> lane) & 1u;
    const u32 zb = zrun + __popc(zmask[r] & lt);
    u32 dst;
    if (is_zero) dst = Z0 + zb;
    else dst = ztot + (u32)(tbase - Z0) + (in_tile - zb);
    if (ok) nxt[dst] = v[r];
    if (next_bit >= 0) {
      const bool nz = ok && (((v[r] >> next_bit) & 1u) == 0u);
      const u32 key = nz ? dst / WM_TILE : 0xffffffffu;
      const unsigned m = __match_any_sync(FULL, key);
      if (nz && lane == __ffs(m) - 1) atomicAdd(&next_tile_zeros[key], (u32)__popc(m));
    }
    zrun += __popc(zmask[r]);
  }
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
  if (n > 0) CPB_LAUNCH(k_wm_count, tiles, WM_THREADS, 0, vals, (u32)n, L - 1, tz[0].get());
  u32* cur = vals;
  u32* nxt = scratch;
  for (int l = 0; l < L; ++l) {
    const int bit = L - 1 - l;
    const int next_bit = (l + 1 < L) ? bit - 1 : -1;
    DBuf<u32>& tzc = tz[l & 1];
    DBuf<u32>& tzn = tz[(l + 1) & 1];
    exclusive_scan_u32(tzc.get(), toff.get(), tiles + 1);  // last input entry is 0 -> toff[tiles] = total
    if (next_bit >= 0) tzn.zero();
    CPB_LAUNCH(k_wm_level, tiles, WM_THREADS, 0, cur, nxt, (u32)n, bit, next_bit, toff.get(), tiles, tzn.get(),
               blocks.get() + (size_t)l * nblk * 8, nblk, z.get() + l);
    u32* t = cur; cur = nxt; nxt = t;
  }
}

}  // namespace cpb
