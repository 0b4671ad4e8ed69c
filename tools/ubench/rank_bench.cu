// microbenchmark of the in-warp ranking loop of the one-sweep pass (no global traffic): variants of the chain
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
#define FULL 0xffffffffu
__device__ __forceinline__ u32 hash(u32 x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// VAR 0: match only; 1: interleaved match + LDS + STS(leader) [kernel]; 2: all matches first, then chain; 3: match + leader ATOMS + SHFL (CUB);
// 4: chain only (no match; mask faked)
template <int VAR, int SKEW> __global__ void __launch_bounds__(256, 3) k(u32* out, int tiles) {
  __shared__ u32 s_wh[8][256];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  u32 acc = 0;
  for (int t = 0; t < tiles; ++t) {
    for (int k = 0; k < 8; ++k) s_wh[k][tid] = 0;
    __syncthreads();
    u32 d[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      u32 h = hash((blockIdx.x * tiles + t) * 4096u + r * 256u + tid);
      d[r] = SKEW ? ((h & 255u) & ((h >> 8) & 255u) & ((h >> 16) & 255u)) : (h & 255u);   // SKEW: each bit set w.p. 1/8
    }
    if (VAR == 0) {
#pragma unroll
      for (int r = 0; r < 16; ++r) acc += __popc(__match_any_sync(FULL, d[r]) & lt);
    } else if (VAR == 1) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const u32 m = __match_any_sync(FULL, d[r]);
        u32 old = s_wh[w][d[r]];
        __syncwarp();
        if ((m & lt) == 0u) s_wh[w][d[r]] = old + __popc(m);
        __syncwarp();
        acc += old + __popc(m & lt);
      }
    } else if (VAR == 2) {
      u32 mk[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) mk[r] = __match_any_sync(FULL, d[r]);
      asm volatile("" ::: "memory");
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const u32 m = mk[r];
        u32 old = s_wh[w][d[r]];
        __syncwarp();
        if ((m & lt) == 0u) s_wh[w][d[r]] = old + __popc(m);
        __syncwarp();
        acc += old + __popc(m & lt);
      }
    } else if (VAR == 3) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const u32 m = __match_any_sync(FULL, d[r]);
        const int leader = __ffs(m) - 1;
        u32 old = 0;
        if (lane == leader) old = atomicAdd(&s_wh[w][d[r]], __popc(m));
        old = __shfl_sync(FULL, old, leader);
        acc += old + __popc(m & lt);
      }
    } else if (VAR == 5) {
      // match emulated through shared memory: every lane ORs its bit into the (warp, digit) word, reads the word back
      __shared__ uint2 s_mc[8][256];  // {lane mask of the round, cursor}
      for (int k = 0; k < 8; ++k) s_mc[k][tid] = make_uint2(0u, 0u);
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        atomicOr(&s_mc[w][d[r]].x, 1u << lane);
        __syncwarp();
        const uint2 mc = s_mc[w][d[r]];
        __syncwarp();
        if ((mc.x & lt) == 0u) s_mc[w][d[r]] = make_uint2(0u, mc.y + __popc(mc.x));
        __syncwarp();
        acc += mc.y + __popc(mc.x & lt);
      }
    } else if (VAR == 4) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const u32 m = 1u << lane;
        u32 old = s_wh[w][d[r]];
        __syncwarp();
        if ((m & lt) == 0u) s_wh[w][d[r]] = old + __popc(m);
        __syncwarp();
        acc += old + __popc(m & lt);
      }
    }
    __syncthreads();
  }
  out[blockIdx.x * 256 + tid] = acc;
}
template <int VAR, int SKEW> void run(u32* out, const char* name) {
  const int tiles = 145;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<VAR, SKEW><<<444, 256>>>(out, tiles); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<VAR, SKEW><<<444, 256>>>(out, tiles); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-40s skew=%d  %.3f ms for %d tiles (one pass at C3 = 64314 tiles)\n", name, SKEW, ms, 444 * tiles);
}
int main() {
  u32* out; cudaMalloc(&out, 444 * 256 * 4);
  run<0, 0>(out, "match only"); run<0, 1>(out, "match only");
  run<1, 0>(out, "match + LDS + STS interleaved"); run<1, 1>(out, "match + LDS + STS interleaved");
  run<2, 0>(out, "matches first, then chain"); run<2, 1>(out, "matches first, then chain");
  run<3, 0>(out, "match + leader ATOMS + SHFL"); run<3, 1>(out, "match + leader ATOMS + SHFL");
  run<5, 0>(out, "smem OR-mask match + cursor"); run<5, 1>(out, "smem OR-mask match + cursor");
  run<4, 0>(out, "chain only"); run<4, 1>(out, "chain only");
  return 0;
}
