#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
#define FULL 0xffffffffu
__device__ __forceinline__ u32 hash(u32 x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// GEN 0: arithmetic progression (32 distinct, changes per round); 1: random 8-bit; 2: random 5-bit; 3: random 16-bit
// DEP 0: 16 independent matches per tile; 1: dependent chain
template <int GEN, int DEP> __global__ void __launch_bounds__(256, 3) k(u32* out, int tiles) {
  const int tid = threadIdx.x, lane = tid & 31;
  const unsigned lt = (1u << lane) - 1u;
  u32 acc = 0;
  for (int t = 0; t < tiles; ++t) {
    u32 d[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const u32 h = hash((blockIdx.x * tiles + t) * 4096u + r * 256u + tid);
      d[r] = GEN == 0 ? ((lane * 13 + r + t) & 255u) : GEN == 1 ? (h & 255u) : GEN == 2 ? (h & 31u) : (h & 0xffffu);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const u32 m = __match_any_sync(FULL, DEP ? (d[r] + (acc & 256u)) : d[r]);
      acc += __popc(m & lt);
    }
  }
  out[blockIdx.x * 256 + tid] = acc;
}
template <int GEN, int DEP> void run(u32* out, const char* name, int blocks) {
  const int tiles = 145 * 444 / blocks;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<GEN, DEP><<<blocks, 256>>>(out, tiles); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<GEN, DEP><<<blocks, 256>>>(out, tiles); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double matches_per_sm = (double)blocks * tiles * 8 * 16 / 148;
  printf("%-28s dep=%d blocks=%4d  %.3f ms   cycles per match per SM = %.1f\n", name, DEP, blocks, ms, ms * 1e-3 * 1.96e9 / matches_per_sm);
}
int main() {
  u32* out; cudaMalloc(&out, 1184 * 256 * 4);
  for (int blocks : {148, 444}) {
    run<0, 0>(out, "progression (32 distinct)", blocks); run<0, 1>(out, "progression (32 distinct)", blocks);
    run<1, 0>(out, "random 8-bit", blocks); run<1, 1>(out, "random 8-bit", blocks);
    run<2, 0>(out, "random 5-bit", blocks); run<2, 1>(out, "random 5-bit", blocks);
    run<3, 0>(out, "random 16-bit", blocks); run<3, 1>(out, "random 16-bit", blocks);
  }
  return 0;
}
