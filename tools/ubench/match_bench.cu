// microbenchmark: MATCH.ANY vs 8 ballots, latency (1 warp / SM) and throughput (32 warps / SM)
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned u32;
__device__ __forceinline__ u32 ballot_match8(u32 d) {
  u32 m = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const bool p = (d >> b) & 1u;
    const u32 bal = __ballot_sync(0xffffffffu, p);
    m &= p ? bal : ~bal;
  }
  return m;
}
template <int MODE> __global__ void k(u32* out, int distinct_mask, int iters, long long* cyc) {
  u32 v = (threadIdx.x * 2654435761u >> 7) & distinct_mask;
  u32 acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    u32 m;
    if (MODE == 0) m = __match_any_sync(0xffffffffu, v);
    else m = ballot_match8(v);
    acc += __popc(m);
    v = (v + (m & 1u) + i) & distinct_mask;  // dependent chain
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  u32* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8);
  const int iters = 2048;
  for (int mode = 0; mode < 2; ++mode)
    for (int dm : {0, 3, 15, 255})
      for (int threads : {32, 256, 1024}) {
        if (mode == 0) k<0><<<148, threads>>>(out, dm, iters, cyc); else k<1><<<148, threads>>>(out, dm, iters, cyc);
        cudaDeviceSynchronize();
        if (mode == 0) k<0><<<148, threads>>>(out, dm, iters, cyc); else k<1><<<148, threads>>>(out, dm, iters, cyc);
        cudaDeviceSynchronize();
        printf("%s distinct_mask=%3d warps/SM=%2d cycles/op(warp)=%.1f  cycles/op(SM throughput)=%.2f\n", mode ? "ballot8" : "match  ", dm, threads / 32,
               (double)*cyc / iters, (double)*cyc / iters / (threads / 32));
      }
  return 0;
}
