// Microbenchmark: sustained LDS.128 throughput for the relax kernel's access pattern
// (lane = (x&3, y), rows 52 floats apart, planes 22*52 floats apart) vs a plain linear pattern.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int SZD = 52, SYD = 22;
template <int MODE>
__global__ void k(float* out, int iters) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  const int n = 22 * SYD * SZD;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base;
  if (MODE == 0) base = (((warp >> 2) * 4 + (lane >> 3) + 7) * SYD + ((lane & 7) + 7)) * SZD + (warp & 3) * 8;  // kernel pattern
  else base = threadIdx.x * 4;                                                                                    // linear
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      // 16 "columns": different (i,j) offsets, 4 granules each
      const int off = MODE == 0 ? (((c % 5) - 2) * SYD * SZD + ((c / 5) - 1) * SZD) : (c * 1024) % 8192;
      const float* p = sm + base + off + (it & 1) * 4;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"((unsigned)__cvta_generic_to_shared(p + 4 * g)));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  long long t1 = clock64();
  if (acc.x + acc.y + acc.z + acc.w == 1.2345f) out[0] = acc.x;
  if (threadIdx.x == 0) out[1] = (float)(t1 - t0);
}
int main() {
  float* d; cudaMalloc(&d, 64);
  const int smem = 22 * SYD * SZD * 4 + 8192 * 4;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      const int iters = 200;
      if (mode == 0) { k<0><<<1, warps * 32, smem>>>(d, iters); k<0><<<1, warps * 32, smem>>>(d, iters); }
      else { k<1><<<1, warps * 32, smem>>>(d, iters); k<1><<<1, warps * 32, smem>>>(d, iters); }
      cudaDeviceSynchronize();
      float h[2]; cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
      const double lds = (double)iters * 64 * warps;
      printf("%s warps=%2d: %.2f cycles per LDS.128 (warp-wide, 512 B) -> %.1f B/clk/SM (+4 FADD each)\n", mode ? "linear " : "kernel ",
             warps, h[1] / lds, 512.0 * lds / h[1]);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
