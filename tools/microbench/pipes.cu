// Microbenchmark: issue/pipe rates of FADD vs FADD2 (f32x2), FMNMX vs FMNMX3, LDS.128, on one SM
// (1 block) and full chip. Prints cycles per warp-instruction per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define REP 512
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a[8]; u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
  float c = seed * 3.f; u64 pc = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  u64 nz = ((u64)__float_as_uint(-0.0f * seed) << 32) | __float_as_uint(-0.0f * seed);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < REP / 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) a[i] = __fadd_rn(a[i], c);
        if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
        if (MODE == 2) a[i] = fminf(a[i], c + i);
        if (MODE == 3) a[i] = fminf(fminf(a[i], c), a[(i + 1) & 7]);
        if (MODE == 4) a[i] = __fmul_rn(a[i], c);
        if (MODE == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));
        if (MODE == 6) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pc), "l"(nz));
        if (MODE == 7) { a[i] = __fadd_rn(a[i], c); a[i] = fminf(a[i], c + i); }  // fma-pipe + alu-pipe mix
      }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1 + MODE] = (float)(t1 - t0);
}
__global__ void lds(float* out, int iters) {
  __shared__ float4 buf[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float4 acc = make_float4(0, 0, 0, 0);
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 64; ++r) {
      float4 v = buf[(idx + r * 13) & 2047];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  long long t1 = clock64();
  if (acc.x == 1.2345f) out[0] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[20] = (float)(t1 - t0);
}
template <int MODE> void run(const char* name, float* d, int warps, int ops_per_inst) {
  int iters = 64;
  k<MODE><<<1, warps * 32>>>(d, iters, 1.0f);
  k<MODE><<<1, warps * 32>>>(d, iters, 1.0f);
  cudaDeviceSynchronize();
  float h[32]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  double insts_per_smsp = (double)iters * REP * (MODE == 7 ? 2 : 1) * warps / 4.0;
  printf("%-28s warps=%2d cycles=%9.0f  cyc/warp-inst/SMSP=%.3f\n", name, warps, h[1 + MODE], h[1 + MODE] / insts_per_smsp);
}
int main() {
  float* d; cudaMalloc(&d, 256); cudaMemset(d, 0, 256);
  for (int warps : {4, 8, 16}) {
    run<0>("FADD", d, warps, 1); run<1>("FADD2 (add.f32x2)", d, warps, 2); run<4>("FMUL", d, warps, 1);
    run<5>("FMUL2", d, warps, 2); run<6>("FFMA2(-0)", d, warps, 2); run<2>("FMNMX", d, warps, 1);
    run<3>("FMNMX3", d, warps, 1); run<7>("FADD+FMNMX mix", d, warps, 1);
  }
  for (int warps : {4, 8, 16}) {
    lds<<<1, warps * 32>>>(d, 64); lds<<<1, warps * 32>>>(d, 64); cudaDeviceSynchronize();
    float h[32]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("LDS.128+4FADD warps=%2d cycles=%.0f  cyc per LDS.128 per SM = %.3f\n", warps, h[20], h[20] / (64.0 * 64 * warps));
  }
  return 0;
}
