// Microbenchmark: the packed inner loop of relax_column on registers only (no shared-memory
// loads): how many cycles per (offset x 8 nodes) can one SM sustain with 8 / 16 warps?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
constexpr int KZ = 8, WIN = 24;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 r, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2_exact(u64 a, u64 b, u64 nz) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(nz)); return d; }
template <uint32_t KMASK, bool SCALAR>
__device__ __forceinline__ void relax_column(const float (&W)[WIN], const float (&T)[WIN], const float* hdp,
                                             const float (&vn)[KZ], const u64 (&vnE)[4], const u64 (&vnO)[3], u64 nz2, float (&acc)[KZ]) {
  float pend[KZ]; int np = 0, hi = 0;
#pragma unroll
  for (int b = 0; b <= 16; ++b) {
    if (KMASK & (1u << b)) {
      const float hd = hdp[hi++];
      float cand[KZ];
      if (SCALAR) {
#pragma unroll
        for (int k = 0; k < KZ; ++k) cand[k] = __fadd_rn(__fmul_rn(hd, __fadd_rn(vn[k], W[k + b])), T[k + b]);
      } else {
        const u64 hd2 = pack2(hd, hd);
        if ((b & 1) == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const u64 sum = add2(vnE[j], pack2(W[2 * j + b], W[2 * j + b + 1]));
            const u64 c2 = add2(mul2_exact(hd2, sum, nz2), pack2(T[2 * j + b], T[2 * j + b + 1]));
            unpack2(c2, cand[2 * j], cand[2 * j + 1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int k = 2 * j + 1;
            const u64 sum = add2(vnO[j], pack2(W[k + b], W[k + b + 1]));
            const u64 c2 = add2(mul2_exact(hd2, sum, nz2), pack2(T[k + b], T[k + b + 1]));
            unpack2(c2, cand[k], cand[k + 1]);
          }
          cand[0] = __fadd_rn(__fmul_rn(hd, __fadd_rn(vn[0], W[b])), T[b]);
          cand[7] = __fadd_rn(__fmul_rn(hd, __fadd_rn(vn[7], W[7 + b])), T[7 + b]);
        }
      }
      if (np & 1) {
#pragma unroll
        for (int k = 0; k < KZ; ++k) acc[k] = fminf(fminf(acc[k], pend[k]), cand[k]);
      } else {
#pragma unroll
        for (int k = 0; k < KZ; ++k) pend[k] = cand[k];
      }
      ++np;
    }
  }
  if (np & 1) {
#pragma unroll
    for (int k = 0; k < KZ; ++k) acc[k] = fminf(acc[k], pend[k]);
  }
}
template <bool SCALAR>
__global__ void k(const float* __restrict__ in, const float* __restrict__ hd, float* out, int iters, float negz) {
  float W[WIN], T[WIN], vn[KZ], acc[KZ];
  for (int i = 0; i < WIN; ++i) { W[i] = in[threadIdx.x * 64 + i]; T[i] = in[threadIdx.x * 64 + 32 + i]; }
  for (int i = 0; i < KZ; ++i) { vn[i] = in[i + threadIdx.x]; acc[i] = 1e30f; }
  u64 vnE[4], vnO[3];
  for (int j = 0; j < 4; ++j) vnE[j] = pack2(vn[2 * j], vn[2 * j + 1]);
  for (int j = 0; j < 3; ++j) vnO[j] = pack2(vn[2 * j + 1], vn[2 * j + 2]);
  const u64 nz2 = pack2(negz, negz);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    relax_column<0x2aa8u, SCALAR>(W, T, hd + (it & 63) * 8, vn, vnE, vnO, nz2, acc);   // k = -5,-3,-1,1,3,5 (6 odd... bits 3,5,7,9,11,13)
    relax_column<0x1d70u, SCALAR>(W, T, hd + (it & 31) * 8 + 3, vn, vnE, vnO, nz2, acc); // 7 offsets mixed parity
#pragma unroll
    for (int i = 0; i < WIN; ++i) { W[i] += 1e-9f * acc[i & 7]; }  // keep windows loop-variant (24 FADD overhead / 13 offsets)
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < KZ; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[4096] = (float)(t1 - t0);
}
int main() {
  float *in, *hd, *out; cudaMalloc(&in, 64 * 1024 * 4); cudaMalloc(&hd, 4096 * 4); cudaMalloc(&out, 8192 * 4);
  float h[64 * 1024]; for (int i = 0; i < 64 * 1024; ++i) h[i] = 0.1f + (i % 97) * 0.01f;
  cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice); cudaMemcpy(hd, h, 4096 * 4, cudaMemcpyHostToDevice);
  for (int scalar = 0; scalar < 2; ++scalar)
    for (int warps : {4, 8, 16}) {
      const int iters = 2000;
      if (scalar) { k<true><<<1, warps * 32>>>(in, hd, out, iters, -0.0f); k<true><<<1, warps * 32>>>(in, hd, out, iters, -0.0f); }
      else { k<false><<<1, warps * 32>>>(in, hd, out, iters, -0.0f); k<false><<<1, warps * 32>>>(in, hd, out, iters, -0.0f); }
      cudaDeviceSynchronize();
      float c; cudaMemcpy(&c, out + 4096, 4, cudaMemcpyDeviceToHost);
      const double offs = 13.0 * iters;                 // offsets per warp
      printf("%s warps=%2d: %.1f cycles per offset per warp-slot (= x%d warps/SMSP) -> %.2f cycles per offset-warp on an SMSP; ideal FMA 24, issue 16\n",
             scalar ? "scalar" : "packed", warps, c / offs, warps / 4 ? warps / 4 : 1, c / offs / (warps / 4.0));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
