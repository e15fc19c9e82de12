// tools/microbench_pairblock.cu -- the k_dp_reg steady-state block (9 species, 2 rows/lane, 2 end codons) as a
// register-only loop: what is the issue/pipe ceiling of this instruction stream at the kernel's occupancy?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2s(float2 a, float b) { return add2(a, make_float2(b, b)); }
__device__ __forceinline__ float max3f(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
constexpr int NK = 9;
template <int MODE>
__global__ void __launch_bounds__(128, 4) k(float* out, int iters, const float* __restrict__ sg, float om) {
  float2 S0[NK], S1[NK], S2[NK];
  for (int k = 0; k < NK; k++) { S0[k] = make_float2(threadIdx.x * 0.01f + k, k); S1[k] = make_float2(1.f + k, 2.f); S2[k] = make_float2(3.f, 4.f + k); }
  float sv[2 * NK];
  for (int k = 0; k < 2 * NK; k++) sv[k] = sg[k];
  float2 tot = make_float2(0.f, 0.f);
  asm volatile("" : "+f"(om));
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    float2 sumA, sumB;
    if (MODE == 0) {  // packed
#pragma unroll
      for (int k = 0; k < NK; k++) {
        S0[k] = add2s(S0[k], sv[k]); S1[k] = add2s(S1[k], om); S2[k] = add2s(S2[k], om);
        float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
        sumA = k == 0 ? m : add2(sumA, m);
      }
#pragma unroll
      for (int k = 0; k < NK; k++) {
        S0[k] = add2s(S0[k], sv[NK + k]); S1[k] = add2s(S1[k], om); S2[k] = add2s(S2[k], om);
        float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
        sumB = k == 0 ? m : add2(sumB, m);
      }
    } else {  // scalar
#pragma unroll
      for (int k = 0; k < NK; k++) {
        S0[k].x += sv[k]; S0[k].y += sv[k]; S1[k].x += om; S1[k].y += om; S2[k].x += om; S2[k].y += om;
        float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
        if (k == 0) sumA = m; else { sumA.x += m.x; sumA.y += m.y; }
      }
#pragma unroll
      for (int k = 0; k < NK; k++) {
        S0[k].x += sv[NK + k]; S0[k].y += sv[NK + k]; S1[k].x += om; S1[k].y += om; S2[k].x += om; S2[k].y += om;
        float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
        if (k == 0) sumB = m; else { sumB.x += m.x; sumB.y += m.y; }
      }
    }
    if (fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 1e30f) { tot = add2(tot, sumA); tot = add2(tot, sumB); }
  }
  float r = tot.x + tot.y; for (int k = 0; k < NK; k++) r += S0[k].x + S0[k].y + S1[k].x + S2[k].y; out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int ctas_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int ctas = sms * ctas_per_sm, iters = 20000; float* d; cudaMalloc(&d, sizeof(float) * ctas * 128);
  float h[18]; for (int i = 0; i < 18; i++) h[i] = -0.3f + 0.05f * i; float* dsg; cudaMalloc(&dsg, sizeof(h)); cudaMemcpy(dsg, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); double best = 1e30;
  for (int r = 0; r < 3; r++) { cudaEventRecord(a); k<MODE><<<ctas, 128>>>(d, iters, dsg, -2.f); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r) best = ms < best ? ms : best; }
  double cells = (double)ctas * 128 * iters * 2 * 2 * NK;  // 2 rows x 2 steps x NK
  double warp_iters_per_smsp = (double)ctas * 4 * iters / (sms * 4);
  printf("%-10s %d CTA/SM (%d warps/SMSP): %.3f ms, %.2f T cells/s, %.1f cycles per pair-block per SMSP-slot (@1.965 GHz)\n", name, ctas_per_sm, ctas_per_sm, best, cells / best / 1e9,
         best * 1e-3 * 1.965e9 / warp_iters_per_smsp);
}
int main() { for (int occ : {1, 2, 4}) { run<0>("packed", occ); run<1>("scalar", occ); } return 0; }
