// tools/microbench_latency.cu -- dependent-issue latency of FADD / FADD2 / FMNMX3 on sm_100a (one warp per SM)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm volatile("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float max3f(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float addf(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
template <int MODE, int ILP>
__global__ void k(float* out, long long* cyc, int iters, float x) {
  float2 v[ILP]; for (int i = 0; i < ILP; i++) v[i] = make_float2(threadIdx.x + i, i);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (MODE == 0) v[i].x = addf(v[i].x, x);
        if (MODE == 1) v[i] = add2(v[i], make_float2(x, x));
        if (MODE == 2) v[i].x = max3f(v[i].x, x, v[i].y);
        if (MODE == 3) { v[i] = add2(v[i], make_float2(x, x)); v[i].x = max3f(v[i].x, v[i].y, x); }   // FADD2 -> FMNMX3 -> FADD2 dependent
      }
    }
  }
  long long t1 = clock64();
  float r = 0; for (int i = 0; i < ILP; i++) r += v[i].x + v[i].y; out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int ILP> void run(const char* name, int ops_per_inner) {
  float* d; long long* c; cudaMalloc(&d, 4096); cudaMalloc(&c, 8 * 8); int iters = 2000;
  k<MODE, ILP><<<1, 32>>>(d, c, iters, 0.5f); k<MODE, ILP><<<1, 32>>>(d, c, iters, 0.5f);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s ILP %d: %.2f cycles per dependent step (%.2f cyc/instr)\n", name, ILP, (double)h / (iters * 8.0), (double)h / (iters * 8.0 * ILP * ops_per_inner));
}
int main() {
  run<0, 1>("FADD chain", 1); run<1, 1>("FADD2 chain", 1); run<2, 1>("FMNMX3 chain", 1); run<3, 1>("FADD2->FMNMX3 chain", 2);
  run<0, 4>("FADD x4 indep", 1); run<1, 4>("FADD2 x4 indep", 1); run<1, 8>("FADD2 x8 indep", 1); run<2, 4>("FMNMX3 x4 indep", 1); run<2, 8>("FMNMX3 x8 indep", 1);
  run<1, 2>("FADD2 x2 indep", 1); run<0, 8>("FADD x8 indep", 1);
  return 0;
}
