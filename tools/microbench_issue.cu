// tools/microbench_issue.cu -- issue-rate microbenchmarks for the DP instruction mix on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_issue microbench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
constexpr int CH = 8;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float sg, float om) {
  float2 S0[CH], S1[CH], S2[CH], acc[CH]; int ia[CH], ib[CH]; for (int i = 0; i < CH; i++) { ia[i] = threadIdx.x + i; ib[i] = i; }
  for (int i = 0; i < CH; i++) { S0[i] = make_float2(threadIdx.x * 0.001f + i, i); S1[i] = make_float2(1.f, 2.f); S2[i] = make_float2(3.f, 4.f); acc[i] = make_float2(0.f, 0.f); }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (MODE == 0) {  // FADD only: 8 scalar adds
        S0[i].x += sg; S0[i].y += sg; S1[i].x += om; S1[i].y += om; S2[i].x += om; S2[i].y += om; acc[i].x += sg; acc[i].y += om;
      } else if (MODE == 1) {  // FADD2 only: 4 packed adds (8 lane-ops)
        S0[i] = add2(S0[i], make_float2(sg, sg)); S1[i] = add2(S1[i], make_float2(om, om)); S2[i] = add2(S2[i], make_float2(om, om)); acc[i] = add2(acc[i], make_float2(sg, om));
      } else if (MODE == 2) {  // scalar cell mix x2: 8 FADD + 2 FMNMX3
        S0[i].x += sg; S0[i].y += sg; S1[i].x += om; S1[i].y += om; S2[i].x += om; S2[i].y += om;
        acc[i].x += max3f(S0[i].x, S1[i].x, S2[i].x); acc[i].y += max3f(S0[i].y, S1[i].y, S2[i].y);
      } else if (MODE == 3) {  // packed cell mix: 4 FADD2 + 2 FMNMX3 for 2 cells
        S0[i] = add2(S0[i], make_float2(sg, sg)); S1[i] = add2(S1[i], make_float2(om, om)); S2[i] = add2(S2[i], make_float2(om, om));
        acc[i] = add2(acc[i], make_float2(max3f(S0[i].x, S1[i].x, S2[i].x), max3f(S0[i].y, S1[i].y, S2[i].y)));
      } else if (MODE == 4) {  // FMNMX3 only
        acc[i].x = max3f(acc[i].x, S0[i].x, S1[i].x); acc[i].y = max3f(acc[i].y, S0[i].y, S2[i].y);
        S0[i].x = max3f(S0[i].x, S1[i].y, acc[i].y); S0[i].y = max3f(S0[i].y, S2[i].x, acc[i].x);
      } else if (MODE == 6) {  // FMNMX (2-input) x4
        acc[i].x = fmaxf(acc[i].x, S0[i].x); acc[i].y = fmaxf(acc[i].y, S0[i].y);
        S1[i].x = fmaxf(S1[i].x, S2[i].y); S1[i].y = fmaxf(S1[i].y, S2[i].x);
      } else if (MODE == 8) {  // 4 FADD2 + 1 FMNMX3
        S0[i] = add2(S0[i], make_float2(sg, sg)); S1[i] = add2(S1[i], make_float2(om, om)); S2[i] = add2(S2[i], make_float2(om, om));
        acc[i] = add2(acc[i], make_float2(max3f(S0[i].x, S1[i].x, S2[i].x), S0[i].y));
      } else if (MODE == 9) {  // 4 FADD2 + 2 integer ALU ops
        S0[i] = add2(S0[i], make_float2(sg, sg)); S1[i] = add2(S1[i], make_float2(om, om)); S2[i] = add2(S2[i], make_float2(om, om));
        acc[i] = add2(acc[i], S0[i]);
        ia[i] = (ia[i] ^ it) + 0x9e37; ib[i] = (ib[i] & 0xfffff) + ia[i];
      } else if (MODE == 10) {  // 2 FADD2 + 2 FMNMX3
        S0[i] = add2(S0[i], make_float2(sg, sg));
        acc[i] = add2(acc[i], make_float2(max3f(S0[i].x, S1[i].x, S2[i].x), max3f(S0[i].y, S1[i].y, S2[i].y)));
      } else if (MODE == 11) {  // 4 FADD2 (one reg-reg) + 2 FMNMX3 on different registers each time (no reuse)
        S0[i] = add2(S0[i], S1[i]); S1[i] = add2(S1[i], make_float2(om, om)); S2[i] = add2(S2[i], make_float2(om, om));
        acc[i] = add2(acc[i], make_float2(max3f(S0[i].x, S1[i].x, S2[i].x), max3f(S0[i].y, S1[i].y, S2[i].y)));
      } else if (MODE == 5) {  // packed adds + 2-input max (FMNMX x4 per 2 cells)
        S0[i] = add2(S0[i], make_float2(sg, sg)); S1[i] = add2(S1[i], make_float2(om, om)); S2[i] = add2(S2[i], make_float2(om, om));
        acc[i] = add2(acc[i], make_float2(fmaxf(fmaxf(S0[i].x, S1[i].x), S2[i].x), fmaxf(fmaxf(S0[i].y, S1[i].y), S2[i].y)));
      }
    }
  }
  float r = 0; for (int i = 0; i < CH; i++) r += acc[i].x + acc[i].y + S0[i].x + S0[i].y + S1[i].x + S1[i].y + (float)(ia[i] + ib[i]); out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, double laneops_per_iter, double cells_per_iter, int ctas_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int ctas = sms * ctas_per_sm, iters = 4096; float* d; cudaMalloc(&d, sizeof(float) * ctas * 256);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); double best = 1e30;
  for (int r = 0; r < 4; r++) { cudaEventRecord(a); k<MODE><<<ctas, 256>>>(d, iters, 0.25f, -2.f); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r) best = ms < best ? ms : best; }
  double thr = (double)ctas * 256 * iters * CH;
  printf("%-28s occ %d CTA/SM  %.3f ms  lane-ops %.2f T/s  cells %.2f T/s  warp-instr/clk/SMSP(@1.965GHz) %.3f\n", name, ctas_per_sm, best, thr * laneops_per_iter / best / 1e9, thr * cells_per_iter / best / 1e9, 0.);
  cudaFree(d);
}
int main() {
  for (int occ : {4}) {
    run<6>("FMNMX x4", 4, 0, occ); run<8>("4FADD2+1FMNMX3", 9, 2, occ); run<9>("4FADD2+2 int ALU", 10, 2, occ); run<10>("2FADD2+2FMNMX3", 6, 2, occ); run<11>("4FADD2(regreg)+2FMNMX3", 10, 2, occ);
    run<0>("FADD x8", 8, 0, occ); run<1>("FADD2 x4 (8 lane-ops)", 8, 0, occ); run<2>("scalar mix 8FADD+2FMNMX3", 10, 2, occ);
    run<3>("packed mix 4FADD2+2FMNMX3", 10, 2, occ); run<4>("FMNMX3 x4", 4, 0, occ); run<5>("packed 4FADD2+4FMNMX", 12, 2, occ);
  }
  return 0;
}
