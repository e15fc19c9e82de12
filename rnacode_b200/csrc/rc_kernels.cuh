// rc_kernels.cuh -- the CUDA kernels of libRNAcode_cuda (sm_100a).
//
//   k_pack   (a) alignment bytes -> class bytes (codes both strands + gap/N/X flags), 128-bit loads/stores
//   k_prep   (a) per block: position->column maps of the reference row for both strands and the
//                sample-invariant frameshift indicator z, packed per (tile, species)
//   k_sigma  (b) codon-pair substitution scores sigma for every (instance, strand, species, position)
//   k_dp     (c) 3-state frameshift DP over all (start, end) pairs + per-row digest of getHSS
//   k_hss    (c) sequential part of getHSS over the row digests; per-sample maxima and native HSS list
//   k_hss_dense  exact fallback of (c) on a materialised S matrix
//
// Reference semantics (all file:line relative to /root/reference): see SURVEY.md Appendix A/C.1.
#pragma once
#include "rc_device.cuh"

namespace rc {

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP), 3-input max (SASS FMNMX3)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void cp_async4(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory loads by 32-bit shared address (the table phase of k_dp_smpf does all its addressing in 32 bits)
__device__ __forceinline__ unsigned lds_u32(unsigned a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f32(unsigned a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(unsigned a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f2(unsigned a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// ---------------------------------------------------------------------------------------------
// (a) k_pack: raw bytes -> class bytes for every instance of every block.
// grid = (x: block, y: grid-stride over 16-byte chunks).  HBM-bound: reads 16 B sample + 16 B native
// (L2 resident), writes 16 B.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const BlockDev* __restrict__ blocks, const unsigned char* __restrict__ raw,
                                              unsigned char* __restrict__ cls, const unsigned char* __restrict__ lut) {
  __shared__ unsigned char s_lut[256];
  s_lut[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  const BlockDev bd = blocks[blockIdx.x];
  const int chunks_per_inst = bd.inst_stride >> 4;
  const long long total = (long long)chunks_per_inst * bd.n_inst;
  const size_t rowbytes = (size_t)bd.N * bd.cols;
  const uint4* nat4 = reinterpret_cast<const uint4*>(raw + bd.nat_off);
  uint4* cls4 = reinterpret_cast<uint4*>(cls + bd.cls_off);
  for (long long ch = (long long)blockIdx.y * blockDim.x + threadIdx.x; ch < total; ch += (long long)gridDim.y * blockDim.x) {
    const int within = (int)(ch % chunks_per_inst), inst = (int)(ch / chunks_per_inst);
    const uint4 g = nat4[within];  // native bytes: where the native alignment has '-', the sample gets '-' (src/misc.c:141-145)
    uint4 v = g;
    if (inst > 0) {
      // samples sit back to back (stride N*cols, as on the host): a 16-byte window is in general unaligned
      const unsigned char* src = raw + bd.raw_off + (size_t)(inst - 1) * rowbytes + (size_t)within * 16;
      const unsigned mis = (unsigned)(reinterpret_cast<size_t>(src) & 15);
      if (mis == 0) {
        v = *reinterpret_cast<const uint4*>(src);
      } else {
        const unsigned* w = reinterpret_cast<const unsigned*>(src - (mis & 3));
        const unsigned sh = (mis & 3) * 8;
        const unsigned w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];  // the buffer ends with 64 bytes of slack
        v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                       __funnelshift_r(w3, w4, sh));
      }
    }
    unsigned vin[4] = {v.x, v.y, v.z, v.w}, gin[4] = {g.x, g.y, g.z, g.w}, out[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
      unsigned o = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        unsigned c = (vin[w] >> (8 * b)) & 0xffu, n = (gin[w] >> (8 * b)) & 0xffu;
        unsigned k = (n == (unsigned)'-') ? CLS_GAP : (unsigned)s_lut[c];
        o |= k << (8 * b);
      }
      out[w] = o;
    }
    cls4[ch] = make_uint4(out[0], out[1], out[2], out[3]);
  }
}

// ---------------------------------------------------------------------------------------------
// (a) k_pack2: the packed form of the alignment rows that kernel (b) inside k_dp_smpf consumes: per (group of 32 instances,
// strand, row) the 2-bit nucleotide codes of the row's characters at the reference's non-gap columns, in the strand's reading
// direction, 16 positions per 32-bit word, lane-interleaved ([word][lane]) -- so that a row of 32 instances is one contiguous
// run for a TMA bulk copy, lane = instance reads its own word without bank conflicts, and a codon (three consecutive
// reference positions) is six adjacent bits: one funnel shift and one mask instead of three byte loads and their shifts.
// calculateSigma only ever looks at those columns (src/score.c:384-391), a '-' or any other symbol there counts as 'A'
// (code 0, ntMap), and the reverse strand complements upper-case ACGTU only (revAln) -- all of which the class byte already
// encodes.  'N' / 'X' (src/score.c:394-404) cannot be expressed in two bits: rows that hold one at a reference position are
// marked per lane in a flag word, their codons take the byte-wise path.
// One CTA per group: the class bytes of as many rows of the group's 32 instances as fit are staged in shared memory with
// coalesced word loads (odd word pitch, many loads in flight), then every thread packs the sixteen positions of one
// (strand, row, word, instance) from its instance's staged bytes -- lane = instance, so the 128-byte lines of the output are
// written whole.  grid = (x: block, y: grid-stride over groups).  cols0 must exist (k_prep<1>).
// 0.5 byte per character and strand.
// ---------------------------------------------------------------------------------------------
constexpr int P2_MAX_COLS = 1020;         // longest row k_pack2 stages (blocks scored by k_dp_smpf are far shorter: their sigma table fits smem)
constexpr int P2_STAGE_BYTES = 64 * 1024;  // staging budget: rows per pass = P2_STAGE_BYTES / (32 * pitch)

__global__ void __launch_bounds__(256) k_pack2(const BlockDev* __restrict__ blocks, const unsigned char* __restrict__ cls,
                                               const int* __restrict__ cols0, unsigned* __restrict__ p2, unsigned* __restrict__ p2f,
                                               int max_L, int stage_bytes) {
  extern __shared__ __align__(16) unsigned char p2_smem[];  // s_c0[2][max_L] ints | s_row[rows per pass][32][pitch]
  const BlockDev bd = blocks[blockIdx.x];
  if (bd.p2_words == 0) return;  // only blocks whose sigma values are formed from packed rows (k_dp_smpf, k_sigma_p2)
  int* s_c0base = reinterpret_cast<int*>(p2_smem);
  unsigned char* s_row = p2_smem + (size_t)2 * max_L * sizeof(int);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = bd.N, L = bd.L, W = bd.p2_words, cols = bd.cols;
  const int groups = (bd.n_inst + 31) >> 5;
  int pitch = (cols + 6 + 3) / 4 * 4;
  if (((pitch / 4) & 1) == 0) pitch += 4;  // odd word count: the 32 instances' bytes of a column sit in 32 banks
  const int rpp = max(1, min(N, stage_bytes / (32 * pitch)));  // rows per pass
  for (int t = threadIdx.x; t < 2 * L; t += blockDim.x) s_c0base[(t / L) * max_L + t % L] = cols0[bd.cols0_off + (size_t)(t / L) * (L + 1) + 1 + t % L];
  for (int g = blockIdx.y; g < groups; g += gridDim.y) {
    for (int r0 = 0; r0 < N; r0 += rpp) {
      const int nr = min(rpp, N - r0);
      __syncthreads();  // s_c0 written; the previous pass is done with s_row
      // stage rows r0 .. r0+nr-1 of the 32 instances: one warp per (row, instance), aligned word copies;
      // column c of (row rr, instance li) at s_row[(rr*32 + li)*pitch + shift(rr) + c]
      for (int pr = warp; pr < nr * 32; pr += 8) {
        const int rr = pr >> 5, li = pr & 31, inst = g * 32 + li;
        const size_t roff = (size_t)(r0 + rr) * cols;
        const unsigned shift = (unsigned)roff & 3u;
        const unsigned* src4 = reinterpret_cast<const unsigned*>(cls + bd.cls_off + (size_t)(inst < bd.n_inst ? inst : 0) * bd.inst_stride + roff - shift);
        unsigned* dst4 = reinterpret_cast<unsigned*>(s_row + (size_t)pr * pitch);
        // asynchronous word copies (LDGSTS): nothing waits between the (row, instance) runs of a warp, all of a pass's loads
        // are in flight together (with ordinary loads every run cost a round trip to memory: 2.5 ms for 10 000 blocks of 10 x 120)
        for (int w = lane; w < (int)(shift + cols + 3) / 4; w += 32) {
          if (inst < bd.n_inst) cp_async4(smem_u32(dst4 + w), src4 + w);
          else dst4[w] = 0u;
        }
      }
      cp_async_wait_all();
      __syncthreads();
      // (strand, row, word) triples, one warp each; lane = instance
      // (the triple of a task is carried along instead of divided out: the divisions were a third of the kernel's instructions)
      int w = warp, rr = 0, s = 0;
      while (w >= W) {
        w -= W;
        if (++rr == nr) { rr = 0; s++; }
      }
      for (int task = warp; task < 2 * nr * W; task += 8, w += 8) {
        while (w >= W) {
          w -= W;
          if (++rr == nr) { rr = 0; s++; }
        }
        const int r = r0 + rr;
        const int sh = s ? 2 : 0;
        const int npos = min(16, L - 16 * w);
        const unsigned rbase = (unsigned)(rr * 32 + lane) * (unsigned)pitch + (unsigned)(((size_t)r * cols) & 3u);  // byte offset of column 0
        const unsigned char* rb = s_row + rbase;
        const int* c0 = s_c0base + s * max_L + 16 * w;
        unsigned word = 0u, flag = 0u;
        const int cfirst = c0[0], clast = c0[npos - 1];
        if (npos == 16 && (clast - cfirst == 15 || cfirst - clast == 15)) {
          // the sixteen positions are sixteen adjacent columns (no gap of the reference inside): four words of class bytes,
          // classified four characters at a time (SIMD in a register), instead of sixteen byte loads.  clo = lowest column.
          const unsigned clo = (unsigned)min(cfirst, clast);
          const unsigned boff = rbase + clo, al = boff & ~3u, fs = (boff & 3u) * 8u;  // warp-uniform alignment
          const unsigned* wp = reinterpret_cast<const unsigned*>(s_row + al);
          const unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = fs ? wp[4] : 0u;
          const unsigned q[4] = {__funnelshift_r(w0, w1, fs), __funnelshift_r(w1, w2, fs), __funnelshift_r(w2, w3, fs),
                                 __funnelshift_r(w3, w4, fs)};
#pragma unroll
          for (int i = 0; i < 4; i++) {
            unsigned x = (q[i] >> sh) & 0x03030303u;  // 2-bit code of each of the four bytes
            x |= x >> 6;                              // bytes 0,1 -> bits 0-3 ; bytes 2,3 -> bits 16-19
            x = (x | (x >> 12)) & 0xffu;              // four codes in eight bits, lowest column first
            word |= x << (8 * i);
            flag |= q[i];
          }
          if (clast < cfirst) {  // reverse strand: position t is column clo + 15 - t -- reverse the order of the 2-bit fields
            word = __brev(word);
            word = ((word & 0x55555555u) << 1) | ((word >> 1) & 0x55555555u);
          }
        } else {
#pragma unroll 4
          for (int t = 0; t < npos; t++) {
            const unsigned b = rb[c0[t]];
            word |= ((b >> sh) & 3u) << (2 * t);
            flag |= b;
          }
        }
        p2[bd.p2_off + ((((size_t)g * 2 + s) * N + r) * W + w) * 32 + lane] = word;
        const unsigned m = __ballot_sync(0xffffffffu, (flag & ((CLS_N | CLS_X) * 0x01010101u)) != 0u);  // any of up to four bytes
        if (lane == 0 && m) atomicOr(&p2f[bd.p2f_off + ((size_t)g * 2 + s) * N + r], m);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (a) k_prep: one CTA per block.
//   cols0[strand][x], x = 1..L : forward 0-based column of the x-th non-gap character of the reference
//   row when read in that strand's direction (pos2col, src/misc.c:250-269, as a prefix sum).
//   z word, layout 0, per (strand, frame, tile, species): bit c = z != 0 at step c of the tile, bit 16+c = z == -1;
//   layout 1, per (strand, frame, tile, step): bit 2k = z != 0 for species k, bit 2k+1 = z == -1
//   (getBlock, src/misc.c:198-244: |gaps_k - gaps_0| mod 3 over the columns of the codon ending at
//   position x plus the reference-gap columns in front of it; from column 1 for x == 3).
// ---------------------------------------------------------------------------------------------
// PHASE 0: both parts in one CTA per block (dense fallback set-up); 1: cols0 only (one CTA per block);
// 2: z words only, the work of a block spread over gridDim.y CTAs (cols0 comes from a PHASE 1 launch)
template <int PHASE>
__global__ void __launch_bounds__(256) k_prep(const BlockDev* __restrict__ blocks, const unsigned char* __restrict__ cls,
                                              int* __restrict__ cols0, unsigned* __restrict__ ztiles) {
  __shared__ int s_cnt[256];
  const BlockDev bd = blocks[blockIdx.x];
  const unsigned char* nat = cls + bd.cls_off;  // instance 0 = native
  const int cols = bd.cols, L = bd.L;
  int* c0f = cols0 + bd.cols0_off;
  int* c0r = c0f + (L + 1);
  if (PHASE != 2) {
  // 1) prefix count of non-gap characters of row 0
  const int per = (cols + 255) / 256;
  const int lo = min(cols, (int)threadIdx.x * per), hi = min(cols, lo + per);
  int cnt = 0;
  for (int c = lo; c < hi; c++) cnt += !(nat[c] & CLS_GAP);
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 256; i++) {
      int t = s_cnt[i];
      s_cnt[i] = run;
      run += t;
    }
  }
  __syncthreads();
  int p = s_cnt[threadIdx.x];
  for (int c = lo; c < hi; c++)
    if (!(nat[c] & CLS_GAP)) {
      ++p;
      c0f[p] = c;
      c0r[L + 1 - p] = c;
    }
  if (threadIdx.x == 0) c0f[0] = c0r[0] = -1;
  __syncthreads();  // cols0 written by this CTA is visible to it below (global writes + barrier)
  }
  if (PHASE == 1) return;
  // 2) z words
  const int NK = bd.NK;
  for (int s = 0; s < 2; s++) {
    const int* c0 = s ? c0r : c0f;
    for (int f = 0; f < 3; f++) {
      const int sites = bd.sites[f], nt = bd.ntiles[f];
      unsigned* zt = ztiles + bd.z_off[s][f];
      const int work = nt * bd.zstride;
      for (int w = blockIdx.y * blockDim.x + threadIdx.x; w < work; w += gridDim.y * blockDim.x) {
        const int tile = w / bd.zstride, u = w % bd.zstride;
        unsigned word = 0;
        // layout 0: u = species, loop over the tile's steps; layout 1: u = step, loop over species;
        // layout 3: u = chunk*TILE + step, loop over the chunk's species
        int n_inner = bd.layout ? NK : TILE, kfirst = 0;
        if (bd.layout == 3) {
          const int gch = u / TILE;
          kfirst = gch * bd.chunk_base + min(gch, bd.chunk_rem);
          n_inner = bd.chunk_base + (gch < bd.chunk_rem ? 1 : 0);
        }
        int j5 = 0;
        if (bd.layout == 5) {  // [chunk][padded step]: w = chunk * (nt * TILE) + step; chunks are counted in quads of species
          const int gch = w / (nt * TILE);
          j5 = w % (nt * TILE);
          kfirst = 4 * (gch * bd.chunk_base + min(gch, bd.chunk_rem));
          n_inner = 4 * (bd.chunk_base + (gch < bd.chunk_rem ? 1 : 0));
        }
        for (int v = 0; v < n_inner; v++) {
          const int k = (bd.layout == 3 || bd.layout == 5) ? kfirst + v : (bd.layout ? v : u);
          const int c = bd.layout == 3 ? u % TILE : (bd.layout ? u : v);
          const int j = bd.layout == 5 ? j5 : tile * TILE + c;
          if (k >= NK || j >= sites) continue;
          const unsigned char* rowk = nat + (size_t)(k + 1) * cols;
          const int x = 3 * j + 3 + f;
          // forward-column range covered by the block; on the reverse strand the range is mirrored,
          // the gap counts are the same
          int a, b;
          if (s == 0) {
            a = (x > 3) ? c0[x - 3] + 1 : 0;
            b = c0[x];
          } else {
            a = c0[x];
            b = (x > 3) ? c0[x - 3] - 1 : cols - 1;
          }
          int gk = 0;
          for (int col = a; col <= b; col++) gk += (rowk[col] & CLS_GAP) ? 1 : 0;
          const int g0 = (b - a + 1) - 3;
          int diff = gk - g0;
          diff = diff < 0 ? -diff : diff;
          const int m = diff % 3;
          if (bd.layout) {
            if (m != 0) word |= (m == 2 ? 3u : 1u) << (2 * v);  // bit 2v: z != 0, bit 2v+1: z == -1 (v = species inside the word)
          } else {
            if (m != 0) word |= 1u << c;
            if (m == 2) word |= 1u << (16 + c);
          }
        }
        zt[w] = word;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (b) k_sigma: sigma[k][x] of calculateSigma (src/score.c:375-426) for every instance / strand.
// grid = (x: item, y: grid-stride over (inst, strand, position)).  Each thread owns one reference
// position and walks the species, so the three column look-ups and the reference codon are shared.
// Tables (BLOSUM as float, genetic code) are staged into shared memory with one TMA bulk copy.
// ---------------------------------------------------------------------------------------------
struct __align__(16) SigmaTables {
  float blosum[576];
  signed char transcode[64];
};

// calculateSigma (src/score.c:406-425) as two look-ups for k_sigma_smp: t[codonA*64 + codonB] = e, where e & 0x3ff indexes
// val[] -- the BLOSUM entry of the two peptides as float, or one of the constants 0 (identical codons), stopPenalty_0,
// stopPenalty_k -- and e >> 10 is the Hamming distance h whose expected score is subtracted (0 for the constants).
constexpr int PT_ZERO = 576, PT_STOP0 = 577, PT_STOPK = 578;
struct __align__(16) PairTables {
  unsigned short t[4096];
  float val[580];
};
// PairTables for codons taken from the packed rows of k_pack2, where the codon's FIRST position sits in the low bits:
// t[qa'*64 + qb'] with q' = c1 | c2 << 2 | c3 << 4 (PairTables: q = c1 << 4 | c2 << 2 | c3).
__host__ __device__ inline unsigned pt_swap(unsigned q) { return ((q & 3u) << 4) | (q & 0xcu) | ((q >> 4) & 3u); }

__global__ void __launch_bounds__(256)
    k_sigma(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const unsigned char* __restrict__ cls,
            const int* __restrict__ cols0, const float* __restrict__ scores, const SigmaTables* __restrict__ tables,
            const unsigned* __restrict__ ztiles, float* __restrict__ sigma, Params prm, int skip3) {
  __shared__ SigmaTables s_tab;
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&s_bar, (unsigned)sizeof(SigmaTables));
    bulk_g2s(&s_tab, tables, (unsigned)sizeof(SigmaTables), &s_bar);
  }
  __syncthreads();
  mbar_wait(&s_bar, 0);

  const Item it = items[blockIdx.x];
  const BlockDev bd = blocks[it.block];
  if (bd.layout == 2 || bd.layout == 1 || (skip3 && bd.layout == 3)) return;  // served by k_sigma_smp / k_sigma_rows / k_sigma_rows3
  const int L = bd.L, N = bd.N, NK = bd.NK, cols = bd.cols;
  const int npos = L - 2;
  const long long total = (long long)it.ninst * 2 * npos;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.y * blockDim.x) {
    int xi, s, inst_l;
    if (bd.layout == 2) {  // sample-major output: consecutive threads = consecutive instances (coalesced stores)
      inst_l = (int)(idx % it.ninst);
      const int rest = (int)(idx / it.ninst);
      s = rest & 1;
      xi = rest >> 1;
    } else {
      xi = (int)(idx % npos);
      const int rest = (int)(idx / npos);
      s = rest & 1;
      inst_l = rest >> 1;
    }
    const int x = xi + 3;
    const int* c0 = cols0 + bd.cols0_off + (size_t)s * (L + 1);
    const int c1 = c0[x - 2], c2 = c0[x - 1], c3 = c0[x];
    const unsigned char* base = cls + bd.cls_off + (size_t)(it.inst0 + inst_l) * bd.inst_stride;
    const unsigned a1 = base[c1], a2 = base[c2], a3 = base[c3];
    const int sh = s ? 2 : 0;
    const unsigned qa = (((a1 >> sh) & 3u) << 4) | (((a2 >> sh) & 3u) << 2) | ((a3 >> sh) & 3u);
    const unsigned nA = (a1 | a2 | a3) & CLS_N;
    const int pepA = s_tab.transcode[qa];
    const int f = xi % 3, j = xi / 3;
    const int tile = j / TILE, c = j % TILE;
    float* out;
    int ks = bd.sig_ks;
    if (bd.layout == 2) {
      // sample-major: [group of 32 instances][step][species quad][lane][4]
      const int rsb = (NK + 3) / 4 * 4;
      out = sigma + it.sigma_off[s][f] + ((size_t)(inst_l >> 5) * bd.sites[f] + j) * rsb * 32 + (inst_l & 31) * 4;
      ks = 0;  // addressed explicitly below
    } else {
      out = sigma + it.sigma_off[s][f] + ((size_t)inst_l * bd.ntiles[f] + tile) * bd.sig_tile + (size_t)c * bd.sig_cs;
    }
    const float* sc = scores + bd.scores_off + (size_t)s * N * 4;
    // layout 1: the step's z word (2 bits per species).  Entries of species with a frameshift at this codon are never
    // read by the recurrence (src/score.c:512-533 ignores sigma); they are stored as +0
    unsigned zword = (bd.layout == 1) ? ztiles[bd.z_off[s][f] + (size_t)tile * bd.zstride + c] : 0u;
    // layout 3 (k_dp_chain): species chunk gch holds csize species from cfirst on; its rows are [chunk][step][rs3]
    const int rs3 = (bd.nkw + 1 + 3) / 4 * 4;
    int gch = 0, cfirst = 0, csize = bd.chunk_base + (bd.chunk_rem > 0 ? 1 : 0);
    float* out3 = nullptr;
    const unsigned* z3 = nullptr;
    if (bd.layout == 3) {
      out3 = sigma + it.sigma_off[s][f] + ((size_t)inst_l * bd.ntiles[f] + tile) * bd.sig_tile + (size_t)c * rs3;
      z3 = ztiles + bd.z_off[s][f] + (size_t)tile * bd.zstride + c;
      zword = z3[0];
    }
    for (int k = 0; k < NK; k++) {
      if (bd.layout == 3 && k == cfirst + csize) {  // close the chunk: dummy species (if any) and the chunk's z word
        float* oc = out3 + (size_t)gch * TILE * rs3;
        for (int q = csize; q < bd.nkw; q++) oc[q] = 0.0f;
        oc[bd.nkw] = __uint_as_float(zword);
        gch++;
        cfirst = k;
        csize = bd.chunk_base + (gch < bd.chunk_rem ? 1 : 0);
        zword = z3[(size_t)gch * TILE];
      }
      const unsigned char* rowk = base + (size_t)(k + 1) * cols;
      const unsigned b1 = rowk[c1], b2 = rowk[c2], b3 = rowk[c3];
      const unsigned qb = (((b1 >> sh) & 3u) << 4) | (((b2 >> sh) & 3u) << 2) | ((b3 >> sh) & 3u);
      float v;
      if (nA | ((b1 | b2 | b3) & CLS_N) | (b1 & b2 & b3 & CLS_X)) {
        v = 0.0f;  // src/score.c:394-404
      } else if (qa == qb) {
        v = 0.0f;  // Hamming distance 0, tested before the stop codons (:409)
      } else {
        const int pepB = s_tab.transcode[qb];
        if (pepA < 0)
          v = prm.stop0;  // :414-416
        else if (pepB < 0)
          v = prm.stopk;  // :418-420
        else {
          const unsigned d = qa ^ qb;
          const int h = ((d & 0x30u) != 0) + ((d & 0x0cu) != 0) + ((d & 0x03u) != 0);
          v = s_tab.blosum[pepA * 24 + pepB] - sc[(k + 1) * 4 + h];  // observed - expected, float32 (:422-425)
        }
      }
      if (bd.layout == 3) {
        if ((zword >> (2 * (k - cfirst))) & 1u) v = 0.0f;
        out3[(size_t)gch * TILE * rs3 + (k - cfirst)] = v;
        continue;
      }
      if (bd.layout == 1 && ((zword >> (2 * k)) & 1u)) v = 0.0f;
      if (bd.layout == 2) out[(k >> 2) * 128 + (k & 3)] = v;
      else out[(size_t)k * ks] = v;
    }
    if (bd.layout == 3) {
      float* oc = out3 + (size_t)gch * TILE * rs3;
      for (int q = csize; q < bd.nkw; q++) oc[q] = 0.0f;
      oc[bd.nkw] = __uint_as_float(zword);
      if (j == bd.sites[f] - 1)  // rows of the last tile past the end of the frame: sigma = 0, no frameshift
        for (int cc = c + 1; cc < TILE; cc++)
          for (int g2 = 0; g2 < bd.nchunk; g2++)
            for (int q = 0; q <= bd.nkw; q++) out3[(size_t)g2 * TILE * rs3 + (size_t)(cc - c) * rs3 + q] = 0.0f;
    }
    if (bd.layout == 1) {  // the step's z word rides in the sigma row, right after the NK sigma values
      out[NK] = __uint_as_float(zword);
      if (j == bd.sites[f] - 1)  // rows of the last tile past the end of the frame: sigma = 0, no frameshift
        for (int cc = c + 1; cc < TILE; cc++)
          for (int q = 0; q <= NK; q++) out[(size_t)(cc - c) * bd.sig_cs + q] = 0.0f;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (b) k_sigma_rows: sigma for layout 1 (k_dp_reg).  One thread per (instance, strand, frame, end codon): it
// produces the whole step row -- NK sigma values, the z word, padding -- and stores it as RS/4 float4, so a warp
// writes 32 consecutive rows (fully coalesced 128-bit stores; k_sigma's position-major mapping scatters 4-byte
// stores over the three frames).  Steps past the end of the frame in the last tile get all-zero rows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k_sigma_rows(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const unsigned char* __restrict__ cls,
                 const int* __restrict__ cols0, const float* __restrict__ scores, const PairTables* __restrict__ tables,
                 const unsigned* __restrict__ ztiles, float* __restrict__ sigma, Params prm) {
  __shared__ PairTables s_tab;
  __shared__ float s_sc[2 * (REG_MAX_NK + 1) * 4];  // expected scores [strand][row][h], h = 0 -> 0 (see PairTables)
  __shared__ __align__(8) uint64_t s_bar;
  const Item it = items[blockIdx.x];
  const BlockDev bd = blocks[it.block];
  if (bd.layout != 1) return;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&s_bar, (unsigned)sizeof(PairTables));
    bulk_g2s(&s_tab, tables, (unsigned)sizeof(PairTables), &s_bar);
  }
  for (int t = threadIdx.x; t < 2 * bd.N * 4; t += blockDim.x) s_sc[t] = (t & 3) ? scores[bd.scores_off + t] : 0.0f;
  __syncthreads();
  mbar_wait(&s_bar, 0);
  const int L = bd.L, N = bd.N, NK = bd.NK, cols = bd.cols;
  const int rs = bd.sig_cs;
  const int st0 = bd.ntiles[0] * TILE, st1 = bd.ntiles[1] * TILE, st2 = bd.ntiles[2] * TILE;  // padded steps per frame
  const int per_is = st0 + st1 + st2;
  const long long total = (long long)it.ninst * 2 * per_is;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.y * blockDim.x) {
    int r = (int)(idx % per_is);
    const int is = (int)(idx / per_is);
    const int s = is & 1, inst_l = is >> 1;
    int f = 0;
    if (r >= st0) { r -= st0; f = 1; }
    if (f == 1 && r >= st1) { r -= st1; f = 2; }
    const int j = r;
    float4* out = reinterpret_cast<float4*>(sigma + it.sigma_off[s][f] + ((size_t)inst_l * bd.ntiles[f]) * bd.sig_tile +
                                            (size_t)j * rs);
    if (j >= bd.sites[f]) {  // padding rows of the last tile
      for (int q = 0; q < rs / 4; q++) out[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      continue;
    }
    const int x = 3 * j + 3 + f;
    const int* c0 = cols0 + bd.cols0_off + (size_t)s * (L + 1);
    const int c1 = c0[x - 2], c2 = c0[x - 1], c3 = c0[x];
    const unsigned char* base = cls + bd.cls_off + (size_t)(it.inst0 + inst_l) * bd.inst_stride;
    const unsigned a1 = base[c1], a2 = base[c2], a3 = base[c3];
    const int sh = s ? 2 : 0;
    const unsigned qa = (((a1 >> sh) & 3u) << 4) | (((a2 >> sh) & 3u) << 2) | ((a3 >> sh) & 3u);
    const unsigned nA = (a1 | a2 | a3) & CLS_N;
    const unsigned short* trow = s_tab.t + (qa << 6);
    const float* sc = s_sc + s * N * 4;
    const unsigned zword = ztiles[bd.z_off[s][f] + j];  // zstride == TILE: one word per step
    for (int q = 0; q < rs / 4; q++) {
      // the twelve bytes of the quad's codons first (independent loads in flight), then the arithmetic
      unsigned bb[4][3];
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int k = 4 * q + t;
        const unsigned char* rowk = base + (size_t)(k < NK ? k + 1 : 0) * cols;
        bb[t][0] = rowk[c1];
        bb[t][1] = rowk[c2];
        bb[t][2] = rowk[c3];
      }
      float v4[4];
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int k = 4 * q + t;
        const unsigned b1 = bb[t][0], b2 = bb[t][1], b3 = bb[t][2];
        const unsigned qb = (((b1 >> sh) & 3u) << 4) | (((b2 >> sh) & 3u) << 2) | ((b3 >> sh) & 3u);
        const unsigned e = trow[qb];
        // src/score.c:394-425 through PairTables, see k_sigma_smp; entries with a frameshift are never read by the recurrence: +0
        float v = s_tab.val[e & 0x3ffu] - sc[(k < NK ? k + 1 : 0) * 4 + (e >> 10)];
        const unsigned zero = nA | ((b1 | b2 | b3) & CLS_N) | (b1 & b2 & b3 & CLS_X) | ((zword >> (2 * (k & 15))) & 1u);
        if (zero || k > NK) v = 0.0f;
        if (k == NK) v = __uint_as_float(zword);
        v4[t] = v;
      }
      out[q] = make_float4(v4[0], v4[1], v4[2], v4[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (b) k_sigma_rows3: the same for layout 3 (k_dp_chain): one thread per (instance, strand, frame, end codon) writes the step's
// rows of ALL species chunks -- per chunk its sigma values, dummy species (+0), the chunk's z word and the padding -- as float4,
// so that a warp (32 consecutive end codons = two tiles) stores 768 contiguous bytes per chunk and tile.  (k_sigma wrote these
// rows value by value with threads running over reference positions, i.e. over the three frames' arrays in turn.)
// Tile layout: [chunk][step][rs3], rs3 = roundup(nkw + 1, 4); rows of the last tile past the frame's end are zero.
// ---------------------------------------------------------------------------------------------
constexpr int SIG3_MAX_N = 501;  // rows of the reference's largest alignment (MAX_NUM_NAMES, src/rnaz_utils.h:7) + 1
__global__ void __launch_bounds__(256)
    k_sigma_rows3(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const unsigned char* __restrict__ cls,
                  const int* __restrict__ cols0, const float* __restrict__ scores, const PairTables* __restrict__ tables,
                  const unsigned* __restrict__ ztiles, float* __restrict__ sigma) {
  __shared__ PairTables s_tab;
  __shared__ float s_sc[2 * SIG3_MAX_N * 4];  // expected scores [strand][row][h], h = 0 -> 0 (see PairTables)
  __shared__ __align__(8) uint64_t s_bar;
  const Item it = items[blockIdx.x];
  const BlockDev bd = blocks[it.block];
  if (bd.layout != 3) return;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&s_bar, (unsigned)sizeof(PairTables));
    bulk_g2s(&s_tab, tables, (unsigned)sizeof(PairTables), &s_bar);
  }
  for (int t = threadIdx.x; t < 2 * bd.N * 4; t += blockDim.x) s_sc[t] = (t & 3) ? scores[bd.scores_off + t] : 0.0f;
  __syncthreads();
  mbar_wait(&s_bar, 0);
  const int L = bd.L, N = bd.N, cols = bd.cols;
  const int nkw = bd.nkw, rs3 = (nkw + 1 + 3) / 4 * 4;
  const int st0 = bd.ntiles[0] * TILE, st1 = bd.ntiles[1] * TILE, st2 = bd.ntiles[2] * TILE;  // padded steps per frame
  const int per_is = st0 + st1 + st2;
  const long long total = (long long)it.ninst * 2 * per_is;
  for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.y * blockDim.x) {
    int r = (int)(idx % per_is);
    const int is = (int)(idx / per_is);
    const int s = is & 1, inst_l = is >> 1;
    int f = 0;
    if (r >= st0) { r -= st0; f = 1; }
    if (f == 1 && r >= st1) { r -= st1; f = 2; }
    const int j = r, tile = j / TILE, c = j % TILE;
    float* out0 = sigma + it.sigma_off[s][f] + ((size_t)inst_l * bd.ntiles[f] + tile) * bd.sig_tile + (size_t)c * rs3;
    if (j >= bd.sites[f]) {  // padding rows of the last tile: sigma = 0, no frameshift
      for (int g = 0; g < bd.nchunk; g++)
        for (int q = 0; q < rs3 / 4; q++)
          reinterpret_cast<float4*>(out0 + (size_t)g * TILE * rs3)[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      continue;
    }
    const int x = 3 * j + 3 + f;
    const int* c0 = cols0 + bd.cols0_off + (size_t)s * (L + 1);
    const int c1 = c0[x - 2], c2 = c0[x - 1], c3 = c0[x];
    const unsigned char* base = cls + bd.cls_off + (size_t)(it.inst0 + inst_l) * bd.inst_stride;
    const unsigned a1 = base[c1], a2 = base[c2], a3 = base[c3];
    const int sh = s ? 2 : 0;
    const unsigned qa = (((a1 >> sh) & 3u) << 4) | (((a2 >> sh) & 3u) << 2) | ((a3 >> sh) & 3u);
    const unsigned nA = (a1 | a2 | a3) & CLS_N;
    const unsigned short* trow = s_tab.t + (qa << 6);
    const float* sc = s_sc + s * N * 4;
    const unsigned* z3 = ztiles + bd.z_off[s][f] + (size_t)tile * bd.zstride + c;  // + chunk * TILE
    int cfirst = 0;
    for (int g = 0; g < bd.nchunk; g++) {
      const int csize = bd.chunk_base + (g < bd.chunk_rem ? 1 : 0);
      const unsigned zword = z3[(size_t)g * TILE];
      float4* out = reinterpret_cast<float4*>(out0 + (size_t)g * TILE * rs3);
      for (int q = 0; q < rs3 / 4; q++) {
        unsigned bb[4][3];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int kk = 4 * q + t;
          const unsigned char* rowk = base + (size_t)(kk < csize ? cfirst + kk + 1 : 0) * cols;
          bb[t][0] = rowk[c1];
          bb[t][1] = rowk[c2];
          bb[t][2] = rowk[c3];
        }
        float v4[4];
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int kk = 4 * q + t;
          const unsigned b1 = bb[t][0], b2 = bb[t][1], b3 = bb[t][2];
          const unsigned qb = (((b1 >> sh) & 3u) << 4) | (((b2 >> sh) & 3u) << 2) | ((b3 >> sh) & 3u);
          const unsigned e = trow[qb];
          // src/score.c:394-425 through PairTables; entries of a species with a frameshift are never read by the recurrence: +0
          float v = s_tab.val[e & 0x3ffu] - sc[(kk < csize ? cfirst + kk + 1 : 0) * 4 + (e >> 10)];
          const unsigned zero = nA | ((b1 | b2 | b3) & CLS_N) | (b1 & b2 & b3 & CLS_X) | ((zword >> (2 * (kk & 15))) & 1u);
          if (zero || kk >= csize) v = 0.0f;            // dummy species and padding: +0
          if (kk == nkw) v = __uint_as_float(zword);    // the chunk's z word rides behind its nkw species slots
          v4[t] = v;
        }
        out[q] = make_float4(v4[0], v4[1], v4[2], v4[3]);
      }
      cfirst += csize;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (b) k_sigma_smp: sigma for the sample-major layouts (k_dp_smp / k_dp_smps).  One CTA per (item, group of 32
// instances, strand, chunk of SIG_PCH reference positions, share of the species quads).  sigma only ever reads the
// alignment at the reference's non-gap columns, so the CTA stages COMPACTED class bytes -- for its position chunk,
// column t of the staged row is the alignment column of reference position x_lo + t (cols0, gathered once into
// shared memory) -- of the group's reference rows and of four species rows at a time: the codon ending at
// position x then sits at three consecutive staged columns, the staging buffer has a fixed size whatever the
// block length, and long blocks give many CTAs.  Padded row pitch => conflict-free byte reads with lane = instance;
// each thread produces the four sigma values of one (position, instance) and stores them as one float4, so a warp
// writes 512 contiguous bytes of the [step][species quad][lane][4] table.
// ---------------------------------------------------------------------------------------------
constexpr int SIG_PCH = 216;              // reference positions per CTA; the window of raw columns (SIG_PCH + 2 + reference gaps) should fit SIG_PITCH - 4
constexpr int SIG_PITCH = 260;            // bytes per staged row: 65 words, odd => 32 lanes hit 32 banks

__global__ void __launch_bounds__(256)
    k_sigma_smp(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const unsigned char* __restrict__ cls,
                const int* __restrict__ cols0, const float* __restrict__ scores, const PairTables* __restrict__ tables,
                float* __restrict__ sigma, int nqz) {
  __shared__ PairTables s_tab;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_col[2][SIG_PITCH];
  __shared__ float s_sc[2][4][4];  // expected scores of the quad's species: [strand][species][h], h = 0 -> 0
  __shared__ long long s_base[2][3];  // float offset of the quad's output per (strand, frame): indexing Item / BlockDev arrays
                                      // by a run-time frame inside the loop would put them into local memory
  __shared__ __align__(16) unsigned char s_ref[32 * SIG_PITCH];
  extern __shared__ __align__(16) unsigned char s_sp[];  // [4 * 32 * SIG_PITCH] (dynamic: static shared memory ends at 48 KB)
  const Item it = items[blockIdx.x];
  const BlockDev bd = blocks[it.block];
  const int group = blockIdx.y;
  if ((bd.layout != 2 && bd.layout != 5) || bd.smp_fused || bd.sig_p2 || group * 32 >= it.ninst) return;  // fused: k_dp_smpf builds its own table; sig_p2: k_sigma_p2
  // blockIdx.z = ((position chunk * 2) + strand) * nqz + quad share
  const int qz = blockIdx.z % nqz, s_cta = (blockIdx.z / nqz) & 1, pc = blockIdx.z / (2 * nqz);
  const int L = bd.L, N = bd.N, NK = bd.NK, cols = bd.cols;
  const int npos = L - 2;
  // Short blocks (whole rows fit the staging buffer): the rows are staged once, uncompacted, and the CTA of strand 0
  // serves both strands from them.  Longer blocks: compacted staging, one CTA per strand and position chunk.
  const bool small = cols <= SIG_PITCH - 4;  // room for the word-copy shift below
  const int xi_lo = small ? 0 : pc * SIG_PCH;
  if (qz * 4 >= NK || xi_lo >= npos || (small && (s_cta == 1 || pc > 0))) return;  // nothing to do for this CTA
  const int xi_n = small ? npos : min(SIG_PCH, npos - xi_lo);  // positions of this CTA
  const int n_staged = small ? cols : xi_n + 2;                // staged columns per row
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&s_bar, (unsigned)sizeof(PairTables));
    bulk_g2s(&s_tab, tables, (unsigned)sizeof(PairTables), &s_bar);
  }
  // Window of alignment columns that holds the CTA's positions.  If it fits the staging buffer (always for short blocks;
  // for a position chunk of a longer block unless the reference row has > ~250 gap columns inside it) the rows' bytes
  // wlo..whi are staged as they are, with aligned 32-bit copies, and a codon is read at its three columns; otherwise the
  // columns are gathered byte by byte into a compacted row (codon = three consecutive staged bytes).
  int wlo = 0, whi = cols - 1;
  if (!small) {
    const int ca = cols0[bd.cols0_off + (size_t)s_cta * (L + 1) + xi_lo + 1];
    const int cb = cols0[bd.cols0_off + (size_t)s_cta * (L + 1) + xi_lo + xi_n + 2];
    wlo = min(ca, cb);
    whi = max(ca, cb);
  }
  const int wn = whi - wlo + 1;
  const bool windowed = wn <= SIG_PITCH - 4;  // room for the word-copy shift
  // s_col[s][t]: (windowed: window-relative) alignment column of reference position xi_lo + 1 + t on strand s
  // (positions x-2 .. x of xi are xi+1 .. xi+3)
  for (int t = threadIdx.x; t < xi_n + 2; t += blockDim.x) {
    if (small) {
      s_col[0][t] = cols0[bd.cols0_off + 1 + t];
      s_col[1][t] = cols0[bd.cols0_off + (L + 1) + 1 + t];
    } else {
      s_col[s_cta][t] = cols0[bd.cols0_off + (size_t)s_cta * (L + 1) + xi_lo + 1 + t] - (windowed ? wlo : 0);
    }
  }
  __syncthreads();
  const int ninst_g = min(32, it.ninst - group * 32);
  const unsigned char* gbase = cls + bd.cls_off + (size_t)(it.inst0 + group * 32) * bd.inst_stride;
  const int wid = threadIdx.x >> 5, ln = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // reference rows of the 32 instances: one warp per row, lanes along the staged columns.  Windowed: instances start
  // 16-byte aligned, rows and windows at any byte, so a staged row is shifted by its source's misalignment and window
  // column t sits at dst[shift + t].
  const unsigned shift_ref = (unsigned)wlo & 3u;
  for (int li = wid; li < 32; li += nwarps) {
    const unsigned char* src = gbase + (size_t)li * bd.inst_stride;
    unsigned char* dst = s_ref + li * SIG_PITCH;
    if (windowed) {
      const unsigned* src4 = reinterpret_cast<const unsigned*>(src + wlo - shift_ref);
      for (int w = ln; w < (int)(shift_ref + wn + 3) / 4; w += 32) reinterpret_cast<unsigned*>(dst)[w] = (li < ninst_g) ? src4[w] : 0u;
    } else {
      for (int t = ln; t < n_staged; t += 32) dst[t] = (li < ninst_g) ? src[s_col[s_cta][t]] : (unsigned char)0;
    }
  }
  __syncthreads();
  mbar_wait(&s_bar, 0);
  const int rsb = (NK + 3) / 4 * 4;
  for (int kq = qz; kq < rsb / 4; kq += nqz) {  // the quads of a wide alignment are spread over nqz CTAs
    // stage species rows 1+4kq .. 4+4kq
    for (int pr = wid; pr < 4 * 32; pr += nwarps) {  // (species of the quad, instance) pairs: one warp per row
      const int kk = pr >> 5, li = pr & 31;
      const int row = 1 + 4 * kq + kk;
      const bool ok = li < ninst_g && row < N;
      const unsigned char* src = gbase + (size_t)li * bd.inst_stride + (size_t)row * cols;
      unsigned char* dst = s_sp + (kk * 32 + li) * SIG_PITCH;
      if (windowed) {
        const unsigned shift = (unsigned)((size_t)row * cols + wlo) & 3u;  // = (src + wlo) & 3
        const unsigned* src4 = reinterpret_cast<const unsigned*>(src + wlo - shift);
        for (int w = ln; w < (int)(shift + wn + 3) / 4; w += 32) reinterpret_cast<unsigned*>(dst)[w] = ok ? src4[w] : 0u;
      } else {
        for (int t = ln; t < n_staged; t += 32) dst[t] = ok ? src[s_col[s_cta][t]] : (unsigned char)0;
      }
    }
    // output address of quad kq: layout 2 is [group][step][quad][lane][4]; layout 5 is [chunk][group][step][3 quads][lane][4]
    // with quad kq = quad qq of chunk ch
    size_t rowmul, step_stride, qoff;  // out = base + rowmul * sites[f] + j * step_stride + qoff + lane * 4
    if (bd.layout == 5) {
      const int big = bd.chunk_rem * (bd.chunk_base + 1);
      const int ch = kq < big ? kq / (bd.chunk_base + 1) : bd.chunk_rem + (kq - big) / bd.chunk_base;
      const int qq = kq < big ? kq % (bd.chunk_base + 1) : (kq - big) % bd.chunk_base;
      const int ngrp = (it.ninst + 31) / 32;
      rowmul = ((size_t)ch * ngrp + group) * 3 * 128;
      step_stride = 3 * 128;
      qoff = (size_t)qq * 128;
    } else {
      rowmul = (size_t)group * (rsb / 4) * 128;
      step_stride = (size_t)(rsb / 4) * 128;
      qoff = (size_t)kq * 128;
    }
    if (threadIdx.x < 32) {  // [strand][species of the quad][h]
      const int ss = threadIdx.x >> 4, kk = (threadIdx.x >> 2) & 3, h = threadIdx.x & 3, row = 1 + 4 * kq + kk;
      s_sc[ss][kk][h] = (h > 0 && row < N) ? scores[bd.scores_off + ((size_t)ss * N + row) * 4 + h] : 0.0f;
    } else if (threadIdx.x < 38) {
      const int ss = (threadIdx.x - 32) / 3, f = (threadIdx.x - 32) % 3;
      s_base[ss][f] = (long long)(it.sigma_off[ss][f] + rowmul * bd.sites[f] + qoff);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;  // = e & 31 below: the CTA has whole warps
    const int n_units = (small ? 2 : 1) * xi_n;
    for (int u = threadIdx.x >> 5; u < n_units; u += blockDim.x >> 5) {
      int s, tl, i1, i2, i3;  // strand, position inside the CTA's range, staged columns of the codon
      if (small) {
        s = u & 1;
        tl = u >> 1;
      } else {
        s = s_cta;
        tl = u;
      }
      if (windowed) {
        i1 = s_col[s][tl];
        i2 = s_col[s][tl + 1];
        i3 = s_col[s][tl + 2];
      } else {
        i1 = tl;
        i2 = tl + 1;
        i3 = tl + 2;
      }
      const int xi = xi_lo + tl;
      const int sh = s ? 2 : 0;
      const float* sc = &s_sc[s][0][0];
      const unsigned char* rr = s_ref + lane * SIG_PITCH + (windowed ? shift_ref : 0u);
      const unsigned a1 = rr[i1], a2 = rr[i2], a3 = rr[i3];
      const unsigned qa = (((a1 >> sh) & 3u) << 4) | (((a2 >> sh) & 3u) << 2) | ((a3 >> sh) & 3u);
      const unsigned nA = (a1 | a2 | a3) & CLS_N;
      const unsigned short* trow = s_tab.t + (qa << 6);
      float v4[4];
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const unsigned char* rk = s_sp + (kk * 32 + lane) * SIG_PITCH + (windowed ? ((unsigned)((size_t)(1 + 4 * kq + kk) * cols + wlo) & 3u) : 0u);
        const unsigned b1 = rk[i1], b2 = rk[i2], b3 = rk[i3];
        const unsigned qb = (((b1 >> sh) & 3u) << 4) | (((b2 >> sh) & 3u) << 2) | ((b3 >> sh) & 3u);
        const unsigned e = trow[qb];
        const float v = s_tab.val[e & 0x3ffu] - sc[kk * 4 + (e >> 10)];  // observed - expected (:422-425), or constant - 0
        const unsigned zero = nA | ((b1 | b2 | b3) & CLS_N) | (b1 & b2 & b3 & CLS_X);  // src/score.c:394-404
        v4[kk] = (zero || 4 * kq + kk >= NK) ? 0.0f : v;
      }
      const int j = xi / 3, f = xi - 3 * j;
      float* out = sigma + s_base[s][f] + (size_t)j * step_stride + lane * 4;
      *reinterpret_cast<float4*>(out) = make_float4(v4[0], v4[1], v4[2], v4[3]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// (c) k_dp
// One warp = one task = (instance, strand, frame, group of 32*R consecutive start codons); lane l owns
// the R rows  row_base + l*R + t.  All lanes walk the end codon j together, so sigma and z are
// warp-uniform: they are staged per tile of 16 steps into the warp's shared-memory ring by TMA bulk
// copies (2 stages, mbarrier complete_tx) and read back with broadcast LDS.128.  The 3-float state per
// (species, row) lives in shared memory between tiles ([k][state][t][lane], conflict-free); inside a
// tile it is in registers.  Float operations and their order are exactly the reference's
// (src/score.c:506-533, :834-843): per species the recurrence advances left to right, the species sum
// is accumulated in k order, then max(sum, Delta) / (N-1).
//
// getHSS digest (src/score.c:888-961).  While a row produces its entries S[b][i] in i order, the lane
// folds them from a fresh state with the tie rule's length test dropped ("fresh lenient fold":
// accept e iff e > last accepted or |e - last accepted| < 1e-4).  Whatever state the sequential scan
// enters the row with, (1) every entry it can accept is accepted by the fresh fold, (2) after its
// first acceptance it coincides with the fresh fold, so the row ends in the fresh fold's final
// (value vF, end jF) or is left untouched, and (3) whether it accepts anything depends only on the row
// maximum Emax and on the accepted entries within 1e-4 of Emax (the tie band).  RowRec keeps exactly
// that; k_hss replays it.  Band overflow (> band_slots distinct near-ties) marks the row and the
// alignment is re-scored through the dense path.
// ---------------------------------------------------------------------------------------------
struct RowSt {
  float lb;  // last accepted value of the fresh fold (-inf before the first)
  float M;   // row maximum so far (-inf before the first)
  int jF;    // end codon of the last accepted entry
  int nb;    // band entries in use | 0x8000 overflow
};

// fabs(e - cur) < 0.0001 is evaluated in double by the reference (src/score.c:953-954).  For a float
// difference d this is |d| <= 0.0001f: (double)0.0001f < 0.0001 < (double)nextafterf(0.0001f, 1).
__device__ __noinline__ RowSt hss_accept(RowSt s, float e, int j, RowRec* rec, int slots) {
  const float d = e - s.lb;
  if (!(d >= -0.0001f)) return s;  // e > lb  or  |e - lb| within tolerance
  s.lb = e;
  s.jF = j;
  int nb = s.nb & 0xff, ovf = s.nb & 0x8000;
  if (e > s.M) {
    if (s.M - e < -0.0001f) nb = 0;  // everything seen so far is out of the band of the new maximum
    s.M = e;
  }
  if (nb > 0 && rec->be[nb - 1] == e) {
    rec->bj[nb - 1] = (unsigned short)j;  // exact tie: only the longest one can matter
  } else {
    if (nb == slots) {  // drop entries that fell out of the band
      int w = 0;
      for (int m = 0; m < nb; m++) {
        const float b = rec->be[m];
        if (!(b - s.M < -0.0001f)) {
          rec->be[w] = b;
          rec->bj[w] = rec->bj[m];
          w++;
        }
      }
      nb = w;
    }
    if (nb == slots) {
      ovf = 0x8000;
    } else {
      rec->be[nb] = e;
      rec->bj[nb] = (unsigned short)j;
      nb++;
    }
  }
  s.nb = nb | ovf;
  return s;
}

template <int R>
struct DpSmem {
  static __host__ __device__ size_t state_bytes(int NK) { return (size_t)NK * 3 * R * 32 * sizeof(float); }
  static __host__ __device__ size_t sig_bytes(int NK) { return (size_t)NK * TILE * sizeof(float); }
  static __host__ __device__ size_t z_bytes(int zstride) { return (size_t)zstride * sizeof(unsigned); }
  static __host__ __device__ size_t per_warp(int NK, int zstride, int nstages = 2) {
    return state_bytes(NK) + nstages * (sig_bytes(NK) + z_bytes(zstride)) + 16;
  }
};

template <int R, bool DENSE>
__global__ void __launch_bounds__(DP_WARPS * 32)
    k_dp(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const CtaDesc* __restrict__ ctas,
         const float* __restrict__ sigma, const unsigned* __restrict__ ztiles, RowRec* __restrict__ recs,
         float* __restrict__ dense, Params prm, int band_slots, int smem_NK, int smem_zstride, int nst) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CtaDesc cd = ctas[blockIdx.x];
  const Item& it = items[cd.item];
  const BlockDev& bd = blocks[it.block];
  const int strand = cd.sf / 3, frame = cd.sf % 3;
  const int sites = bd.sites[frame], ntiles = bd.ntiles[frame], NK = bd.NK, zstride = bd.zstride;
  const int ngroups = (sites + 32 * R - 1) / (32 * R);
  // A CTA descriptor covers DP_WARPS tasks.  Very wide alignments need so much shared-memory state per task that the
  // kernel is launched with fewer warps; each warp then takes several of the CTA's tasks, one after the other.
  const int nwarps = blockDim.x >> 5;
  // nst: stages of the sigma / z ring (2; 1 when the state of a 500-row alignment leaves no room for a second one)
  unsigned char* wsm = smem + (size_t)warp * DpSmem<R>::per_warp(smem_NK, smem_zstride, nst);
  float* st = reinterpret_cast<float*>(wsm);
  unsigned char* ring = wsm + DpSmem<R>::state_bytes(smem_NK);
  const size_t stage_bytes = DpSmem<R>::sig_bytes(smem_NK) + DpSmem<R>::z_bytes(smem_zstride);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + nst * stage_bytes);
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  unsigned ring_it = 0;  // tiles this warp has consumed so far, over all its tasks (stage = ring_it % nst, phase = ring_it / nst)
#pragma unroll 1
  for (int tsk = warp; tsk < DP_WARPS; tsk += nwarps) {
  const int task = cd.task0 + tsk;
  if (task >= it.ninst * ngroups) return;  // warps are independent: no CTA-wide barrier below
  const int inst_l = task / ngroups, g = task % ngroups;
  const int row_base = g * 32 * R;
  const int r0 = row_base + lane * R;  // first of this lane's R rows

  const float* sig_src = sigma + it.sigma_off[strand][frame] + (size_t)inst_l * ntiles * NK * TILE;
  const unsigned* z_src = ztiles + bd.z_off[strand][frame];
  const unsigned sig_tx = (unsigned)(NK * TILE * sizeof(float)), z_tx = (unsigned)(zstride * sizeof(unsigned));

  const int t0 = row_base / TILE;
  const int t_last_diag = (row_base + 32 * R - 1) / TILE;
  if (lane == 0) {
    for (int q = 0; q < nst && t0 + q < ntiles; q++) {
      const unsigned sq = (ring_it + q) % nst;
      unsigned char* dst = ring + sq * stage_bytes;
      mbar_expect_tx(&bars[sq], sig_tx + z_tx);
      bulk_g2s(dst, sig_src + (size_t)(t0 + q) * NK * TILE, sig_tx, &bars[sq]);
      bulk_g2s(dst + DpSmem<R>::sig_bytes(smem_NK), z_src + (size_t)(t0 + q) * zstride, z_tx, &bars[sq]);
    }
  }
  for (int i = 0; i < NK * 3 * R; i++) st[i * 32 + lane] = 0.0f;
  __syncwarp();

  RowSt rs[R];
  RowRec* rec0 = recs + it.rec_off[strand][frame] + (size_t)inst_l * sites + r0;
#pragma unroll
  for (int t = 0; t < R; t++) {
    rs[t].lb = -INFINITY;
    rs[t].M = -INFINITY;
    rs[t].jF = 0;
    rs[t].nb = 0;
  }
  float* dense_row[R];
  if (DENSE) {
#pragma unroll
    for (int t = 0; t < R; t++) {
      const long long r = r0 + t;
      // row r of the frame starts at offset r*sites - r(r-1)/2 and holds entries j = r..sites-1
      dense_row[t] = dense + it.dense_off[strand][frame] + (size_t)inst_l * ((size_t)sites * (sites + 1) / 2) +
                     (r * sites - r * (r - 1) / 2) - r;
    }
  }

  const float Delta = prm.Delta, Omega = prm.Omega, omega = prm.omega;
  const float fNK = bd.fNK, rcpNK = bd.rcpNK;

  for (int tile = t0; tile < ntiles; tile++, ring_it++) {
    const int s = (int)(ring_it % nst);
    const unsigned parity = (ring_it / nst) & 1u;
    const float* sg = reinterpret_cast<const float*>(ring + s * stage_bytes);
    const unsigned* zt = reinterpret_cast<const unsigned*>(ring + s * stage_bytes + DpSmem<R>::sig_bytes(smem_NK));
    const int j0 = tile * TILE;
    const bool diag = tile <= t_last_diag;
    mbar_wait(&bars[s], parity);

    float sum[TILE][R];
#pragma unroll
    for (int c = 0; c < TILE; c++)
#pragma unroll
      for (int t = 0; t < R; t++) sum[c][t] = 0.0f;

    for (int k = 0; k < NK; k++) {
      const unsigned zz = zt[k];
      const float4* sp = reinterpret_cast<const float4*>(sg + k * TILE);
      const float4 q0 = sp[0], q1 = sp[1], q2 = sp[2], q3 = sp[3];
      const float sv[TILE] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w,
                              q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
      float S0[R], S1[R], S2[R];
      float* stk = st + (size_t)k * 3 * R * 32 + lane;
#pragma unroll
      for (int t = 0; t < R; t++) {
        S0[t] = stk[(0 * R + t) * 32];
        S1[t] = stk[(1 * R + t) * 32];
        S2[t] = stk[(2 * R + t) * 32];
      }
      if (zz == 0u && !diag) {
        // no frameshift for this species anywhere in the tile and every row already started (src/score.c:506-510)
#pragma unroll
        for (int c = 0; c < TILE; c++) {
#pragma unroll
          for (int t = 0; t < R; t++) {
            S0[t] += sv[c];
            S1[t] += omega;
            S2[t] += omega;
            sum[c][t] += max3f(S0[t], S1[t], S2[t]);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < TILE; c++) {
          const int j = j0 + c;
          const bool nz = (zz >> c) & 1u, neg = (zz >> (16 + c)) & 1u;
#pragma unroll
          for (int t = 0; t < R; t++) {
            if (diag && j == r0 + t) {  // the row starts here from (0,0,0) (src/score.c:500-504)
              S0[t] = 0.0f;
              S1[t] = 0.0f;
              S2[t] = 0.0f;
            }
            float n0, n1, n2;
            if (!nz) {
              n0 = S0[t] + sv[c];
              n1 = S1[t] + omega;
              n2 = S2[t] + omega;
            } else if (!neg) {  // z = +1 (src/score.c:512-521)
              n0 = fmaxf(S0[t] + Delta, S2[t] + Omega);
              n1 = fmaxf(S0[t] + Omega, S1[t] + Delta);
              n2 = fmaxf(S1[t] + Omega, S2[t] + Delta);
            } else {  // z = -1 (src/score.c:523-533)
              n0 = fmaxf(S0[t] + Delta, S1[t] + Omega);
              n1 = fmaxf(S1[t] + Delta, S2[t] + Omega);
              n2 = fmaxf(S2[t] + Delta, S0[t] + Omega);
            }
            S0[t] = n0;
            S1[t] = n1;
            S2[t] = n2;
            sum[c][t] += max3f(n0, n1, n2);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < R; t++) {
        stk[(0 * R + t) * 32] = S0[t];
        stk[(1 * R + t) * 32] = S1[t];
        stk[(2 * R + t) * 32] = S2[t];
      }
    }
    __syncwarp();  // every lane has consumed this stage
    if (lane == 0 && tile + nst < ntiles) {
      unsigned char* dst = ring + s * stage_bytes;
      mbar_expect_tx(&bars[s], sig_tx + z_tx);
      bulk_g2s(dst, sig_src + (size_t)(tile + nst) * NK * TILE, sig_tx, &bars[s]);
      bulk_g2s(dst + DpSmem<R>::sig_bytes(smem_NK), z_src + (size_t)(tile + nst) * zstride, z_tx, &bars[s]);
    }

    // S[b][i] = max(sum, Delta) / (N-1)  (src/score.c:841-843; S[b][i-1], S[b][i-2] are always 0) and the
    // positive-entry filter of getHSS (:891).  The quotient is the correctly rounded IEEE one:
    // q = m*rcp, r = fma(-d, q, m), q' = fma(r, rcp, q) with rcp = RN(1/d) (checked exhaustively on the host).
    const bool chk = diag || (tile == ntiles - 1);
#pragma unroll
    for (int c = 0; c < TILE; c++) {
      const int j = j0 + c;
#pragma unroll
      for (int t = 0; t < R; t++) {
        const float m = fmaxf(sum[c][t], Delta);
        bool live = true;
        if (chk) live = (j >= r0 + t) && (j < sites);
        if (DENSE) {
          if (live) {
            const float q = m * rcpNK;
            const float e = __fmaf_rn(__fmaf_rn(-fNK, q, m), rcpNK, q);
            dense_row[t][j] = e;
          }
        } else if (m > 0.0f && live) {
          const float q = m * rcpNK;
          const float e = __fmaf_rn(__fmaf_rn(-fNK, q, m), rcpNK, q);
          rs[t] = hss_accept(rs[t], e, j, rec0 + t, band_slots);
        }
      }
    }
  }

  if (!DENSE) {
#pragma unroll
    for (int t = 0; t < R; t++) {
      if (r0 + t < sites) {
        RowRec* rec = rec0 + t;
        rec->Emax = rs[t].M;
        rec->vF = rs[t].lb;
        rec->jF = (unsigned short)rs[t].jF;
        rec->n = (unsigned short)rs[t].nb;
      }
    }
  }
  __syncwarp();  // every copy into the ring has been waited for: the ring is free for the next task
  }
}

// ---------------------------------------------------------------------------------------------
// (c) k_dp_reg: the DP for alignments with at most REG_MAX_NK scored species (the common case: the
// reference's examples have 3..9).  Same task decomposition, staging and getHSS digest as k_dp, but the
// loop nest is step-major: for each end codon the warp walks all species, so the state stays in
// registers for the whole task and the species sum is a scalar per row.  Each lane owns TWO rows and
// keeps every state as a float2 (row0, row1): the adds are issued as packed FADD2 (add.rn.f32x2, IEEE
// round-to-nearest per half, bit-identical to two FADDs), which halves the issue slots of the add
// chains.  sigma tiles are laid out [step][RS] with RS = roundup(NK+1, 4): NK sigma values followed by
// the step's z word (2 bits per species), so one row of broadcast LDS.128 feeds a step, and a step
// without any frameshift costs one test.  Steps with a frameshift branch per species (warp-uniform).
// The kernel requires Delta <= 0 (then max(sum, Delta) > 0 <=> sum > 0 and the quotient is taken from
// sum itself); the host routes Delta > 0 to k_dp.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2s(float2 a, float b) { return add2(a, make_float2(b, b)); }

// RowRec-resident variant of hss_accept: the fold state lives in the row's record (global memory, touched
// only on the rare positive entries), so the hot loop carries no per-row registers for it.
// rec->vF = last accepted value (-inf before the first), rec->Emax = row maximum, rec->jF, rec->n as in RowRec.
__device__ __forceinline__ void hss_accept_rec(RowRec* rec, float e, int j, int slots) {
  const float lb = rec->vF;
  const float d = e - lb;
  if (!(d >= -0.0001f)) return;
  rec->vF = e;
  rec->jF = (unsigned short)j;
  int nb = rec->n & 0xff, ovf = rec->n & 0x8000;
  float M = rec->Emax;
  if (e > M) {
    if (M - e < -0.0001f) nb = 0;
    M = e;
    rec->Emax = e;
  }
  if (nb > 0 && rec->be[nb - 1] == e) {
    rec->bj[nb - 1] = (unsigned short)j;
  } else {
    if (nb == slots) {
      int w = 0;
      for (int m = 0; m < nb; m++) {
        const float b = rec->be[m];
        if (!(b - M < -0.0001f)) {
          rec->be[w] = b;
          rec->bj[w] = rec->bj[m];
          w++;
        }
      }
      nb = w;
    }
    if (nb == slots) {
      ovf = 0x8000;
    } else {
      rec->be[nb] = e;
      rec->bj[nb] = (unsigned short)j;
      nb++;
    }
  }
  rec->n = (unsigned short)(nb | ovf);
}

template <int NK>
struct RegCfg {
  static constexpr int RS = (NK + 1 + 3) / 4 * 4;  // floats per step row: NK sigma + z word, padded to 16 bytes
  static constexpr int SIG_TILE = TILE * RS;
  static constexpr int STAGE_BYTES = SIG_TILE * 4;
};

// loads one step row (sigma values + z word) from shared memory at 32-bit shared address `a`
template <int NK>
__device__ __forceinline__ void reg_load_row(unsigned a, float (&sv)[RegCfg<NK>::RS]) {
#pragma unroll
  for (int q = 0; q < RegCfg<NK>::RS / 4; q++)
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(sv[4 * q]), "=f"(sv[4 * q + 1]), "=f"(sv[4 * q + 2]), "=f"(sv[4 * q + 3])
                 : "r"(a + 16 * q));
}

// frameshift update of one species (src/score.c:512-533)
__device__ __forceinline__ void reg_shift(bool neg, float Delta, float Omega, float2& a0, float2& a1, float2& a2) {
  // z = +1 (:512-521): (0<-2, 1<-0, 2<-1);  z = -1 (:523-533): (0<-1, 1<-2, 2<-0)
  const float2 x0 = neg ? a1 : a2, x1 = neg ? a2 : a0, x2 = neg ? a0 : a1;
  const float2 d0 = add2s(a0, Delta), d1 = add2s(a1, Delta), d2 = add2s(a2, Delta);
  const float2 o0 = add2s(x0, Omega), o1 = add2s(x1, Omega), o2 = add2s(x2, Omega);
  a0 = make_float2(fmaxf(d0.x, o0.x), fmaxf(d0.y, o0.y));
  a1 = make_float2(fmaxf(d1.x, o1.x), fmaxf(d1.y, o1.y));
  a2 = make_float2(fmaxf(d2.x, o2.x), fmaxf(d2.y, o2.y));
}

// One end codon for all species: state update + species sum (no getHSS test).  The kernel holds exactly
// one copy of this code (instruction-cache footprint matters more than the few uniform branches).
#ifndef RC_FLAG_UNIFIED
#define RC_FLAG_UNIFIED 0
#endif
// Branch-free update of one species that may (f) or may not have a frameshift at this codon; f and neg are
// warp-uniform.  n_i = max(S_i + p_i, X_i + q): without a frameshift p = (sigma, omega, omega) and q = -inf, so the
// second term drops out (max(d, -inf) == d); with one p = Delta, q = Omega and X is the state rotated in the
// direction of the shift (src/score.c:506-533).  Same float operations on the values that matter as the branchy
// form, no branch: the species of a group with a frameshift cost 24 instead of 6-19 instructions, in a straight line.
__device__ __forceinline__ void reg_any(bool f, bool neg, float sig, float omega, float Delta, float Omega, float2& a0,
                                        float2& a1, float2& a2) {
  const float p0 = f ? Delta : sig, p12 = f ? Delta : omega, q = f ? Omega : -INFINITY;
  const float2 x0 = neg ? a1 : a2, x1 = neg ? a2 : a0, x2 = neg ? a0 : a1;
  const float2 d0 = add2s(a0, p0), d1 = add2s(a1, p12), d2 = add2s(a2, p12);
  const float2 o0 = add2s(x0, q), o1 = add2s(x1, q), o2 = add2s(x2, q);
  a0 = make_float2(fmaxf(d0.x, o0.x), fmaxf(d0.y, o0.y));
  a1 = make_float2(fmaxf(d1.x, o1.x), fmaxf(d1.y, o1.y));
  a2 = make_float2(fmaxf(d2.x, o2.x), fmaxf(d2.y, o2.y));
}

// HAS_IN: the species sum continues a partial sum `sin` handed over by the warp that owns the preceding species
// (k_dp_chain); otherwise it starts with the first species (0 + m == m).
template <int NK, bool HAS_IN = false>
__device__ __forceinline__ float2 reg_update(float2 (&S0)[NK], float2 (&S1)[NK], float2 (&S2)[NK],
                                             const float (&sv)[RegCfg<NK>::RS], bool diag, int j, int r0, float Delta,
                                             float Omega, float omega, float2 sin = make_float2(0.0f, 0.0f)) {
  const unsigned zw = __float_as_uint(sv[NK]);
  if (diag) {
    // a row starts from (0,0,0) at its first end codon (src/score.c:500-504)
    if (j == r0) {
#pragma unroll
      for (int k = 0; k < NK; k++) S0[k].x = S1[k].x = S2[k].x = 0.0f;
    }
    if (j == r0 + 1) {
#pragma unroll
      for (int k = 0; k < NK; k++) S0[k].y = S1[k].y = S2[k].y = 0.0f;
    }
  }
  float2 sum;
  if (zw == 0u) {  // no species has a frameshift at this codon (src/score.c:506-510)
#pragma unroll
    for (int k = 0; k < NK; k++) {
      S0[k] = add2s(S0[k], sv[k]);
      S1[k] = add2s(S1[k], omega);
      S2[k] = add2s(S2[k], omega);
      const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
      sum = (k == 0) ? (HAS_IN ? add2(sin, m) : m) : add2(sum, m);  // species sum in k order (src/score.c:834-838); 0 + m == m
    }
  } else {
#if RC_FLAG_UNIFIED == 2
    // some species has a frameshift here: every species through the branch-free form
#pragma unroll
    for (int k = 0; k < NK; k++) {
      const unsigned z2 = (zw >> (2 * k)) & 3u;
      reg_any((z2 & 1u) != 0u, (z2 & 2u) != 0u, sv[k], omega, Delta, Omega, S0[k], S1[k], S2[k]);
      const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
      sum = (k == 0) ? (HAS_IN ? add2(sin, m) : m) : add2(sum, m);
    }
#else
    // some species has a frameshift here: test groups of three species; a hit group takes the branch-free form
#pragma unroll
    for (int g = 0; g < NK; g += 3) {
      constexpr unsigned GM3 = 0x3fu;
      const unsigned gm = (NK - g >= 3) ? GM3 : ((1u << (2 * (NK - g))) - 1u);
      if (((zw >> (2 * g)) & gm) == 0u) {
#pragma unroll
        for (int k = g; k < g + 3 && k < NK; k++) {
          S0[k] = add2s(S0[k], sv[k]);
          S1[k] = add2s(S1[k], omega);
          S2[k] = add2s(S2[k], omega);
        }
      } else {
#pragma unroll
        for (int k = g; k < g + 3 && k < NK; k++) {
          const unsigned z2 = (zw >> (2 * k)) & 3u;
#if RC_FLAG_UNIFIED
          reg_any((z2 & 1u) != 0u, (z2 & 2u) != 0u, sv[k], omega, Delta, Omega, S0[k], S1[k], S2[k]);
#else
          if (z2 == 0u) {
            S0[k] = add2s(S0[k], sv[k]);
            S1[k] = add2s(S1[k], omega);
            S2[k] = add2s(S2[k], omega);
          } else {
            reg_shift((z2 & 2u) != 0u, Delta, Omega, S0[k], S1[k], S2[k]);
          }
#endif
        }
      }
#pragma unroll
      for (int k = g; k < g + 3 && k < NK; k++) {
        const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
        sum = (k == 0) ? (HAS_IN ? add2(sin, m) : m) : add2(sum, m);
      }
    }
#endif
  }
  return sum;
}

// Two consecutive frameshift-free end codons in one straight-line block: the species-sum chain of the
// first overlaps with the state updates of the second.
template <int NK, bool HAS_IN = false>
__device__ __forceinline__ void reg_pair_fast(float2 (&S0)[NK], float2 (&S1)[NK], float2 (&S2)[NK],
                                              const float (&svA)[RegCfg<NK>::RS], const float (&svB)[RegCfg<NK>::RS],
                                              float omega, float2& sumA, float2& sumB,
                                              float2 sinA = make_float2(0.0f, 0.0f), float2 sinB = make_float2(0.0f, 0.0f)) {
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2s(S0[k], svA[k]);
    S1[k] = add2s(S1[k], omega);
    S2[k] = add2s(S2[k], omega);
    const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
    sumA = (k == 0) ? (HAS_IN ? add2(sinA, m) : m) : add2(sumA, m);
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2s(S0[k], svB[k]);
    S1[k] = add2s(S1[k], omega);
    S2[k] = add2s(S2[k], omega);
    const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
    sumB = (k == 0) ? (HAS_IN ? add2(sinB, m) : m) : add2(sumB, m);
  }
}

// Tiles in which the warp's rows start ("diagonal" tiles).  A row starts from (0,0,0) at its first end codon
// (src/score.c:500-504).  The states are zero-initialised, and until a row has started every addend that reaches it
// is replaced by +0 (per-lane selects on sigma, omega, Delta, Omega): 0 + 0 == 0 and max(0, 0) == 0 keep the state
// exactly (0,0,0) up to the row's first codon, from where the reference's operations apply unchanged.  This lets
// the diagonal tiles use the same two-codon straight-line blocks as the steady state.  With the lane's rows
// r0 (even) and r0+1 and a codon pair (j, j+1), j even:  P = (j >= r0): row r0 live at j, both rows live at j+1;
// Q = (j > r0): row r0+1 live at j.
template <int NK, bool HAS_IN = false>
__device__ __forceinline__ void reg_pair_diag(float2 (&S0)[NK], float2 (&S1)[NK], float2 (&S2)[NK],
                                              const float (&svA)[RegCfg<NK>::RS], const float (&svB)[RegCfg<NK>::RS],
                                              float omega, bool P, bool Q, float2& sumA, float2& sumB,
                                              float2 sinA = make_float2(0.0f, 0.0f), float2 sinB = make_float2(0.0f, 0.0f)) {
  const float2 omA = make_float2(P ? omega : 0.0f, Q ? omega : 0.0f);
  const float omB = P ? omega : 0.0f;
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2(S0[k], make_float2(P ? svA[k] : 0.0f, Q ? svA[k] : 0.0f));
    S1[k] = add2(S1[k], omA);
    S2[k] = add2(S2[k], omA);
    const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
    sumA = (k == 0) ? (HAS_IN ? add2(sinA, m) : m) : add2(sumA, m);
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2s(S0[k], P ? svB[k] : 0.0f);
    S1[k] = add2s(S1[k], omB);
    S2[k] = add2s(S2[k], omB);
    const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
    sumB = (k == 0) ? (HAS_IN ? add2(sinB, m) : m) : add2(sumB, m);
  }
}

// frameshift update of one species with per-row penalties (src/score.c:512-533)
__device__ __forceinline__ void reg_shift2(bool neg, float2 Delta, float2 Omega, float2& a0, float2& a1, float2& a2) {
  const float2 x0 = neg ? a1 : a2, x1 = neg ? a2 : a0, x2 = neg ? a0 : a1;
  const float2 d0 = add2(a0, Delta), d1 = add2(a1, Delta), d2 = add2(a2, Delta);
  const float2 o0 = add2(x0, Omega), o1 = add2(x1, Omega), o2 = add2(x2, Omega);
  a0 = make_float2(fmaxf(d0.x, o0.x), fmaxf(d0.y, o0.y));
  a1 = make_float2(fmaxf(d1.x, o1.x), fmaxf(d1.y, o1.y));
  a2 = make_float2(fmaxf(d2.x, o2.x), fmaxf(d2.y, o2.y));
}

// One end codon of a diagonal tile, any frameshift pattern; mx / my: the lane's two rows are live at this codon.
template <int NK, bool HAS_IN = false>
__device__ __forceinline__ float2 reg_update_diag(float2 (&S0)[NK], float2 (&S1)[NK], float2 (&S2)[NK],
                                                  const float (&sv)[RegCfg<NK>::RS], bool mx, bool my, float Delta,
                                                  float Omega, float omega, float2 sin = make_float2(0.0f, 0.0f)) {
  const unsigned zw = __float_as_uint(sv[NK]);
  const float2 om2 = make_float2(mx ? omega : 0.0f, my ? omega : 0.0f);
  const float2 D2 = make_float2(mx ? Delta : 0.0f, my ? Delta : 0.0f);
  const float2 O2 = make_float2(mx ? Omega : 0.0f, my ? Omega : 0.0f);
  float2 sum;
#pragma unroll
  for (int k = 0; k < NK; k++) {
    const unsigned z2 = (zw >> (2 * k)) & 3u;
    if (z2 == 0u) {
      S0[k] = add2(S0[k], make_float2(mx ? sv[k] : 0.0f, my ? sv[k] : 0.0f));
      S1[k] = add2(S1[k], om2);
      S2[k] = add2(S2[k], om2);
    } else {
      reg_shift2((z2 & 2u) != 0u, D2, O2, S0[k], S1[k], S2[k]);
    }
    const float2 m = make_float2(max3f(S0[k].x, S1[k].x, S2[k].x), max3f(S0[k].y, S1[k].y, S2[k].y));
    sum = (k == 0) ? (HAS_IN ? add2(sin, m) : m) : add2(sum, m);
  }
  return sum;
}

// getHSS only looks at positive entries (src/score.c:891); with Delta <= 0, max(sum, Delta) > 0 <=> sum > 0.
// S[b][i] = sum / (N-1) (src/score.c:841-843) as the correctly rounded quotient (see k_dp).  lb mirrors
// rec->vF (the fresh fold's last accepted value) in a register so that entries the fold rejects cost no
// memory access.  Out of line: positive entries are rare and the hot loop must stay small.
__device__ __noinline__ float reg_check_row(float sum, int j, int rstart, int sites, float fNK, float rcpNK, RowRec* rec,
                                            int band_slots, float lb) {
  if (sum > 0.0f && j >= rstart && j < sites) {
    const float q = sum * rcpNK;
    const float e = __fmaf_rn(__fmaf_rn(-fNK, q, sum), rcpNK, q);
    if (e - lb >= -0.0001f) {
      lb = e;
      const float M = rec->Emax;
      if (M - e < -0.0001f) {
        // the common case right after a row starts: a new row maximum beyond the old tie band -- the band restarts
        // with this entry (what hss_accept_rec does for it, without its band bookkeeping)
        const unsigned short ovf = rec->n & 0x8000;
        rec->Emax = e;
        rec->vF = e;
        rec->be[0] = e;
        rec->jF = (unsigned short)j;
        rec->n = (unsigned short)(1 | ovf);
        rec->bj[0] = (unsigned short)j;
      } else {
        hss_accept_rec(rec, e, j, band_slots);
      }
    }
  }
  return lb;
}

// Variant with the positive-entry filter and the fresh-fold test inline: only accepted entries leave the loop
// (four arguments instead of nine).  Used where positive entries are frequent (short rows, k_dp_smp).
__device__ __noinline__ void hss_accept_call(RowRec* rec, float e, int j, int band_slots) {
  const float M = rec->Emax;
  if (M - e < -0.0001f) {
    const unsigned short ovf = rec->n & 0x8000;
    rec->Emax = e;
    rec->vF = e;
    rec->be[0] = e;
    rec->jF = (unsigned short)j;
    rec->n = (unsigned short)(1 | ovf);
    rec->bj[0] = (unsigned short)j;
  } else {
    hss_accept_rec(rec, e, j, band_slots);
  }
}
__device__ __forceinline__ float reg_check_inl(float sum, int j, int rstart, int sites, float fNK, float rcpNK, RowRec* rec,
                                               int band_slots, float lb) {
  if (sum > 0.0f && j >= rstart && j < sites) {
    const float q = sum * rcpNK;
    const float e = __fmaf_rn(__fmaf_rn(-fNK, q, sum), rcpNK, q);
    if (e - lb >= -0.0001f) {
      lb = e;
      hss_accept_call(rec, e, j, band_slots);
    }
  }
  return lb;
}

// Sample-major kernels: 64 rows of 32 different alignments share a warp, so at almost every end codon SOME row has an
// accepted entry and a record update per entry (divergent, ~40 instructions, shared-memory read-modify-write) was half
// of k_dp_smp's time on 40-codon frames.  The common event -- a new row maximum beyond the tie band, after which the
// whole fold state is "one band entry (M, jF)" -- is therefore kept in registers (RowFold); the record is only
// materialised (fold_flush) when an entry falls inside the band or below the maximum, and once at the end of the row.
struct RowFold {
  float M;  // row maximum (= rec->Emax)
  int jF;   // >= 0: the state is the single entry (M, jF), last accepted value = M, and the record is stale; < 0: see the record
};
__device__ __forceinline__ void fold_init(RowFold& f) {
  f.M = -INFINITY;
  f.jF = -1;
}
__device__ __forceinline__ void fold_flush(RowRec* rec, float M, int jF) {
  const unsigned short ovf = rec->n & 0x8000;
  rec->Emax = M;
  rec->vF = M;
  rec->be[0] = M;
  rec->jF = (unsigned short)jF;
  rec->n = (unsigned short)(1 | ovf);
  rec->bj[0] = (unsigned short)jF;
}
__device__ __noinline__ void fold_slow(RowRec* rec, float e, int j, int band_slots, float M, int jF) {
  if (jF >= 0) fold_flush(rec, M, jF);
  hss_accept_rec(rec, e, j, band_slots);
}
__device__ __forceinline__ void fold_entry(float sum, int j, int rstart, int sites, float fNK, float rcpNK, RowRec* rec,
                                           int band_slots, RowFold& f) {
  if (sum > 0.0f && j >= rstart && j < sites) {
    const float q = sum * rcpNK;
    const float e = __fmaf_rn(__fmaf_rn(-fNK, q, sum), rcpNK, q);
    if (f.M - e < -0.0001f) {  // new maximum beyond the band (implies acceptance: last accepted value <= M)
      f.M = e;
      f.jF = j;
    } else {
      const float lb = f.jF >= 0 ? f.M : rec->vF;
      if (e - lb >= -0.0001f) {
        fold_slow(rec, e, j, band_slots, f.M, f.jF);
        f.M = fmaxf(f.M, e);
        f.jF = -1;
      }
    }
  }
}

// The fold state of the getHSS digest lives in a per-warp shared-memory copy of the two row records while
// the rows are being scored (positive entries are frequent at the start of a row, and a global-memory
// read-modify-write per entry would stall the warp); the records are written to HBM once, at the end of the rows.
__device__ __forceinline__ void rec_init(RowRec* r) {
  RowRec init;
  init.Emax = -INFINITY;
  init.vF = -INFINITY;
  init.be[0] = init.be[1] = init.be[2] = 0.0f;
  init.jF = 0;
  init.n = 0;
  init.bj[0] = init.bj[1] = init.bj[2] = 0;
  init.pad = 0;
  reinterpret_cast<uint4*>(r)[0] = reinterpret_cast<const uint4*>(&init)[0];
  reinterpret_cast<uint4*>(r)[1] = reinterpret_cast<const uint4*>(&init)[1];
}
__device__ __forceinline__ void rec_copy(RowRec* dst, const RowRec* src) {
  reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(src)[0];
  reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(src)[1];
}

#ifndef RC_REG_MINB
#define RC_REG_MINB 4
#endif
#ifndef RC_REG_CHECK_INL
#define RC_REG_CHECK_INL 1
#endif
#if RC_REG_CHECK_INL
#define RC_REG_CHECK reg_check_inl
#else
#define RC_REG_CHECK reg_check_row
#endif
#ifndef RC_REG_TILE
#define RC_REG_TILE 32  // end codons per TMA stage of k_dp_reg (a multiple of TILE; layout 1 rows are contiguous over tiles)
#endif
#ifndef RC_REG_DIAG_MASKED
#define RC_REG_DIAG_MASKED 1
#endif
#ifndef RC_REG_SINGLE
#define RC_REG_SINGLE 0
#endif
template <int NK>
__global__ void __launch_bounds__(DP_WARPS * 32, RC_REG_MINB)
    k_dp_reg(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const CtaDesc* __restrict__ ctas,
             const float* __restrict__ sigma, RowRec* __restrict__ recs, Params prm, int band_slots) {
  constexpr int R = 2;
  constexpr int RS = RegCfg<NK>::RS;
  constexpr int RT = RC_REG_TILE;
  constexpr int STAGE_BYTES = RT * RS * 4;
  // ring of two stages + the two mbarriers + room for the loop's read-ahead of two step rows
  __shared__ __align__(128) unsigned char smem[DP_WARPS][2 * STAGE_BYTES + 16 + 2 * RS * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CtaDesc cd = ctas[blockIdx.x];
  const Item& it = items[cd.item];
  const BlockDev& bd = blocks[it.block];
  const int strand = cd.sf / 3, frame = cd.sf % 3;
  const int sites = bd.sites[frame];
  const int nsteps = bd.ntiles[frame] * TILE;  // padded end codons of the frame: a multiple of RT (rows past `sites` are zero)
  const int ntiles = nsteps / RT;
  const int ngroups = (sites + 32 * R - 1) / (32 * R);
  const int task = cd.task0 + warp;
  if (task >= it.ninst * ngroups) return;
  const int inst_l = task / ngroups, g = task % ngroups;
  const int row_base = g * 32 * R;
  const int r0 = row_base + lane * R;

  unsigned char* ring = smem[warp];
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 2 * STAGE_BYTES);
  unsigned ring_a = smem_u32(ring);
  asm volatile("" : "+r"(ring_a));  // keep the shared-window address in a register (no per-step re-derivation)
  const float* sig_src = sigma + it.sigma_off[strand][frame] + (size_t)inst_l * nsteps * RS;
  const int t0 = row_base / RT;
  const int t_last_diag = (row_base + 32 * R - 1) / RT;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    for (int s = 0; s < 2 && t0 + s < ntiles; s++) {
      mbar_expect_tx(&bars[s], STAGE_BYTES);
      bulk_g2s(ring + s * STAGE_BYTES, sig_src + (size_t)(t0 + s) * RT * RS, STAGE_BYTES, &bars[s]);
    }
  }
  // fold state of the getHSS digest: two records per lane in shared memory, flushed at the end of the task
  __shared__ __align__(16) RowRec srec[DP_WARPS][R * 32];
  RowRec* rec0 = &srec[warp][lane * R];
  rec_init(rec0);
  rec_init(rec0 + 1);
  __syncwarp();

  float2 S0[NK], S1[NK], S2[NK];
#pragma unroll
  for (int k = 0; k < NK; k++) S0[k] = S1[k] = S2[k] = make_float2(0.0f, 0.0f);
  float2 lb = make_float2(-INFINITY, -INFINITY);
  const float Delta = prm.Delta, Omega = prm.Omega;
  float omega = prm.omega;
  asm volatile("" : "+f"(omega));  // a vector register operand for the packed adds instead of a constant reload per step
  const float fNK = bd.fNK, rcpNK = bd.rcpNK;

#pragma unroll 1
  for (int tile = t0; tile < ntiles; tile++) {
    const int s = (tile - t0) & 1;
    const unsigned parity = ((tile - t0) >> 1) & 1;
    const unsigned a0 = ring_a + s * STAGE_BYTES;
    const int j0 = tile * RT;
    constexpr int tsteps = RT;
    const bool diag = tile <= t_last_diag;  // rows start inside this tile
    mbar_wait(&bars[s], parity);
#if RC_REG_DIAG_MASKED
    if (diag) {
      // rows start inside the tile: masked two-codon blocks (reg_pair_diag), in a loop of their own
      float svA[RS], svB[RS];
      reg_load_row<NK>(a0, svA);
      reg_load_row<NK>(a0 + RS * 4, svB);
#pragma unroll 1
      for (int c = 0; c < tsteps; c += 2) {
        float2 sumA, sumB;
        const bool P = j0 + c >= r0, Q = j0 + c > r0;
        if ((__float_as_uint(svA[NK]) | __float_as_uint(svB[NK])) == 0u) {
          reg_pair_diag<NK>(S0, S1, S2, svA, svB, omega, P, Q, sumA, sumB);
        } else {
          sumA = reg_update_diag<NK>(S0, S1, S2, svA, P, Q, Delta, Omega, omega);
          sumB = reg_update_diag<NK>(S0, S1, S2, svB, P, P, Delta, Omega, omega);
        }
        reg_load_row<NK>(a0 + (c + 2) * RS * 4, svA);
        reg_load_row<NK>(a0 + (c + 3) * RS * 4, svB);
        if (fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 0.0f) {
          if (sumA.x > 0.0f) lb.x = RC_REG_CHECK(sumA.x, j0 + c, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumA.y > 0.0f) lb.y = RC_REG_CHECK(sumA.y, j0 + c, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
          if (sumB.x > 0.0f) lb.x = RC_REG_CHECK(sumB.x, j0 + c + 1, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumB.y > 0.0f) lb.y = RC_REG_CHECK(sumB.y, j0 + c + 1, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
        }
      }
    } else
#else
    if (diag) {
      // rows start inside the tile: one step at a time with the start-of-row reset
#pragma unroll 1
      for (int c = 0; c < tsteps; c++) {
        float sv[RS];
        reg_load_row<NK>(a0 + c * RS * 4, sv);
        const float2 sum = reg_update<NK>(S0, S1, S2, sv, true, j0 + c, r0, Delta, Omega, omega);
        if (fmaxf(sum.x, sum.y) > 0.0f) {
          lb.x = RC_REG_CHECK(sum.x, j0 + c, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          lb.y = RC_REG_CHECK(sum.y, j0 + c, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
        }
      }
    } else
#endif
    {
      // Two end codons per iteration.  Pairs without any frameshift run one straight-line block (the species-sum
      // chain of the first codon overlaps the state updates of the second); other pairs take the two codons one
      // after the other.  The rows of the next iteration are requested right after the arithmetic (unconditionally,
      // so that the loop carries them without register copies: past the tile they hit the other stage or the pad,
      // and are discarded).
      float svA[RS], svB[RS];
      reg_load_row<NK>(a0, svA);
      reg_load_row<NK>(a0 + RS * 4, svB);
#if RC_REG_SINGLE
      // A codon with a frameshift of some species is taken alone and the loop goes on from the next codon, so that the kernel
      // holds ONE copy of the branchy general update (instruction-cache footprint of the steady loop)
      int c = 0;
#pragma unroll 1
      while (c < tsteps) {
        float2 sumA, sumB = make_float2(0.0f, 0.0f);
        const int jc = j0 + c;
        const bool pair = c + 1 < tsteps && (__float_as_uint(svA[NK]) | __float_as_uint(svB[NK])) == 0u;
        if (pair) {
          reg_pair_fast<NK>(S0, S1, S2, svA, svB, omega, sumA, sumB);
          c += 2;
        } else {
          sumA = reg_update<NK>(S0, S1, S2, svA, false, jc, r0, Delta, Omega, omega);
          c += 1;
        }
        reg_load_row<NK>(a0 + c * RS * 4, svA);
        reg_load_row<NK>(a0 + (c + 1) * RS * 4, svB);
        if (fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 0.0f) {
          if (sumA.x > 0.0f) lb.x = RC_REG_CHECK(sumA.x, jc, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumA.y > 0.0f) lb.y = RC_REG_CHECK(sumA.y, jc, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
          if (sumB.x > 0.0f) lb.x = RC_REG_CHECK(sumB.x, jc + 1, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumB.y > 0.0f) lb.y = RC_REG_CHECK(sumB.y, jc + 1, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
        }
      }
#else
#pragma unroll 1
      for (int c = 0; c < tsteps; c += 2) {
        float2 sumA, sumB;
        const bool clean = (__float_as_uint(svA[NK]) | __float_as_uint(svB[NK])) == 0u;
        if (clean) {
          reg_pair_fast<NK>(S0, S1, S2, svA, svB, omega, sumA, sumB);
        } else {
          sumA = reg_update<NK>(S0, S1, S2, svA, false, j0 + c, r0, Delta, Omega, omega);
          sumB = reg_update<NK>(S0, S1, S2, svB, false, j0 + c + 1, r0, Delta, Omega, omega);
        }
        reg_load_row<NK>(a0 + (c + 2) * RS * 4, svA);
        reg_load_row<NK>(a0 + (c + 3) * RS * 4, svB);
        if (fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 0.0f) {
          if (sumA.x > 0.0f) lb.x = RC_REG_CHECK(sumA.x, j0 + c, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumA.y > 0.0f) lb.y = RC_REG_CHECK(sumA.y, j0 + c, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
          if (sumB.x > 0.0f) lb.x = RC_REG_CHECK(sumB.x, j0 + c + 1, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumB.y > 0.0f) lb.y = RC_REG_CHECK(sumB.y, j0 + c + 1, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
        }
      }
#endif
    }
    __syncwarp();
    if (lane == 0 && tile + 2 < ntiles) {
      mbar_expect_tx(&bars[s], STAGE_BYTES);
      bulk_g2s(ring + s * STAGE_BYTES, sig_src + (size_t)(tile + 2) * RT * RS, STAGE_BYTES, &bars[s]);
    }
  }
  // rows without any positive entry keep n == 0; accepted rows hold Emax / vF / jF / band
  RowRec* grec = recs + it.rec_off[strand][frame] + (size_t)inst_l * sites + r0;
#pragma unroll
  for (int t = 0; t < R; t++)
    if (r0 + t < sites) rec_copy(grec + t, rec0 + t);
}

// ---------------------------------------------------------------------------------------------
// (c) k_dp_regtu: k_dp_reg with THREE instead of four additions per cell.
// Without a frameshift a step adds sigma to S0 and the SAME constant omega to S1 and S2 (src/score.c:506-510), and the cell only
// needs max3(S0, S1, S2).  Rounding to nearest is monotone, so max(S1 (+) omega, S2 (+) omega) == max(S1, S2) (+) omega bit for
// bit: the kernel carries T = max(S1, S2) -- one addition per step -- and takes max(S0, T).  The smaller of the two, U, and which
// of them is S1 are only needed at the next frameshift of that species (src/score.c:512-533 mixes the three states): U is
// brought up to date there, n deferred additions of omega at once.  n roundings collapse because omega is -2^k (the default
// -2.0; the host sends every other value to k_dp_reg): a sum stays exactly representable while its magnitude does not leave
// the binade of the value the deferred run started from, so whole runs are one FMA, and a run that crosses a binade boundary
// is cut there -- one FMA up to and including the first rounded step, then on from the rounded value (catch_up; checked on
// the host against the step-by-step loop, bit for bit, on 2*10^7 random (value, k, n); tools/catch_up_check.c).
// Rows that started after the species' previous frameshift have S1 == S2 == T, no catching up.
// Everything else -- task shape, TMA ring, masked diagonal tiles, species sum in k order, getHSS fold -- is k_dp_reg's.
// ---------------------------------------------------------------------------------------------
// u (+) w, n times, w = -2^k (k = kexp).  See above; returns exactly what the loop `for (i < n) u = u + w` returns.
__device__ __forceinline__ float catch_up(float u, float w, int kexp, int n) {
  while (n > 0) {
    const unsigned bits = __float_as_uint(u);
    const int expf = (int)((bits >> 23) & 0xffu);
    const int shift = kexp + 150 - expf;  // log2(|w| / ulp(u))
    if (expf == 0 || shift < 0) {         // zero / denormal, or w finer than u's grid (every step rounds): one plain step
      u = u + w;
      n--;
      continue;
    }
    const float B = __uint_as_float((bits & 0x7f800000u) + 0x00800000u);  // top of u's binade: all multiples of ulp(u) up to B exist
    const float r = __fmaf_rn(w, (float)n, u);
    if (fabsf(r) <= B) return r;  // the run never leaves the binade (or ends on its first rounded step): exact
    const int m = (int)((bits & 0x7fffffu) | 0x800000u);
    const int num = ((bits >> 31) ? -m : m) + (1 << 24);  // (u + B) / ulp(u)
    const int J = shift >= 25 ? 0 : (num >> shift);       // steps that stay exact
    u = __fmaf_rn(w, (float)(J + 1), u);                  // ... and the first rounded one
    n -= J + 1;
  }
  return u;
}

struct TuEvent {
  float2 s0, t, u;
  unsigned bits;  // bit 0: row x has S1 >= S2 (S1 is T), bit 1: the same for row y
};
// Frameshift of one species at one end codon (src/score.c:512-533) on the (S0, T, U) form.  n: clean steps since U was last
// current (warp-uniform); fx / fy: the row started after that point (S1 == S2 == T); D2 / O2: Delta / Omega per row, +0 for a row
// that has not started yet (the state stays (0,0,0), see reg_pair_diag).
__device__ __noinline__ TuEvent tu_event(float2 s0, float2 t, float2 u, unsigned bits, bool neg, int n, bool fx, bool fy, float2 D2,
                                         float2 O2, float omega, int kexp) {
  u.x = fx ? t.x : catch_up(u.x, omega, kexp, n);
  u.y = fy ? t.y : catch_up(u.y, omega, kexp, n);
  const bool bx = (bits & 1u) != 0u, by = (bits & 2u) != 0u;
  float2 a0 = s0;
  float2 a1 = make_float2(bx ? t.x : u.x, by ? t.y : u.y);
  float2 a2 = make_float2(bx ? u.x : t.x, by ? u.y : t.y);
  reg_shift2(neg, D2, O2, a0, a1, a2);
  TuEvent e;
  e.s0 = a0;
  e.t = make_float2(fmaxf(a1.x, a2.x), fmaxf(a1.y, a2.y));
  e.u = make_float2(fminf(a1.x, a2.x), fminf(a1.y, a2.y));
  e.bits = (a1.x >= a2.x ? 1u : 0u) | (a1.y >= a2.y ? 2u : 0u);
  return e;
}

// two frameshift-free end codons, all rows live
template <int NK>
__device__ __forceinline__ void tu_pair_fast(float2 (&S0)[NK], float2 (&T)[NK], const float (&svA)[RegCfg<NK>::RS],
                                             const float (&svB)[RegCfg<NK>::RS], float omega, float2& sumA, float2& sumB) {
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2s(S0[k], svA[k]);
    T[k] = add2s(T[k], omega);
    const float2 m = make_float2(fmaxf(S0[k].x, T[k].x), fmaxf(S0[k].y, T[k].y));
    sumA = (k == 0) ? m : add2(sumA, m);
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2s(S0[k], svB[k]);
    T[k] = add2s(T[k], omega);
    const float2 m = make_float2(fmaxf(S0[k].x, T[k].x), fmaxf(S0[k].y, T[k].y));
    sumB = (k == 0) ? m : add2(sumB, m);
  }
}
// the same in a tile in which rows start (masks as in reg_pair_diag)
template <int NK>
__device__ __forceinline__ void tu_pair_diag(float2 (&S0)[NK], float2 (&T)[NK], const float (&svA)[RegCfg<NK>::RS],
                                             const float (&svB)[RegCfg<NK>::RS], float omega, bool P, bool Q, float2& sumA,
                                             float2& sumB) {
  const float2 omA = make_float2(P ? omega : 0.0f, Q ? omega : 0.0f);
  const float omB = P ? omega : 0.0f;
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2(S0[k], make_float2(P ? svA[k] : 0.0f, Q ? svA[k] : 0.0f));
    T[k] = add2(T[k], omA);
    const float2 m = make_float2(fmaxf(S0[k].x, T[k].x), fmaxf(S0[k].y, T[k].y));
    sumA = (k == 0) ? m : add2(sumA, m);
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = add2s(S0[k], P ? svB[k] : 0.0f);
    T[k] = add2s(T[k], omB);
    const float2 m = make_float2(fmaxf(S0[k].x, T[k].x), fmaxf(S0[k].y, T[k].y));
    sumB = (k == 0) ? m : add2(sumB, m);
  }
}
// one end codon, any frameshift pattern; mx / my: the lane's rows are live at this codon (both true outside diagonal tiles)
// U_a: shared address of this lane's U of species 0 ([species][lane] float2, 256 bytes per species); tU_a: of the warp's
// tU[0] -- both only touched at a frameshift, so they stay out of the registers of the step loop
template <int NK>
__device__ __forceinline__ float2 tu_update(float2 (&S0)[NK], float2 (&T)[NK], unsigned U_a, unsigned& bx, unsigned& by,
                                            unsigned tU_a, const float (&sv)[RegCfg<NK>::RS], bool mx, bool my, int j, int r0,
                                            float Delta, float Omega, float omega, int kexp) {
  const unsigned zw = __float_as_uint(sv[NK]);
  const float2 om2 = make_float2(mx ? omega : 0.0f, my ? omega : 0.0f);
  float2 sum;
#pragma unroll
  for (int k = 0; k < NK; k++) {
    const unsigned z2 = (zw >> (2 * k)) & 3u;
    if (z2 == 0u) {
      S0[k] = add2(S0[k], make_float2(mx ? sv[k] : 0.0f, my ? sv[k] : 0.0f));
      T[k] = add2(T[k], om2);
    } else {
      const float2 D2 = make_float2(mx ? Delta : 0.0f, my ? Delta : 0.0f);
      const float2 O2 = make_float2(mx ? Omega : 0.0f, my ? Omega : 0.0f);
      const int tk = (int)lds_u32(tU_a + 4u * k);
      const TuEvent e = tu_event(S0[k], T[k], lds_f2(U_a + 256u * k), ((bx >> k) & 1u) | (((by >> k) & 1u) << 1), (z2 & 2u) != 0u,
                                 j - 1 - tk, r0 > tk, r0 + 1 > tk, D2, O2, omega, kexp);
      S0[k] = e.s0;
      T[k] = e.t;
      sts_f2(U_a + 256u * k, e.u);
      bx = (bx & ~(1u << k)) | ((e.bits & 1u) << k);
      by = (by & ~(1u << k)) | (((e.bits >> 1) & 1u) << k);
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(tU_a + 4u * k), "r"(j) : "memory");  // every lane writes the same value
    }
    const float2 m = make_float2(fmaxf(S0[k].x, T[k].x), fmaxf(S0[k].y, T[k].y));
    sum = (k == 0) ? m : add2(sum, m);
  }
  return sum;
}

template <int NK>
__global__ void __launch_bounds__(DP_WARPS * 32, RC_REG_MINB)
    k_dp_regtu(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const CtaDesc* __restrict__ ctas,
               const float* __restrict__ sigma, RowRec* __restrict__ recs, Params prm, int band_slots) {
  constexpr int R = 2;
  constexpr int RS = RegCfg<NK>::RS;
  constexpr int RT = RC_REG_TILE;
  constexpr int STAGE_BYTES = RT * RS * 4;
  __shared__ __align__(128) unsigned char smem[DP_WARPS][2 * STAGE_BYTES + 16 + 2 * RS * 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CtaDesc cd = ctas[blockIdx.x];
  const Item& it = items[cd.item];
  const BlockDev& bd = blocks[it.block];
  const int strand = cd.sf / 3, frame = cd.sf % 3;
  const int sites = bd.sites[frame];
  const int nsteps = bd.ntiles[frame] * TILE;
  const int ntiles = nsteps / RT;
  const int ngroups = (sites + 32 * R - 1) / (32 * R);
  const int task = cd.task0 + warp;
  if (task >= it.ninst * ngroups) return;
  const int inst_l = task / ngroups, g = task % ngroups;
  const int row_base = g * 32 * R;
  const int r0 = row_base + lane * R;

  unsigned char* ring = smem[warp];
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 2 * STAGE_BYTES);
  unsigned ring_a = smem_u32(ring);
  asm volatile("" : "+r"(ring_a));
  const float* sig_src = sigma + it.sigma_off[strand][frame] + (size_t)inst_l * nsteps * RS;
  const int t0 = row_base / RT;
  const int t_last_diag = (row_base + 32 * R - 1) / RT;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
    for (int s = 0; s < 2 && t0 + s < ntiles; s++) {
      mbar_expect_tx(&bars[s], STAGE_BYTES);
      bulk_g2s(ring + s * STAGE_BYTES, sig_src + (size_t)(t0 + s) * RT * RS, STAGE_BYTES, &bars[s]);
    }
  }
  __shared__ __align__(16) RowRec srec[DP_WARPS][R * 32];
  RowRec* rec0 = &srec[warp][lane * R];
  rec_init(rec0);
  rec_init(rec0 + 1);
  __syncwarp();

  // U per (species, lane) and tU per species (end codon of the species' last frameshift seen by this warp; -1: none yet)
  __shared__ __align__(16) float2 s_U[DP_WARPS][NK][32];
  __shared__ int s_tU[DP_WARPS][NK];
  float2 S0[NK], T[NK];
#pragma unroll
  for (int k = 0; k < NK; k++) {
    S0[k] = T[k] = make_float2(0.0f, 0.0f);
    s_U[warp][k][lane] = make_float2(0.0f, 0.0f);
    s_tU[warp][k] = -1;
  }
  unsigned U_a = smem_u32(&s_U[warp][0][lane]), tU_a = smem_u32(&s_tU[warp][0]);
  asm volatile("" : "+r"(U_a), "+r"(tU_a));
  __syncwarp();
  unsigned bx = 0u, by = 0u;
  float2 lb = make_float2(-INFINITY, -INFINITY);
  const float Delta = prm.Delta, Omega = prm.Omega;
  float omega = prm.omega;
  asm volatile("" : "+f"(omega));
  const int kexp = (int)((__float_as_uint(prm.omega) >> 23) & 0xffu) - 127;
  const float fNK = bd.fNK, rcpNK = bd.rcpNK;

#pragma unroll 1
  for (int tile = t0; tile < ntiles; tile++) {
    const int s = (tile - t0) & 1;
    const unsigned parity = ((tile - t0) >> 1) & 1;
    const unsigned a0 = ring_a + s * STAGE_BYTES;
    const int j0 = tile * RT;
    const bool diag = tile <= t_last_diag;  // rows start inside this tile
    mbar_wait(&bars[s], parity);
    float svA[RS], svB[RS];
    reg_load_row<NK>(a0, svA);
    reg_load_row<NK>(a0 + RS * 4, svB);
#pragma unroll 1
    for (int c = 0; c < RT; c += 2) {
      float2 sumA, sumB;
      const bool clean = (__float_as_uint(svA[NK]) | __float_as_uint(svB[NK])) == 0u;
      if (diag) {
        const bool P = j0 + c >= r0, Q = j0 + c > r0;
        if (clean) {
          tu_pair_diag<NK>(S0, T, svA, svB, omega, P, Q, sumA, sumB);
        } else {
          sumA = tu_update<NK>(S0, T, U_a, bx, by, tU_a, svA, P, Q, j0 + c, r0, Delta, Omega, omega, kexp);
          sumB = tu_update<NK>(S0, T, U_a, bx, by, tU_a, svB, P, P, j0 + c + 1, r0, Delta, Omega, omega, kexp);
        }
      } else if (clean) {
        tu_pair_fast<NK>(S0, T, svA, svB, omega, sumA, sumB);
      } else {
        sumA = tu_update<NK>(S0, T, U_a, bx, by, tU_a, svA, true, true, j0 + c, r0, Delta, Omega, omega, kexp);
        sumB = tu_update<NK>(S0, T, U_a, bx, by, tU_a, svB, true, true, j0 + c + 1, r0, Delta, Omega, omega, kexp);
      }
      reg_load_row<NK>(a0 + (c + 2) * RS * 4, svA);
      reg_load_row<NK>(a0 + (c + 3) * RS * 4, svB);
      if (fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 0.0f) {
        if (sumA.x > 0.0f) lb.x = RC_REG_CHECK(sumA.x, j0 + c, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
        if (sumA.y > 0.0f) lb.y = RC_REG_CHECK(sumA.y, j0 + c, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
        if (sumB.x > 0.0f) lb.x = RC_REG_CHECK(sumB.x, j0 + c + 1, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
        if (sumB.y > 0.0f) lb.y = RC_REG_CHECK(sumB.y, j0 + c + 1, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
      }
    }
    __syncwarp();
    if (lane == 0 && tile + 2 < ntiles) {
      mbar_expect_tx(&bars[s], STAGE_BYTES);
      bulk_g2s(ring + s * STAGE_BYTES, sig_src + (size_t)(tile + 2) * RT * RS, STAGE_BYTES, &bars[s]);
    }
  }
  RowRec* grec = recs + it.rec_off[strand][frame] + (size_t)inst_l * sites + r0;
#pragma unroll
  for (int t = 0; t < R; t++)
    if (r0 + t < sites) rec_copy(grec + t, rec0 + t);
}

// ---------------------------------------------------------------------------------------------
// (c) k_dp_chain: the register-resident DP for WIDE alignments (N-1 > 16).  The species are cut into W chunks
// of at most NK species; the CTA has W warps that all work on the same task (instance, strand, frame, 64 start
// codons), warp g owning chunk g with its 3*NK float2 states in registers exactly like k_dp_reg.  The species
// sum of the reference is one k-ordered float chain (src/score.c:834-838), so it is pipelined through the warps:
// warp g continues the partial sums warp g-1 left in shared memory for every (end codon, row) of a tile
// (RC_CHAIN_STAGES hand-off buffers with full/empty mbarriers per boundary), adds its own species in order and passes the
// result on; the last warp owns the getHSS digest.  Warp g therefore runs about one tile behind warp g-1.
// Chunks that are one species short carry a dummy species (sigma = +0, z = 0, state 0): with omega <= 0 its
// contribution max3(0, t*omega, t*omega) = +0 leaves the sum unchanged (x + 0 == x).
// sigma layout 3: [instance][tile][chunk][step][RS] (RS = NK sigma values + the chunk's z word, padded to 16 B).
// ---------------------------------------------------------------------------------------------
constexpr int CHAIN_MAX_WARPS = 16;   // warps of one k_dp_chain CTA (launch bound)
#ifndef RC_CHAIN_PASS_WARPS
#define RC_CHAIN_PASS_WARPS 5
#endif
constexpr int CHAIN_PASS_WARPS = RC_CHAIN_PASS_WARPS;   // chunks per pass when an alignment has more (see k_dp_chain)
constexpr int CHAIN_MAX_TASKS = 8;  // upper bound of BlockDev.chain_tasks
#ifndef RC_CHAIN_STAGES
#define RC_CHAIN_STAGES 3
#endif

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int NK>
struct ChainCfg {
  static constexpr int RS = RegCfg<NK>::RS;
  static constexpr int STAGE_BYTES = RegCfg<NK>::STAGE_BYTES;
  static constexpr int RING_BYTES = 2 * STAGE_BYTES + 16 + 2 * RS * 4;  // two stages, two mbarriers, read-ahead pad
  static constexpr int HAND_STAGES = RC_CHAIN_STAGES;                   // hand-off buffers per boundary
  static constexpr int HAND_BYTES = HAND_STAGES * TILE * 32 * 8;        // per boundary: [stage][step][lane] float2
  static __host__ __device__ size_t smem_bytes(int W) {
    return (size_t)W * RING_BYTES + (size_t)(W - 1) * HAND_BYTES + (size_t)(W - 1) * 16 * HAND_STAGES + 64 * sizeof(RowRec);
  }
};

#ifndef RC_CHAIN_MAXREG
#define RC_CHAIN_MAXREG 128
#endif
#ifndef RC_CHAIN_SINGLE
#define RC_CHAIN_SINGLE 0
#endif
#ifndef RC_CHAIN_MERGE
#define RC_CHAIN_MERGE 0
#endif
// PF / PL (compile-time, per launch): this launch is the FIRST / LAST pass of the alignment's species chunks.  Only the first
// warp of a later pass reads partial sums from global memory, only the last warp of an earlier pass writes them, only the last
// warp of the last pass owns the getHSS fold: a pass compiled without the paths it can never take needs fewer registers (the
// step loop spilled at the 128 that three resident CTAs allow).  MULTI = false: a single pass, PF = PL = true.
template <int NK, bool MULTI, bool PF = true, bool PL = true>
__global__ void
#if RC_CHAIN_MAXREG < 128
    __maxnreg__(RC_CHAIN_MAXREG)
#else
    __launch_bounds__(CHAIN_MAX_WARPS * 32)
#endif
    k_dp_chain(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const CtaDesc* __restrict__ ctas,
               const float* __restrict__ sigma, RowRec* __restrict__ recs, Params prm, int band_slots, int c_lo,
               float2* __restrict__ partial) {
  constexpr int R = 2;
  constexpr int RS = ChainCfg<NK>::RS;
  constexpr int SIG_TILE = RegCfg<NK>::SIG_TILE;
  constexpr int STAGE_BYTES = ChainCfg<NK>::STAGE_BYTES;
  constexpr int RING_BYTES = ChainCfg<NK>::RING_BYTES;
  constexpr int HAND_BYTES = ChainCfg<NK>::HAND_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  const int W = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CtaDesc cd = ctas[blockIdx.x];
  const Item& it = items[cd.item];
  const BlockDev& bd = blocks[it.block];
  const int strand = cd.sf / 3, frame = cd.sf % 3;
  const int sites = bd.sites[frame], ntiles = bd.ntiles[frame];
  const int ngroups = (sites + 32 * R - 1) / (32 * R);
  const int ntasks = it.ninst * ngroups;
  const bool first = warp == 0, last = warp == W - 1;  // ends of this CTA's warp pipeline
  // Very wide alignments run in several PASSES (launches) of at most 5 chunks each, so that three or four CTAs stay
  // resident per SM: the warps of a launch own the chunks c_lo .. c_lo + W - 1, the first warp of a later pass continues
  // the partial sums the previous pass left in global memory, the last warp of an earlier pass leaves them there.
  const int chunk = c_lo + warp;
  // ends of the whole species chain: chunk == 0 <=> first && PF, chunk == bd.nchunk - 1 <=> last && PL
  // (MULTI = false: single-pass launch, PF = PL = true, the global hand-over code compiles away)

  unsigned char* ring = smem + (size_t)warp * RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 2 * STAGE_BYTES);
  unsigned char* hand = smem + (size_t)W * RING_BYTES;
  constexpr int HS = ChainCfg<NK>::HAND_STAGES;
  uint64_t* hbar = reinterpret_cast<uint64_t*>(hand + (size_t)(W - 1) * HAND_BYTES);  // [boundary][full 0..HS-1, empty 0..HS-1]
  RowRec* srec = reinterpret_cast<RowRec*>(hbar + 2 * HS * (W - 1));
  unsigned ring_a = smem_u32(ring);
  asm volatile("" : "+r"(ring_a));
  const size_t tile_stride = (size_t)bd.sig_tile;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    if (!last)
      for (int q = 0; q < 2 * HS; q++) mbar_init(&hbar[2 * HS * warp + q], 1);
    mbar_fence_init();
  }
  RowRec* rec0 = &srec[lane * R];
  __syncthreads();  // hand-off barriers are initialised before any neighbour touches them

  const float Delta = prm.Delta, Omega = prm.Omega;
  float omega = prm.omega;
  asm volatile("" : "+f"(omega));
  const float fNK = bd.fNK, rcpNK = bd.rcpNK;
  const unsigned hand_in = smem_u32(hand) + (unsigned)((warp - 1) * HAND_BYTES) + lane * 8;  // valid for warp > 0
  const unsigned hand_out = smem_u32(hand) + (unsigned)(warp * HAND_BYTES) + lane * 8;       // valid for warp < W-1
  uint64_t* full_in = &hbar[2 * HS * (warp - 1)];
  uint64_t* empty_in = full_in + HS;
  uint64_t* full_out = &hbar[2 * HS * warp];
  uint64_t* empty_out = full_out + HS;
  int hs = 0;            // hand-off stage of the current tile, and how often it has been used before
  unsigned hround = 0;
  unsigned ring_it = 0;  // tiles this warp has consumed so far (sigma ring stage = ring_it & 1, phase = ring_it >> 1)

  // The CTA works through bd.chain_tasks consecutive tasks.  The warps only meet through the hand-off barriers, so the
  // first warps start the next task while the last ones finish the current one: the pipeline drains once per CTA,
  // not once per task (rows of a 1000-column block are only ~20 tiles long, the pipeline is W-1 tiles deep).
#pragma unroll 1
  for (int task = cd.task0; task < min(cd.task0 + bd.chain_tasks, ntasks); task++) {
  const int inst_l = task / ngroups, g = task % ngroups;
  const int row_base = g * 32 * R;
  const int r0 = row_base + lane * R;
  const float* sig_src = sigma + it.sigma_off[strand][frame] + (size_t)inst_l * ntiles * bd.sig_tile + (size_t)chunk * SIG_TILE;
  const int t0 = row_base / TILE;
  const int t_last_diag = (row_base + 32 * R - 1) / TILE;
  // partial sums between passes: [instance][row group][tile from the group's first][end codon][lane] float2
  float2* gp = nullptr;
  if (MULTI && ((first && !PF) || (last && !PL))) {
    const size_t per_inst = (size_t)ngroups * ntiles - 2 * (size_t)ngroups * (ngroups - 1);  // tiles
    gp = partial + it.part_off[strand][frame] +
         ((size_t)inst_l * per_inst + (size_t)g * ntiles - 2 * (size_t)g * (g - 1) - t0) * (TILE * 32) + lane;  // + tile * TILE * 32
  }
  if (lane == 0) {
    for (int q = 0; q < 2 && t0 + q < ntiles; q++) {
      const unsigned sq = (ring_it + q) & 1u;
      mbar_expect_tx(&bars[sq], STAGE_BYTES);
      bulk_g2s(ring + sq * STAGE_BYTES, sig_src + (size_t)(t0 + q) * tile_stride, STAGE_BYTES, &bars[sq]);
    }
  }
  if (PL && last) {
    rec_init(rec0);
    rec_init(rec0 + 1);
    __syncwarp();
  }
  float2 S0[NK], S1[NK], S2[NK];
#pragma unroll
  for (int k = 0; k < NK; k++) S0[k] = S1[k] = S2[k] = make_float2(0.0f, 0.0f);
  float2 lb = make_float2(-INFINITY, -INFINITY);

#pragma unroll 1
  for (int tile = t0; tile < ntiles; tile++, ring_it++) {
    const unsigned s = ring_it & 1u;
    const unsigned parity = (ring_it >> 1) & 1u;
    const unsigned a0 = ring_a + s * STAGE_BYTES;
    const unsigned hin = hand_in + hs * (TILE * 256), hout = hand_out + hs * (TILE * 256);
    const int j0 = tile * TILE;
    const bool diag = tile <= t_last_diag;
    mbar_wait(&bars[s], parity);
    if (!first) mbar_wait(&full_in[hs], hround & 1u);                           // partial sums of this tile have arrived
    if (!last && hround > 0) mbar_wait(&empty_out[hs], (hround & 1u) ^ 1u);     // the next warp is done with this stage
#if RC_CHAIN_SINGLE
    {
      // One copy of each update in the loop (instruction-cache footprint: the five warps of a CTA are in different phases of
      // it): a codon with a frameshift, and every codon of a tile in which rows start, is taken alone.
      float svA[RS], svB[RS];
      reg_load_row<NK>(a0, svA);
      reg_load_row<NK>(a0 + RS * 4, svB);
      int c = 0;
#pragma unroll 1
      while (c < TILE) {
        const int c0 = c;
        float2 sumA, sumB = make_float2(0.0f, 0.0f);
        float2 sinA = make_float2(0.0f, 0.0f), sinB = make_float2(0.0f, 0.0f);
        const bool pair = !diag && c + 1 < TILE && (__float_as_uint(svA[NK]) | __float_as_uint(svB[NK])) == 0u;
        if (!first) {
          sinA = lds_f2(hin + c * 256);
          if (pair) sinB = lds_f2(hin + (c + 1) * 256);
        } else if (MULTI && !PF) {
          sinA = gp[((size_t)tile * TILE + c) * 32];
          if (pair) sinB = gp[((size_t)tile * TILE + c + 1) * 32];
        }
        if (pair) {
          reg_pair_fast<NK, true>(S0, S1, S2, svA, svB, omega, sumA, sumB, sinA, sinB);
          c += 2;
        } else if (!diag) {
          sumA = reg_update<NK, true>(S0, S1, S2, svA, false, j0 + c, r0, Delta, Omega, omega, sinA);
          c += 1;
        } else {
          sumA = reg_update_diag<NK, true>(S0, S1, S2, svA, j0 + c >= r0, j0 + c > r0, Delta, Omega, omega, sinA);
          c += 1;
        }
        reg_load_row<NK>(a0 + c * RS * 4, svA);
        reg_load_row<NK>(a0 + (c + 1) * RS * 4, svB);
        if (!last) {
          sts_f2(hout + c0 * 256, sumA);
          if (pair) sts_f2(hout + (c0 + 1) * 256, sumB);
        } else if (MULTI && !PL) {
          gp[((size_t)tile * TILE + c0) * 32] = sumA;
          if (pair) gp[((size_t)tile * TILE + c0 + 1) * 32] = sumB;
        } else if (PL && fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 0.0f) {
          if (sumA.x > 0.0f) lb.x = RC_REG_CHECK(sumA.x, j0 + c0, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumA.y > 0.0f) lb.y = RC_REG_CHECK(sumA.y, j0 + c0, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
          if (sumB.x > 0.0f) lb.x = RC_REG_CHECK(sumB.x, j0 + c0 + 1, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumB.y > 0.0f) lb.y = RC_REG_CHECK(sumB.y, j0 + c0 + 1, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
        }
      }
    }
#else
    {
      float svA[RS], svB[RS];
      reg_load_row<NK>(a0, svA);
      reg_load_row<NK>(a0 + RS * 4, svB);
#pragma unroll 1
      for (int c = 0; c < TILE; c += 2) {
        float2 sumA, sumB;
        float2 sinA = make_float2(0.0f, 0.0f), sinB = make_float2(0.0f, 0.0f);
        if (!first) {
          sinA = lds_f2(hin + c * 256);
          sinB = lds_f2(hin + (c + 1) * 256);
        } else if (MULTI && !PF) {
          sinA = gp[((size_t)tile * TILE + c) * 32];
          sinB = gp[((size_t)tile * TILE + c + 1) * 32];
        }
        const bool clean = (__float_as_uint(svA[NK]) | __float_as_uint(svB[NK])) == 0u;
#if RC_CHAIN_MERGE
        // One copy of the general update for codons with a frameshift, in diagonal tiles or not (outside them the per-row
        // masks are simply true), taken twice in a loop: the kernel shrinks from 47 KB to about 32 KB of SASS -- the five warps
        // of a CTA are in different phases of the loop and `no_instruction` was its largest stall.
        if (!clean) {
          const bool P = !diag || j0 + c >= r0, Q = !diag || j0 + c > r0;
#pragma unroll 1
          for (int h = 0; h < 2; h++) {
            float sv[RS];
#pragma unroll
            for (int q = 0; q < RS; q++) sv[q] = h ? svB[q] : svA[q];
            const float2 sum = reg_update_diag<NK, true>(S0, S1, S2, sv, P, h ? P : Q, Delta, Omega, omega, h ? sinB : sinA);
            if (h) sumB = sum;
            else sumA = sum;
          }
        } else if (!diag) {
          reg_pair_fast<NK, true>(S0, S1, S2, svA, svB, omega, sumA, sumB, sinA, sinB);
        } else {
          reg_pair_diag<NK, true>(S0, S1, S2, svA, svB, omega, j0 + c >= r0, j0 + c > r0, sumA, sumB, sinA, sinB);
        }
#else
        if (!diag) {
          if (clean) {
            reg_pair_fast<NK, true>(S0, S1, S2, svA, svB, omega, sumA, sumB, sinA, sinB);
          } else {
            sumA = reg_update<NK, true>(S0, S1, S2, svA, false, j0 + c, r0, Delta, Omega, omega, sinA);
            sumB = reg_update<NK, true>(S0, S1, S2, svB, false, j0 + c + 1, r0, Delta, Omega, omega, sinB);
          }
        } else {
          const bool P = j0 + c >= r0, Q = j0 + c > r0;
          if (clean) {
            reg_pair_diag<NK, true>(S0, S1, S2, svA, svB, omega, P, Q, sumA, sumB, sinA, sinB);
          } else {
            sumA = reg_update_diag<NK, true>(S0, S1, S2, svA, P, Q, Delta, Omega, omega, sinA);
            sumB = reg_update_diag<NK, true>(S0, S1, S2, svB, P, P, Delta, Omega, omega, sinB);
          }
        }
#endif
        reg_load_row<NK>(a0 + (c + 2) * RS * 4, svA);
        reg_load_row<NK>(a0 + (c + 3) * RS * 4, svB);
        if (!last) {
          sts_f2(hout + c * 256, sumA);
          sts_f2(hout + (c + 1) * 256, sumB);
        } else if (MULTI && !PL) {
          gp[((size_t)tile * TILE + c) * 32] = sumA;
          gp[((size_t)tile * TILE + c + 1) * 32] = sumB;
        } else if (PL && fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 0.0f) {
          if (sumA.x > 0.0f) lb.x = RC_REG_CHECK(sumA.x, j0 + c, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumA.y > 0.0f) lb.y = RC_REG_CHECK(sumA.y, j0 + c, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
          if (sumB.x > 0.0f) lb.x = RC_REG_CHECK(sumB.x, j0 + c + 1, r0, sites, fNK, rcpNK, rec0, band_slots, lb.x);
          if (sumB.y > 0.0f) lb.y = RC_REG_CHECK(sumB.y, j0 + c + 1, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, lb.y);
        }
      }
    }
#endif
    __syncwarp();
    if (lane == 0) {
      if (tile + 2 < ntiles) {
        mbar_expect_tx(&bars[s], STAGE_BYTES);
        bulk_g2s(ring + s * STAGE_BYTES, sig_src + (size_t)(tile + 2) * tile_stride, STAGE_BYTES, &bars[s]);
      }
      if (!first) mbar_arrive(&empty_in[hs]);
      if (!last) mbar_arrive(&full_out[hs]);
    }
    if (++hs == HS) {
      hs = 0;
      hround++;
    }
  }
  if (PL && last) {
    RowRec* grec = recs + it.rec_off[strand][frame] + (size_t)inst_l * sites + r0;
#pragma unroll
    for (int t = 0; t < R; t++)
      if (r0 + t < sites) rec_copy(grec + t, rec0 + t);
    __syncwarp();
  }
  }
}

// ---------------------------------------------------------------------------------------------
// getHSS fold of the sample-major kernels in the space of the species sums (see k_dp_smpf for the argument)
// ---------------------------------------------------------------------------------------------
#ifndef RC_SMP_FOLDS
#define RC_SMP_FOLDS 1  // 0: the round-1 fold on the quotients (fold_entry) in k_dp_smp / k_dp_smps
#endif
struct RowFoldS {
  float Ms;   // largest positive species sum so far (0: none yet); +inf: complex row, the state is in the shared-memory record
  float lo;   // max(Ms - B, 0): a later sum in (lo, Ms) may fall into the tie band of the maximum
  float nB;   // -B; -inf for a complex row (then lo stays 0 and every positive sum takes the exact path)
  int jF;     // end codon of Ms
};
__device__ __forceinline__ void folds_init(RowFoldS& f, float B) {
  f.Ms = 0.0f;
  f.lo = 0.0f;
  f.nB = -B;
  f.jF = 0;
}
// sB = s + nB (computed for both rows of the lane with one packed add)
__device__ __forceinline__ void folds_fast(RowFoldS& f, float s, float sB, int j, bool& amb) {
  const bool p = s >= f.Ms;
  amb = amb || (s > f.lo && !p);
  f.Ms = fmaxf(f.Ms, s);
  f.lo = fmaxf(f.lo, sB);
  f.jF = p ? j : f.jF;
}
__device__ __forceinline__ float quot_nk(float s, float fNK, float rcpNK) {  // the correctly rounded s / (N-1), see k_dp
  const float q = s * rcpNK;
  return __fmaf_rn(__fmaf_rn(-fNK, q, s), rcpNK, q);
}
__device__ __noinline__ RowFoldS folds_exact(RowFoldS f, float s, int j, float fNK, float rcpNK, RowRec* rec, int slots) {
  if (!(s > 0.0f)) return f;  // getHSS only looks at positive entries (src/score.c:891)
  const float e = quot_nk(s, fNK, rcpNK);
  if (f.Ms != INFINITY) {
    if (s >= f.Ms) {  // what the fast path does
      f.Ms = s;
      f.jF = j;
      f.lo = fmaxf(f.lo, s + f.nB);
      return f;
    }
    const float Me = quot_nk(f.Ms, fNK, rcpNK);  // f.Ms > s > 0: something was accepted before
    if (e == Me) {  // equal after rounding: an exact tie, the later (longer) entry stands for it
      f.jF = j;
      return f;
    }
    if (e - Me >= -0.0001f) {  // accepted below the maximum (src/score.c:953-954 without the length test): the band matters now
      rec_init(rec);
      fold_flush(rec, Me, f.jF);
      hss_accept_rec(rec, e, j, slots);
      f.Ms = INFINITY;
      f.lo = 0.0f;
      f.nB = -INFINITY;
    }
    return f;
  }
  if (e - rec->vF >= -0.0001f) hss_accept_rec(rec, e, j, slots);
  return f;
}
// final record of a row: straight from the registers unless the row is complex
__device__ __forceinline__ void folds_store(const RowFoldS& f, RowRec* grec, const RowRec* srec, float fNK, float rcpNK) {
  uint4* g = reinterpret_cast<uint4*>(grec);
  if (f.Ms == INFINITY) {
    rec_copy(grec, srec);
  } else if (f.Ms > 0.0f) {
    const unsigned e = __float_as_uint(quot_nk(f.Ms, fNK, rcpNK));
    g[0] = make_uint4(e, e, e, 0u);                                             // Emax, vF, be[0], be[1]
    g[1] = make_uint4(0u, (unsigned)f.jF | (1u << 16), (unsigned)f.jF, 0u);     // be[2], jF | n << 16, bj[0] | bj[1] << 16, bj[2] | pad
  } else {
    const unsigned ninf = __float_as_uint(-INFINITY);
    g[0] = make_uint4(ninf, ninf, 0u, 0u);
    g[1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

#ifndef RC_FOLD_GATE
#define RC_FOLD_GATE 0  // 1: skip the fold of a step pair when no lane has a sum above its row's lower bound
#endif
// two end codons (j, j+1) of the lane's two rows; the rows' states before the pair are kept for the exact path
__device__ __forceinline__ void folds_pair(RowFoldS& fx, RowFoldS& fy, float2 sumA, float2 sumB, int j, float fNK, float rcpNK,
                                           RowRec* rec0, int band_slots) {
#if RC_FOLD_GATE
  if (!(fmaxf(sumA.x, sumB.x) > fx.lo || fmaxf(sumA.y, sumB.y) > fy.lo)) return;
#endif
  RowFoldS nx = fx, ny = fy;
  bool amb = false;
  const float2 nB = make_float2(fx.nB, fy.nB);
  const float2 aB = add2(sumA, nB), bB = add2(sumB, nB);
  folds_fast(nx, sumA.x, aB.x, j, amb);
  folds_fast(ny, sumA.y, aB.y, j, amb);
  folds_fast(nx, sumB.x, bB.x, j + 1, amb);
  folds_fast(ny, sumB.y, bB.y, j + 1, amb);
  if (amb) {  // a possible near tie after some row's maximum: the two rows again, entry by entry, exactly
    fx = folds_exact(fx, sumA.x, j, fNK, rcpNK, rec0, band_slots);
    fx = folds_exact(fx, sumB.x, j + 1, fNK, rcpNK, rec0, band_slots);
    fy = folds_exact(fy, sumA.y, j, fNK, rcpNK, rec0 + 1, band_slots);
    fy = folds_exact(fy, sumB.y, j + 1, fNK, rcpNK, rec0 + 1, band_slots);
  } else {
    fx = nx;
    fy = ny;
  }
}
// one end codon; ylive: the lane's second row has started (it starts one codon after the first)
__device__ __forceinline__ void folds_single(RowFoldS& fx, RowFoldS& fy, float2 sum, int j, bool ylive, float fNK, float rcpNK,
                                             RowRec* rec0, int band_slots) {
  RowFoldS nx = fx, ny = fy;
  bool amb = false;
  const float2 sB = add2(sum, make_float2(fx.nB, fy.nB));
  folds_fast(nx, sum.x, sB.x, j, amb);
  if (ylive) folds_fast(ny, sum.y, sB.y, j, amb);
  if (amb) {
    fx = folds_exact(fx, sum.x, j, fNK, rcpNK, rec0, band_slots);
    if (ylive) fy = folds_exact(fy, sum.y, j, fNK, rcpNK, rec0 + 1, band_slots);
  } else {
    fx = nx;
    fy = ny;
  }
}

// ---------------------------------------------------------------------------------------------
// (c) k_dp_smp: the DP for SHORT blocks with many null alignments (N-1 <= 16, sites*RSB*128 B fits in shared
// memory).  Lane = one alignment instance (native or null sample), warp = 32 instances of the same block,
// strand and frame, working on the same pair of start codons: the gap pattern and hence z, the row starts and
// all control flow are identical across the lanes (samples inherit the native gaps, src/misc.c:127-148), so
// there is no start-of-row (diagonal) waste at all -- that waste dominates k_dp_reg on blocks of ~40 codons.
// The CTA's sigma table (all end codons of the frame for its 32 instances, lane-interleaved so that each lane
// reads its own values with conflict-free LDS.128) is brought in by one TMA bulk copy; the CTA's warps then
// share it and walk the start-codon pairs (long and short rows paired up for balance).  State handling,
// packed FADD2 arithmetic, float order and the getHSS digest are those of k_dp_reg.
// ---------------------------------------------------------------------------------------------
template <int NK>
__device__ __forceinline__ void smp_load_row(unsigned a, unsigned zw, float (&sv)[RegCfg<NK>::RS]) {
  constexpr int RSB = (NK + 3) / 4 * 4;
#pragma unroll
  for (int q = 0; q < RSB / 4; q++)
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(sv[4 * q]), "=f"(sv[4 * q + 1]), "=f"(sv[4 * q + 2]), "=f"(sv[4 * q + 3])
                 : "r"(a + 512 * q));
  sv[NK] = __uint_as_float(zw);
}

// CHAINED (layout 5, wide alignments): the species are cut into chunks of 1-3 quads; chunk g is one
// launch of this kernel with NK = 4 * (quads of the chunk) and continues, for every (start-codon pair, end codon,
// instance), the k-ordered partial species sum the launch of chunk g-1 left in global memory (`partial`, float2 per
// lane = the two rows of the pair).  The last chunk owns the getHSS digest.  Species past the end of the alignment in
// the last quad are dummies (sigma = +0, z = 0): with omega <= 0 they add max3(0, t*omega, t*omega) = +0.
#ifndef RC_SMP_DIAG_PAIR
#define RC_SMP_DIAG_PAIR 0  // measured: no gain on 40-codon frames (4.54 vs 4.56 ms), 10-20 registers more
#endif
// CHAINED: two CTAs of 8 warps per SM need at most 128 registers per thread (the 12-species chunk with partial sums coming in
// and the fold state would take 138)
// CHAINED: the partial sums a lane continues come from global memory (they were written by the launch of the preceding
// chunk), one float2 per end codon.  Waiting for them was the kernel's largest stall (long_scoreboard 4.1 per issue,
// profiles/r02_k_dp_smp_chunked.md), so every lane requests them SMP_PF end codons ahead with cp.async into a ring slot of
// its own in shared memory (no registers held while the copy is in flight): one commit group per end codon.
#ifndef RC_SMP_PF
#define RC_SMP_PF 4
#endif
constexpr int SMP_PF = RC_SMP_PF;                                   // end codons in flight per lane
constexpr int SMP_PF_BYTES = SMP_MAX_WARPS * SMP_PF * 256;  // ring of a CTA: [warp][slot][lane] float2
// LASTC (chained launches only): the launch of the last chunk, the one that owns the getHSS fold.  A compile-time flag: the
// launches of the other chunks carry no fold state, which is what keeps the 12-species chunk within the 128 registers that
// two resident CTAs allow (the kernel is very sensitive to spills in its step loop: one more live register cost 7 %).
template <int NK, bool CHAINED, bool LASTC = true>
__global__ void __launch_bounds__(SMP_MAX_WARPS * 32, (CHAINED || NK <= 12) ? 2 : 1)
    k_dp_smp(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const CtaDesc* __restrict__ ctas,
             const float* __restrict__ sigma, const unsigned* __restrict__ ztiles, RowRec* __restrict__ recs, Params prm,
             int band_slots, int chunk, float2* __restrict__ partial) {
  constexpr int RS = RegCfg<NK>::RS;
  constexpr int RSB = (NK + 3) / 4 * 4;
  constexpr int ROW_BYTES = (CHAINED ? 12 : RSB) * 32 * 4;  // one end codon, 32 lanes (chained: always room for three quads)
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;  // 4-8 warps share the CTA's sigma table
  const CtaDesc cd = ctas[blockIdx.x];
  const Item& it = items[cd.item];
  const BlockDev& bd = blocks[it.block];
  const int strand = cd.sf / 3, frame = cd.sf % 3;
  const int sites = bd.sites[frame];
  const int group = cd.task0;  // group of 32 instances inside the item
  // Folded group: a group with at most 16 instances (the last one of a block with 101 = 3 * 32 + 5 of them, RNAcode's default
  // -n 100) does not leave its other lanes idle: the lanes are cut into R = 32 / m replicas of m >= #instances lanes, lane l
  // works for instance l % m (it reads that instance's column of the sigma table), and replica l / m takes its own start-codon
  // pair -- R pairs side by side.  The rows of the replicas start 2 codons apart; until the last one has started the steps are
  // masked per lane (reg_update_diag).
  const int ninst_g = min(32, it.ninst - group * 32);
  int fold_m = 32;
  if (bd.smp_fold && ninst_g <= 16) {
    fold_m = 1;
    while (fold_m < ninst_g) fold_m <<= 1;
  }
  const int R = 32 / fold_m, il = lane & (fold_m - 1), rep = lane / fold_m;
  const int inst_l = group * 32 + il;
  const bool valid = il < ninst_g;
  const bool first = !CHAINED || chunk == 0;
  constexpr bool last = !CHAINED || LASTC;
  const int ngrp = (it.ninst + 31) / 32;

  const size_t sig_bytes = (size_t)sites * ROW_BYTES;
  const size_t z_bytes = ((size_t)sites * 4 + 15) / 16 * 16;
  unsigned* zs = reinterpret_cast<unsigned*>(smem + sig_bytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sig_bytes + z_bytes);
  // records of rows whose fold went complex, two per lane and warp (only the launch that owns the digest has room for them)
  RowRec* srec = reinterpret_cast<RowRec*>(smem + sig_bytes + z_bytes + 16 + (CHAINED ? SMP_PF_BYTES : 0));
  // CHAINED: this lane's slot 0 of the prefetch ring for incoming partial sums
  unsigned pf_a = smem_u32(smem + sig_bytes + z_bytes + 16) + (unsigned)(warp * (SMP_PF * 256) + lane * 8);
  asm volatile("" : "+r"(pf_a));
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, (unsigned)(sig_bytes + z_bytes));
    const size_t grp_index = CHAINED ? (size_t)chunk * ngrp + group : (size_t)group;
    bulk_g2s(smem, sigma + it.sigma_off[strand][frame] + grp_index * sites * (ROW_BYTES / 4), (unsigned)sig_bytes, bar);
    bulk_g2s(zs, ztiles + bd.z_off[strand][frame] + (CHAINED ? (size_t)chunk * bd.ntiles[frame] * TILE : 0), (unsigned)z_bytes, bar);
  }
  __syncthreads();
  mbar_wait(bar, 0);

  unsigned sig_a = smem_u32(smem) + il * 16;
  asm volatile("" : "+r"(sig_a));
  const float Delta = prm.Delta, Omega = prm.Omega;
  float omega = prm.omega;
  asm volatile("" : "+f"(omega));
  const float fNK = bd.fNK, rcpNK = bd.rcpNK, foldB = bd.fold_B;
  RowRec* rec_inst = recs + it.rec_off[strand][frame] + (size_t)(valid ? inst_l : 0) * sites;
  // partial sums of this (item, strand, frame, group): [pair][end codon from the pair's first row on][lane]
  const int npairs = (sites + 1) / 2;
  float2* part = nullptr;
  if (CHAINED) {
    const size_t per_group = ((size_t)npairs * sites - (size_t)npairs * (npairs - 1)) * 32;
    part = partial + it.part_off[strand][frame] + (size_t)group * per_group + lane;
  }

  // boustrophedon assignment of start-codon pairs (R at a time) to warps: w, 2W-1-w, 2W+w, 4W-1-w, ... (rows get shorter
  // with the pair index, so every warp receives a similar total length)
  const int nsuper = (npairs + R - 1) / R;
#pragma unroll 1
  for (int turn = 0;; turn++) {
    const int sp = (turn & 1) ? (turn + 1) * nw - 1 - warp : turn * nw + warp;
    if (turn * nw >= nsuper) break;
    if (sp >= nsuper) continue;
    const int p = sp * R + rep;            // this lane's pair (may lie past the end for the last replicas)
    const int r0 = 2 * p;                  // its first row
    const int r_first = 2 * sp * R;        // first row of the warp
    const int j_steady = r_first + 2 * R;  // from here on both rows of every lane have started
    const bool live = p < npairs;
    RowRec* rec0 = srec + (warp * 32 + lane) * 2;
    // pp[j * 32] = entry of end codon j of this lane's pair (a lane without a pair never touches it)
    float2* pp = (CHAINED && live) ? part + ((size_t)p * sites - (size_t)p * (p - 1) - r0) * 32 : nullptr;
    float2 S0[NK], S1[NK], S2[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) S0[k] = S1[k] = S2[k] = make_float2(0.0f, 0.0f);
    RowFoldS sx, sy;
    folds_init(sx, foldB);
    folds_init(sy, foldB);
    int j = r_first;
    const bool pf = CHAINED && !first && live;  // this lane continues partial sums of the preceding chunk
    int jreq = r0;                              // next end codon whose partial sum this lane has to request
    // One request = one commit group; a request past the last end codon re-reads the last entry into a slot nobody reads, so
    // that "at most two groups pending" always means "the entries of j and j + 1 have arrived".
    auto pf_issue = [&](int jtop) {
      if (jreq < jtop + SMP_PF) {
        cp_async8(pf_a + (unsigned)(jreq & (SMP_PF - 1)) * 256u, pp + (size_t)min(jreq, sites - 1) * 32);
        cp_async_commit();
        jreq++;
      }
    };
    if (pf) {
      pf_issue(j);
      pf_issue(j);
    }
#pragma unroll 1
    while (j < sites) {
      if (pf) {
        // top-up: an iteration consumes one or two end codons, two requests keep SMP_PF of them in flight
        pf_issue(j);
        pf_issue(j);
        cp_async_wait<SMP_PF - 2>();
      }
      float svA[RS];
      const unsigned zA = zs[j];
      smp_load_row<NK>(sig_a + j * ROW_BYTES, zA, svA);
      if (R > 1 && j < j_steady) {
        // folded group, rows still starting: one end codon at a time, every addend masked per lane until its row starts
        // (the state stays exactly (0,0,0), the sums 0, and the fold ignores a sum of 0)
        const bool mx = j >= r0, my = j >= r0 + 1;
        float2 sin = make_float2(0.0f, 0.0f);
        if (pf && mx) sin = lds_f2(pf_a + (unsigned)(j & (SMP_PF - 1)) * 256u);
        if (CHAINED && !first && !my) sin.y = 0.0f;
        const float2 sum = reg_update_diag<NK, CHAINED>(S0, S1, S2, svA, mx, my, Delta, Omega, omega, sin);
        if (CHAINED && !last) {
          if (mx) pp[(size_t)j * 32] = sum;
        } else {
          folds_single(sx, sy, sum, j, true, fNK, rcpNK, rec0, band_slots);
        }
        j += 1;
        continue;
      }
      if (j >= j_steady && j + 1 < sites) {
        const unsigned zB = zs[j + 1];
        if ((zA | zB) == 0u) {
          float svB[RS];
          smp_load_row<NK>(sig_a + (j + 1) * ROW_BYTES, zB, svB);
          float2 sumA, sumB, sinA = make_float2(0.0f, 0.0f), sinB = make_float2(0.0f, 0.0f);
          if (pf) {
            sinA = lds_f2(pf_a + (unsigned)(j & (SMP_PF - 1)) * 256u);
            sinB = lds_f2(pf_a + (unsigned)((j + 1) & (SMP_PF - 1)) * 256u);
          }
          reg_pair_fast<NK, CHAINED>(S0, S1, S2, svA, svB, omega, sumA, sumB, sinA, sinB);
          if (CHAINED && !last) {
            if (live) {
              pp[(size_t)j * 32] = sumA;
              pp[(size_t)(j + 1) * 32] = sumB;
            }
          } else {
            folds_pair(sx, sy, sumA, sumB, j, fNK, rcpNK, rec0, band_slots);
          }
          j += 2;
          continue;
        }
      }
      float2 sin = make_float2(0.0f, 0.0f);
      if (pf) sin = lds_f2(pf_a + (unsigned)(j & (SMP_PF - 1)) * 256u);
      const float2 sum = reg_update<NK, CHAINED>(S0, S1, S2, svA, j < r0 + 2, j, r0, Delta, Omega, omega, sin);
      if (CHAINED && !last) {
        if (live) pp[(size_t)j * 32] = sum;
      } else {
        folds_single(sx, sy, sum, j, j > r0, fNK, rcpNK, rec0, band_slots);
      }
      j += 1;
    }
    // the requests past the last end codon are still in flight: they must have landed before the next row pair's requests
    // reuse their slots (two asynchronous copies to one address are not ordered among themselves)
    if (pf) cp_async_wait<0>();
    if (valid && last && r0 < sites) {
      folds_store(sx, rec_inst + r0, rec0, fNK, rcpNK);
      if (r0 + 1 < sites) folds_store(sy, rec_inst + r0 + 1, rec0 + 1, fNK, rcpNK);
    }
  }
}

// k_dp_smps: the same for frames whose sigma table does not fit into shared memory.  The table is streamed through
// shared memory in segments of SMP_SEG end codons (two stages, TMA bulk
// copies): the CTA's warps take consecutive start-codon pairs, one pair each (a "batch"), and walk the segments from
// the batch's first row to the end of the frame together, so every segment is fetched once per batch and the kernel
// has no limit on the frame length.
template <int NK, bool CHAINED>
__global__ void __launch_bounds__(SMP_MAX_WARPS * 32)
    k_dp_smps(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const CtaDesc* __restrict__ ctas,
             const float* __restrict__ sigma, const unsigned* __restrict__ ztiles, RowRec* __restrict__ recs, Params prm,
             int band_slots, int chunk, float2* __restrict__ partial) {
  constexpr int RS = RegCfg<NK>::RS;
  constexpr int RSB = (NK + 3) / 4 * 4;
  constexpr int ROW_BYTES = (CHAINED ? 12 : RSB) * 32 * 4;  // one end codon, 32 lanes (chained: always room for three quads)
  constexpr int T = SMP_SEG;
  constexpr int STAGE_BYTES = T * ROW_BYTES;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;  // 4-8 warps share the CTA's sigma stream
  const CtaDesc cd = ctas[blockIdx.x];
  const Item& it = items[cd.item];
  const BlockDev& bd = blocks[it.block];
  const int strand = cd.sf / 3, frame = cd.sf % 3;
  const int sites = bd.sites[frame];
  const int group = cd.task0;  // group of 32 instances inside the item
  const int inst_l = group * 32 + lane;
  const bool valid = inst_l < it.ninst;
  const bool first = !CHAINED || chunk == 0, last = !CHAINED || chunk == bd.nchunk - 1;
  const int ngrp = (it.ninst + 31) / 32;

  const size_t z_bytes = ((size_t)sites * 4 + 15) / 16 * 16;
  unsigned* zs = reinterpret_cast<unsigned*>(smem + 2 * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGE_BYTES + z_bytes);  // [0], [1]: stages; [2]: z words
  // fold state of the getHSS digest, two records per lane and warp (only the launch that owns the digest has room for it)
  RowRec* srec = reinterpret_cast<RowRec*>(smem + 2 * STAGE_BYTES + z_bytes + 32);
  const size_t grp_index = CHAINED ? (size_t)chunk * ngrp + group : (size_t)group;
  const float* sig_src = sigma + it.sigma_off[strand][frame] + grp_index * sites * (ROW_BYTES / 4);
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_fence_init();
    mbar_expect_tx(&bars[2], (unsigned)z_bytes);
    bulk_g2s(zs, ztiles + bd.z_off[strand][frame] + (CHAINED ? (size_t)chunk * bd.ntiles[frame] * TILE : 0), (unsigned)z_bytes, &bars[2]);
  }
  __syncthreads();
  mbar_wait(&bars[2], 0);

  const unsigned stage_a = smem_u32(smem) + lane * 16;
  const float Delta = prm.Delta, Omega = prm.Omega;
  float omega = prm.omega;
  asm volatile("" : "+f"(omega));
  const float fNK = bd.fNK, rcpNK = bd.rcpNK;
  RowRec* rec_inst = recs + it.rec_off[strand][frame] + (size_t)(valid ? inst_l : 0) * sites;
  // partial sums of this (item, strand, frame, group): [pair][end codon from the pair's first row on][lane]
  const int npairs = (sites + 1) / 2;
  const int nseg = (sites + T - 1) / T;
  float2* part = nullptr;
  if (CHAINED) {
    const size_t per_group = ((size_t)npairs * sites - (size_t)npairs * (npairs - 1)) * 32;
    part = partial + it.part_off[strand][frame] + (size_t)group * per_group + lane;
  }
  unsigned it_count = 0;  // segments consumed so far by the CTA (stage = it_count & 1, phase = it_count >> 1)

#pragma unroll 1
  for (int p0 = 0; p0 < npairs; p0 += nw) {
    const int p = p0 + warp;
    const bool active = p < npairs;
    const int r0 = 2 * p;
    const int seg0 = (2 * p0) / T;  // the batch's first row lies in this segment
    if (threadIdx.x == 0) {
      for (int q = 0; q < 2 && seg0 + q < nseg; q++) {
        const unsigned sq = (it_count + q) & 1u;
        const unsigned bytes = (unsigned)(min(T, sites - (seg0 + q) * T) * ROW_BYTES);
        mbar_expect_tx(&bars[sq], bytes);
        bulk_g2s(smem + sq * STAGE_BYTES, sig_src + (size_t)(seg0 + q) * T * (ROW_BYTES / 4), bytes, &bars[sq]);
      }
    }
    RowRec* rec0 = srec + (warp * 32 + lane) * 2;
#if !RC_SMP_FOLDS
    if (last && active) {
      rec_init(rec0);
      rec_init(rec0 + 1);
    }
#endif
    float2* pp = CHAINED ? part + ((size_t)p * sites - (size_t)p * (p - 1) - r0) * 32 : nullptr;  // pp[j * 32] = entry of end codon j
    float2 S0[NK], S1[NK], S2[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) S0[k] = S1[k] = S2[k] = make_float2(0.0f, 0.0f);
#if RC_SMP_FOLDS
    RowFoldS sx, sy;
    folds_init(sx, bd.fold_B);
    folds_init(sy, bd.fold_B);
#else
    RowFold fx, fy;
    fold_init(fx);
    fold_init(fy);
#endif
    int j = r0;
#pragma unroll 1
    for (int seg = seg0; seg < nseg; seg++, it_count++) {
      const unsigned st = it_count & 1u;
      mbar_wait(&bars[st], (it_count >> 1) & 1u);
      const int jend = min(sites, (seg + 1) * T);
      const unsigned row_a = stage_a + st * STAGE_BYTES - (unsigned)(seg * T) * ROW_BYTES;  // row_a + j * ROW_BYTES = row of end codon j
      if (active) {
#pragma unroll 1
        while (j < jend) {
          float svA[RS];
          const unsigned zA = zs[j];
          smp_load_row<NK>(row_a + j * ROW_BYTES, zA, svA);
          if (j >= r0 + 2 && j + 1 < jend) {
            const unsigned zB = zs[j + 1];
            if ((zA | zB) == 0u) {
              float svB[RS];
              smp_load_row<NK>(row_a + (j + 1) * ROW_BYTES, zB, svB);
              float2 sumA, sumB, sinA = make_float2(0.0f, 0.0f), sinB = make_float2(0.0f, 0.0f);
              if (CHAINED && !first) {
                sinA = pp[(size_t)j * 32];
                sinB = pp[(size_t)(j + 1) * 32];
              }
              reg_pair_fast<NK, CHAINED>(S0, S1, S2, svA, svB, omega, sumA, sumB, sinA, sinB);
              if (CHAINED && !last) {
                pp[(size_t)j * 32] = sumA;
                pp[(size_t)(j + 1) * 32] = sumB;
              }
#if RC_SMP_FOLDS
              else {
                folds_pair(sx, sy, sumA, sumB, j, fNK, rcpNK, rec0, band_slots);
              }
#else
              else if (fmaxf(fmaxf(sumA.x, sumA.y), fmaxf(sumB.x, sumB.y)) > 0.0f && valid) {
                fold_entry(sumA.x, j, r0, sites, fNK, rcpNK, rec0, band_slots, fx);
                fold_entry(sumA.y, j, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, fy);
                fold_entry(sumB.x, j + 1, r0, sites, fNK, rcpNK, rec0, band_slots, fx);
                fold_entry(sumB.y, j + 1, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, fy);
              }
#endif
              j += 2;
              continue;
            }
          }
          float2 sin = make_float2(0.0f, 0.0f);
          if (CHAINED && !first) sin = pp[(size_t)j * 32];
          const float2 sum = reg_update<NK, CHAINED>(S0, S1, S2, svA, j < r0 + 2, j, r0, Delta, Omega, omega, sin);
          if (CHAINED && !last) {
            pp[(size_t)j * 32] = sum;
          }
#if RC_SMP_FOLDS
          else {
            folds_single(sx, sy, sum, j, j > r0, fNK, rcpNK, rec0, band_slots);
          }
#else
          else if (fmaxf(sum.x, sum.y) > 0.0f && valid) {
            fold_entry(sum.x, j, r0, sites, fNK, rcpNK, rec0, band_slots, fx);
            if (r0 + 1 < sites) fold_entry(sum.y, j, r0 + 1, sites, fNK, rcpNK, rec0 + 1, band_slots, fy);
          }
#endif
          j += 1;
        }
      }
      __syncthreads();  // every warp is done with this stage
      if (threadIdx.x == 0 && seg + 2 < nseg) {
        const unsigned bytes = (unsigned)(min(T, sites - (seg + 2) * T) * ROW_BYTES);
        mbar_expect_tx(&bars[st], bytes);
        bulk_g2s(smem + st * STAGE_BYTES, sig_src + (size_t)(seg + 2) * T * (ROW_BYTES / 4), bytes, &bars[st]);
      }
    }
#if RC_SMP_FOLDS
    if (active && valid && last) {
      folds_store(sx, rec_inst + r0, rec0, fNK, rcpNK);
      if (r0 + 1 < sites) folds_store(sy, rec_inst + r0 + 1, rec0 + 1, fNK, rcpNK);
    }
#else
    if (active && valid && last) {
#pragma unroll
      for (int t = 0; t < 2; t++)
        if (r0 + t < sites) {
          const RowFold& f = t ? fy : fx;
          if (f.jF >= 0) fold_flush(rec0 + t, f.M, f.jF);
          rec_copy(rec_inst + r0 + t, rec0 + t);
        }
    }
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// (b)+(c) k_dp_smpf: k_dp_smp with kernel (b) fused in.  The CTA (32 instances of one block, one strand, one frame) builds
// its sigma table itself: the class bytes of its 32 reference rows and of four species rows at a time are staged in shared
// memory with aligned word copies (lane = instance, odd row pitch => conflict-free byte reads), every (end codon, instance)
// gets its four sigma values from the two look-ups of PairTables and stores them as one float4 into the lane-interleaved
// table the DP loop reads -- sigma never exists in HBM (k_sigma_smp wrote and k_dp_smp re-read 0.4 KB per codon and
// instance).  While one CTA of an SM is in its table phase (LSU / integer work) the other runs its DP phase (FP32 pipe).
//
// The getHSS fold (see k_dp) runs in the space of the species sums s, not of the entries e = s / (N-1):
//   * e is a monotone function of s, so the row maximum and exact ties can be tracked on s without the division;
//   * the record of a row only has to keep (a) the row maximum Emax with the LAST end codon that reaches it, and (b) entries
//     AFTER that position which the lenient fold accepts, i.e. which lie within the 1e-4 band below the last accepted value.
//     Band entries BEFORE the maximum are redundant for the replay in k_hss: they are shorter and not larger than the
//     maximum, so whenever one of them passes the tie rule (|e - cur| <= 1e-4 and length >= current length, src/score.c:953-959)
//     the maximum passes it too;
//   * hence the state of almost every row is just (Ms, jF) in registers: a sum s >= Ms replaces it; a sum below Ms - B is
//     rejected for certain, where B = (N-1) * 1.0002e-4 + 2^-21 * (largest possible sum) covers the tolerance and every
//     rounding of the two quotients; only a positive sum in (Ms - B, Ms) -- a near tie after the maximum -- needs the exact
//     test on the quotients.  That rare case runs the reference's rule on the quotients (folds_exact) and from then on the row
//     is "complex": its state lives in a shared-memory RowRec handled by hss_accept_rec, as in k_dp_smp.
// The fast path is branch-free: per entry two compares, two max, one select.
// ---------------------------------------------------------------------------------------------
template <int NK>
struct SmpfCfg {
  static constexpr int RSB = (NK + 3) / 4 * 4;
  static __host__ __device__ size_t align16(size_t v) { return (v + 15) / 16 * 16; }
  // dynamic shared memory: sigma table | z words | barrier | PairTables (packed-codon order) | expected scores of the staged
  // species | flag words of the staged rows | codon columns (byte-wise path) | packed reference row | packed species rows |
  // fold records.  nsp = species rows staged at once (all of the launch's quads)
  static __host__ __device__ size_t off_z(int sites, int row_bytes) { return (size_t)sites * row_bytes; }
  static __host__ __device__ size_t off_bar(int sites, int row_bytes) { return off_z(sites, row_bytes) + align16((size_t)sites * 4); }
  static __host__ __device__ size_t off_tab(int sites, int row_bytes) { return off_bar(sites, row_bytes) + 16; }
  static __host__ __device__ size_t off_sc(int sites, int row_bytes) { return off_tab(sites, row_bytes) + sizeof(PairTables); }
  static __host__ __device__ size_t off_flag(int sites, int row_bytes, int nsp) { return off_sc(sites, row_bytes) + align16((size_t)nsp * 16); }
  static __host__ __device__ size_t off_col(int sites, int row_bytes, int nsp) { return off_flag(sites, row_bytes, nsp) + align16((size_t)(nsp + 1) * 4); }
  static __host__ __device__ size_t off_ref(int sites, int row_bytes, int nsp) { return off_col(sites, row_bytes, nsp) + align16((size_t)3 * sites * 4); }
  static __host__ __device__ size_t off_sp(int sites, int row_bytes, int nsp, int words) { return off_ref(sites, row_bytes, nsp) + (size_t)(words + 1) * 128; }
  static __host__ __device__ size_t off_rec(int sites, int row_bytes, int nsp, int words) {
    return off_sp(sites, row_bytes, nsp, words) + ((size_t)nsp * words + 1) * 128;
  }
  static __host__ __device__ size_t total(int sites, int row_bytes, int nsp, int words, int nw) {
    return off_rec(sites, row_bytes, nsp, words) + (size_t)nw * 64 * sizeof(RowRec);
  }
};

// six adjacent bits of a packed row (see k_pack2): the codon whose first position sits at bit `sh` of the word at shared
// address `a`; two: the codon runs over into the next word (128 bytes on).  sh and two are warp-uniform.
__device__ __forceinline__ unsigned p2_codon(unsigned a, unsigned sh, bool two) {
  const unsigned lo = lds_u32(a);
  if (!two) return (lo >> sh) & 63u;
  return __funnelshift_r(lo, lds_u32(a + 128), sh) & 63u;
}

template <int NK, bool CHAINED>
__global__ void __launch_bounds__(SMP_MAX_WARPS * 32, 2)  // two CTAs of 8 warps per SM: at most 128 registers
    k_dp_smpf(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const CtaDesc* __restrict__ ctas,
              const unsigned* __restrict__ p2, const unsigned* __restrict__ p2f, const unsigned char* __restrict__ cls,
              const int* __restrict__ cols0, const float* __restrict__ scores, const PairTables* __restrict__ tables,
              const unsigned* __restrict__ ztiles, RowRec* __restrict__ recs, Params prm, int band_slots, int chunk,
              float2* __restrict__ partial) {
  constexpr int RS = RegCfg<NK>::RS;
  constexpr int RSB = (NK + 3) / 4 * 4;
  constexpr int ROW_BYTES = (CHAINED ? 12 : RSB) * 32 * 4;  // one end codon, 32 lanes (chained: always room for three quads)
  using Cfg = SmpfCfg<NK>;
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const CtaDesc cd = ctas[blockIdx.x];
  const Item& it = items[cd.item];
  const BlockDev& bd = blocks[it.block];
  // cd.sf = strand * 3 + frame, or 6 + strand: ONE CTA for the three frames of the strand -- they share the tables and the packed
  // rows, only the z words, the sigma table and the DP differ (a third of the CTAs, a third of the set-up latency)
  const bool allf = !CHAINED && cd.sf >= 6;
  const int strand = allf ? cd.sf - 6 : cd.sf / 3;
  const int f_lo = allf ? 0 : cd.sf % 3, f_hi = allf ? 3 : f_lo + 1;
  const int sites_c = bd.sites[f_lo];  // the carve-up of shared memory follows the CTA's first (= longest) frame
  const int group = cd.task0;  // group of 32 instances inside the item
  // Folded group: a group with at most 16 instances (the last one of a block with 101 = 3 * 32 + 5 of them, RNAcode's default
  // -n 100) does not leave its other lanes idle: the lanes are cut into R = 32 / m replicas of m >= #instances lanes, lane l
  // works for instance l % m, and replica l / m takes its own start-codon pair -- R pairs side by side.  The rows of the
  // replicas start 2 codons apart; until the last one has started the steps are masked per lane (reg_update_diag).
  const int ninst_g = min(32, it.ninst - group * 32);
  int fold_m = 32;
  if (!CHAINED && bd.smp_fold && ninst_g <= 16) {
    fold_m = 1;
    while (fold_m < ninst_g) fold_m <<= 1;
  }
  const int R = 32 / fold_m, il = lane & (fold_m - 1), rep = lane / fold_m;
  const int inst_l = group * 32 + il;
  const bool valid = il < ninst_g;
  const bool first = !CHAINED || chunk == 0, last = !CHAINED || chunk == bd.nchunk - 1;
  const int N = bd.N, cols = bd.cols, L = bd.L, W = bd.p2_words;

  // species of this launch: all of them, or the quads of the chunk (layout 5)
  int q_first = 0, q_count = RSB / 4;
  if (CHAINED) {
    q_first = chunk * bd.chunk_base + min(chunk, bd.chunk_rem);
    q_count = bd.chunk_base + (chunk < bd.chunk_rem ? 1 : 0);
  }
  const int NSP = CHAINED ? 12 : RSB;                      // species slots the carve-up provides for
  const int k_first = 4 * q_first;                         // first scored species (0-based) of the launch
  const int n_real = min(4 * q_count, bd.NK - k_first);    // species rows that exist

  unsigned* zs = reinterpret_cast<unsigned*>(smem + Cfg::off_z(sites_c, ROW_BYTES));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Cfg::off_bar(sites_c, ROW_BYTES));
  const PairTables& s_tab = *reinterpret_cast<const PairTables*>(smem + Cfg::off_tab(sites_c, ROW_BYTES));
  float* s_sc = reinterpret_cast<float*>(smem + Cfg::off_sc(sites_c, ROW_BYTES));            // [species of the launch][h], h = 0 -> 0
  unsigned* s_flag = reinterpret_cast<unsigned*>(smem + Cfg::off_flag(sites_c, ROW_BYTES, NSP));  // [0]: reference row, [1 + k]: species
  int* s_col = reinterpret_cast<int*>(smem + Cfg::off_col(sites_c, ROW_BYTES, NSP));      // columns of the frame's codons (byte-wise path)
  unsigned char* s_ref = smem + Cfg::off_ref(sites_c, ROW_BYTES, NSP);                    // packed reference row [word][lane]
  unsigned char* s_sp = smem + Cfg::off_sp(sites_c, ROW_BYTES, NSP, W);                   // packed species rows [species][word][lane]
  RowRec* srec = reinterpret_cast<RowRec*>(smem + Cfg::off_rec(sites_c, ROW_BYTES, NSP, W));

  // ---- table phase: kernel (b) for this CTA's (instances, strand, frame, species of the launch) ------------------------
  const size_t prow = (size_t)W * 128;  // bytes of one packed row of 32 instances
  const size_t grp = (size_t)(it.inst0 >> 5) + group;
  const unsigned* gp2 = p2 + bd.p2_off + ((grp * 2 + strand) * N) * (size_t)W * 32;
  const unsigned* gfl = p2f + bd.p2f_off + (grp * 2 + strand) * N;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, (unsigned)(sizeof(PairTables) + prow + (size_t)n_real * prow));
    bulk_g2s(smem + Cfg::off_tab(sites_c, ROW_BYTES), tables, (unsigned)sizeof(PairTables), bar);
    bulk_g2s(s_ref, gp2, (unsigned)prow, bar);
    bulk_g2s(s_sp, gp2 + (size_t)(1 + k_first) * W * 32, (unsigned)((size_t)n_real * prow), bar);
  }
  // expected scores of the launch's species on this strand, [species][h], h = 0 -> 0 (see PairTables); flag words of the rows
  for (int t = threadIdx.x; t < 4 * NSP; t += blockDim.x) {
    const int kk = t >> 2, h = t & 3, row = 1 + k_first + kk;
    s_sc[t] = (h > 0 && kk < n_real) ? scores[bd.scores_off + ((size_t)strand * N + row) * 4 + h] : 0.0f;
  }
  for (int t = threadIdx.x; t <= NSP; t += blockDim.x) s_flag[t] = t == 0 ? gfl[0] : (t - 1 < n_real ? gfl[k_first + t] : 0u);
#pragma unroll 1
  for (int frame = f_lo; frame < f_hi; frame++) {
  const int sites = bd.sites[frame];
  if (sites <= 0) continue;  // (uniform)
  // z words of the frame (one per end codon; a few hundred bytes: plain loads, in flight while the bulk copies arrive)
  const unsigned* zg = ztiles + bd.z_off[strand][frame] + (CHAINED ? (size_t)chunk * bd.ntiles[frame] * TILE : 0);
  for (int t = threadIdx.x; t < sites; t += blockDim.x) zs[t] = zg[t];
  // codon of site j: reference positions x-2 .. x with x = 3j + 3 + frame, i.e. entries frame+1+3j .. frame+3+3j of cols0
  const int* c0 = cols0 + bd.cols0_off + (size_t)strand * (L + 1) + frame + 1;
  for (int t = threadIdx.x; t < 3 * sites; t += blockDim.x) s_col[t] = c0[t];
  __syncthreads();  // barrier initialised; s_sc, s_flag, zs, s_col written
  if (frame == f_lo) mbar_wait(bar, 0);
  {
    const unsigned lane_bit = 1u << il;
    // class bytes of this lane's instance, for codons of rows with 'N' / 'X' (rare)
    const unsigned char* cbase = cls + bd.cls_off + (size_t)(it.inst0 + (valid ? inst_l : 0)) * bd.inst_stride;
    const unsigned ref_a = smem_u32(s_ref) + il * 4, sp_a = smem_u32(s_sp) + il * 4, prow32 = (unsigned)prow;
    const unsigned tab_t = smem_u32(&s_tab.t[0]), tab_v = smem_u32(&s_tab.val[0]), sc_a = smem_u32(s_sc);
    const unsigned out_a = smem_u32(smem) + lane * 16;
    unsigned fl_all = s_flag[0];  // flag words of the reference row and of all species rows of the launch
    for (int k = 0; k < n_real; k++) fl_all |= s_flag[1 + k];
    const bool slow_warp = fl_all != 0u;  // warp-uniform (every lane reads the same words)
    constexpr int NQ = CHAINED ? 3 : RSB / 4;
    for (int j = warp; j < sites; j += nw) {
      const int p0 = 3 * j + frame;
      const unsigned woff = (unsigned)(p0 >> 4) * 128u, sh = 2u * (unsigned)(p0 & 15);
      const bool two = sh > 26u;
      const unsigned qa = p2_codon(ref_a + woff, sh, two);
      const unsigned trow = tab_t + (qa << 7);  // 64 entries of 2 bytes per reference codon
      unsigned ka = sp_a + woff;                // the codon's word in species row 0 of the launch, then row by row
#pragma unroll
      for (int kq = 0; kq < NQ; kq++) {
        if (CHAINED && kq >= q_count) break;
        float v4[4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int k = 4 * kq + kk;
          float v = 0.0f;
          if (CHAINED ? (k < n_real) : (k < NK)) {  // compile-time for the plain kernel, warp-uniform otherwise
            const unsigned qb = p2_codon(ka, sh, two);
            const unsigned e = lds_u16(trow + 2u * qb);
            // observed - expected (src/score.c:422-425), or constant - 0
            v = lds_f32(tab_v + 4u * (e & 0x3ffu)) - lds_f32(sc_a + 16u * (unsigned)k + 4u * (e >> 10));
            ka += prow32;
          }
          v4[kk] = v;
        }
        if (slow_warp) {  // some lane's rows hold an 'N' or 'X': those lanes test the six characters (src/score.c:394-404)
          unsigned fl_q = s_flag[0];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) fl_q |= s_flag[1 + 4 * kq + kk];
          if ((fl_q & lane_bit) != 0u && valid) {
            const int i1 = s_col[3 * j], i2 = s_col[3 * j + 1], i3 = s_col[3 * j + 2];
            const unsigned a1 = cbase[i1], a2 = cbase[i2], a3 = cbase[i3];
            const unsigned nA = (a1 | a2 | a3) & CLS_N;
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
              if (4 * kq + kk < n_real) {
                const unsigned char* rk = cbase + (size_t)(1 + k_first + 4 * kq + kk) * cols;
                const unsigned b1 = rk[i1], b2 = rk[i2], b3 = rk[i3];
                if (nA | ((b1 | b2 | b3) & CLS_N) | (b1 & b2 & b3 & CLS_X)) v4[kk] = 0.0f;
              }
          }
        }
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(out_a + (unsigned)j * ROW_BYTES + kq * 512), "f"(v4[0]),
                     "f"(v4[1]), "f"(v4[2]), "f"(v4[3])
                     : "memory");
      }
    }
  }
  __syncthreads();  // the table is complete

  // ---- DP phase: k_dp_smp's loop with the fold in species-sum space ------------------------------------------------------
  unsigned sig_a = smem_u32(smem) + lane * 16, zs_a = smem_u32(zs);
  asm volatile("" : "+r"(sig_a), "+r"(zs_a));
  const float Delta = prm.Delta, Omega = prm.Omega;
  float omega = prm.omega;
  asm volatile("" : "+f"(omega));
  const float fNK = bd.fNK, rcpNK = bd.rcpNK, foldB = bd.fold_B;
  RowRec* rec_inst = recs + it.rec_off[strand][frame] + (size_t)(valid ? inst_l : 0) * sites;
  const int npairs = (sites + 1) / 2;
  float2* part = nullptr;
  if (CHAINED) {
    const size_t per_group = ((size_t)npairs * sites - (size_t)npairs * (npairs - 1)) * 32;
    part = partial + it.part_off[strand][frame] + (size_t)group * per_group + lane;
  }
  const int nsuper = (npairs + R - 1) / R;  // start-codon pairs taken R at a time (R = 1: one pair per warp and turn)
#pragma unroll 1
  for (int turn = 0;; turn++) {
    const int sp = (turn & 1) ? (turn + 1) * nw - 1 - warp : turn * nw + warp;  // boustrophedon, see k_dp_smp
    if (turn * nw >= nsuper) break;
    if (sp >= nsuper) continue;
    const int p = sp * R + rep;            // this lane's pair (may lie past the end for the last replicas)
    const int r0 = 2 * p;                  // its first row
    const int r_first = 2 * sp * R;        // first row of the warp
    const int j_steady = r_first + 2 * R;  // from here on both rows of every lane have started
    RowRec* rec0 = srec + (warp * 32 + lane) * 2;
    float2* pp = CHAINED ? part + ((size_t)p * sites - (size_t)p * (p - 1) - r0) * 32 : nullptr;
    float2 S0[NK], S1[NK], S2[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) S0[k] = S1[k] = S2[k] = make_float2(0.0f, 0.0f);
    RowFoldS fx, fy;
    folds_init(fx, foldB);
    folds_init(fy, foldB);
    int j = r_first;
#pragma unroll 1
    while (j < sites) {
      float svA[RS];
      const unsigned zA = lds_u32(zs_a + 4u * (unsigned)j);
      smp_load_row<NK>(sig_a + j * ROW_BYTES, zA, svA);
      if (R > 1 && j < j_steady) {
        // folded group, rows still starting: one end codon at a time, every addend masked per lane until its row starts
        // (the state stays exactly (0,0,0), the sums 0, and the fold ignores a sum of 0)
        const float2 sum = reg_update_diag<NK>(S0, S1, S2, svA, j >= r0, j >= r0 + 1, Delta, Omega, omega);
        folds_single(fx, fy, sum, j, true, fNK, rcpNK, rec0, band_slots);
        j += 1;
        continue;
      }
      if (j >= j_steady && j + 1 < sites) {
        const unsigned zB = lds_u32(zs_a + 4u * (unsigned)j + 4u);
        if ((zA | zB) == 0u) {
          float svB[RS];
          smp_load_row<NK>(sig_a + (j + 1) * ROW_BYTES, zB, svB);
          float2 sumA, sumB, sinA = make_float2(0.0f, 0.0f), sinB = make_float2(0.0f, 0.0f);
          if (CHAINED && !first) {
            sinA = pp[(size_t)j * 32];
            sinB = pp[(size_t)(j + 1) * 32];
          }
          reg_pair_fast<NK, CHAINED>(S0, S1, S2, svA, svB, omega, sumA, sumB, sinA, sinB);
          if (CHAINED && !last) {
            pp[(size_t)j * 32] = sumA;
            pp[(size_t)(j + 1) * 32] = sumB;
          } else {
            folds_pair(fx, fy, sumA, sumB, j, fNK, rcpNK, rec0, band_slots);
          }
          j += 2;
          continue;
        }
      }
      float2 sin = make_float2(0.0f, 0.0f);
      if (CHAINED && !first) sin = pp[(size_t)j * 32];
      const float2 sum = reg_update<NK, CHAINED>(S0, S1, S2, svA, j < r0 + 2, j, r0, Delta, Omega, omega, sin);
      if (CHAINED && !last) {
        pp[(size_t)j * 32] = sum;
      } else {
        folds_single(fx, fy, sum, j, j > r0, fNK, rcpNK, rec0, band_slots);
      }
      j += 1;
    }
    if (valid && last && r0 < sites) {
      folds_store(fx, rec_inst + r0, rec0, fNK, rcpNK);
      if (r0 + 1 < sites) folds_store(fy, rec_inst + r0 + 1, rec0 + 1, fNK, rcpNK);
    }
  }
  __syncthreads();  // every warp is done with the frame's table and z words before the next frame overwrites them
  }
}

// ---------------------------------------------------------------------------------------------
// (b) k_sigma_p2: the sigma tables of the sample-major layouts 2 / 5 that live in HBM (k_dp_smp, k_dp_smps), built from
// k_pack2's packed rows exactly as k_dp_smpf's table phase builds its own -- a codon is six adjacent bits of a packed word,
// sigma = val[t[codonA][codonB] & 0x3ff] - expected[k][t >> 10] (calculateSigma, src/score.c:375-426, through PairTables), rows
// with an 'N' / 'X' at a reference position take the byte-wise test per lane -- instead of k_sigma_smp's three byte loads and
// shifts per character (43 instructions per sigma value there, about 10 here).
// One CTA per (item, group of 32 instances, strand, species chunk): the packed reference row and the chunk's species rows arrive
// by TMA bulk copies and serve all three frames; lane = instance, a warp writes the 512 bytes of one (end codon, species quad)
// with one float4 store per lane.  grid = (item, group * 2 + strand, chunk).
// Dynamic shared memory: PairTables | expected scores [16][4] | flag words [17] | barrier | reference row | species rows.
// ---------------------------------------------------------------------------------------------
struct SigP2Cfg {
  static __host__ __device__ size_t off_sc() { return sizeof(PairTables); }
  static __host__ __device__ size_t off_flag() { return off_sc() + 16 * 4 * sizeof(float); }
  static __host__ __device__ size_t off_bar() { return off_flag() + 32 * sizeof(unsigned); }
  static __host__ __device__ size_t off_ref() { return off_bar() + 16; }
  static __host__ __device__ size_t off_sp(int words) { return off_ref() + (size_t)(words + 1) * 128; }
  static __host__ __device__ size_t total(int words, int nsp) { return off_sp(words) + ((size_t)nsp * words + 1) * 128; }
};

__global__ void __launch_bounds__(256)
    k_sigma_p2(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const unsigned* __restrict__ p2,
               const unsigned* __restrict__ p2f, const unsigned char* __restrict__ cls, const int* __restrict__ cols0,
               const float* __restrict__ scores, const PairTables* __restrict__ tables, float* __restrict__ sigma) {
  extern __shared__ __align__(128) unsigned char smem[];
  const Item& it = items[blockIdx.x];
  const BlockDev& bd = blocks[it.block];
  if (!bd.sig_p2) return;
  const int group = blockIdx.y >> 1, strand = blockIdx.y & 1, chunk = blockIdx.z;
  const bool chained = bd.layout == 5;
  if (group * 32 >= it.ninst || chunk >= (chained ? bd.nchunk : 1)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int N = bd.N, cols = bd.cols, L = bd.L, W = bd.p2_words, NK = bd.NK;
  const int rsb = (NK + 3) / 4 * 4;
  int q_first = 0, q_count = rsb / 4;
  if (chained) {
    q_first = chunk * bd.chunk_base + min(chunk, bd.chunk_rem);
    q_count = bd.chunk_base + (chunk < bd.chunk_rem ? 1 : 0);
  }
  const int k_first = 4 * q_first;
  const int n_real = min(4 * q_count, NK - k_first);         // species rows of this CTA that exist
  const unsigned ROW_BYTES = (chained ? 12u : (unsigned)rsb) * 128u;  // one end codon of the table: quads x 32 lanes x 16 B
  const int ninst_g = min(32, it.ninst - group * 32);
  const bool valid = lane < ninst_g;
  const int ngrp = (it.ninst + 31) / 32;
  const size_t grp_index = chained ? (size_t)chunk * ngrp + group : (size_t)group;

  const PairTables& s_tab = *reinterpret_cast<const PairTables*>(smem);
  float* s_sc = reinterpret_cast<float*>(smem + SigP2Cfg::off_sc());
  unsigned* s_flag = reinterpret_cast<unsigned*>(smem + SigP2Cfg::off_flag());
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SigP2Cfg::off_bar());
  unsigned char* s_ref = smem + SigP2Cfg::off_ref();
  unsigned char* s_sp = smem + SigP2Cfg::off_sp(W);

  const size_t prow = (size_t)W * 128;  // bytes of one packed row of 32 instances
  const size_t grp = (size_t)(it.inst0 >> 5) + group;
  const unsigned* gp2 = p2 + bd.p2_off + ((grp * 2 + strand) * N) * (size_t)W * 32;
  const unsigned* gfl = p2f + bd.p2f_off + (grp * 2 + strand) * N;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, (unsigned)(sizeof(PairTables) + prow + (size_t)n_real * prow));
    bulk_g2s(smem, tables, (unsigned)sizeof(PairTables), bar);
    bulk_g2s(s_ref, gp2, (unsigned)prow, bar);
    bulk_g2s(s_sp, gp2 + (size_t)(1 + k_first) * W * 32, (unsigned)((size_t)n_real * prow), bar);
  }
  for (int t = threadIdx.x; t < 64; t += blockDim.x) {
    const int kk = t >> 2, h = t & 3, row = 1 + k_first + kk;
    s_sc[t] = (h > 0 && kk < n_real) ? scores[bd.scores_off + ((size_t)strand * N + row) * 4 + h] : 0.0f;
  }
  for (int t = threadIdx.x; t <= 16; t += blockDim.x) s_flag[t] = t == 0 ? gfl[0] : (t - 1 < n_real ? gfl[k_first + t] : 0u);
  __syncthreads();  // barrier initialised; s_sc, s_flag written
  mbar_wait(bar, 0);

  const unsigned lane_bit = 1u << lane;
  const unsigned char* cbase = cls + bd.cls_off + (size_t)(it.inst0 + group * 32 + (valid ? lane : 0)) * bd.inst_stride;
  const unsigned ref_a = smem_u32(s_ref) + lane * 4, sp_a = smem_u32(s_sp) + lane * 4, prow32 = (unsigned)prow;
  const unsigned tab_t = smem_u32(&s_tab.t[0]), tab_v = smem_u32(&s_tab.val[0]), sc_a = smem_u32(s_sc);
  unsigned fl_all = s_flag[0];
  for (int k = 0; k < n_real; k++) fl_all |= s_flag[1 + k];
  const bool slow_warp = fl_all != 0u;  // warp-uniform
  // codon of site j of frame f: reference positions x-2 .. x with x = 3j + 3 + f, i.e. entries f+1+3j .. f+3+3j of cols0
  const int* c0s = cols0 + bd.cols0_off + (size_t)strand * (L + 1);
#pragma unroll 1
  for (int frame = 0; frame < 3; frame++) {
    const int sites = bd.sites[frame];
    float* outg = sigma + it.sigma_off[strand][frame] + grp_index * sites * (ROW_BYTES / 4) + lane * 4;
#pragma unroll 1
    for (int j = warp; j < sites; j += nw) {
      const int p0 = 3 * j + frame;
      const unsigned woff = (unsigned)(p0 >> 4) * 128u, sh = 2u * (unsigned)(p0 & 15);
      const bool two = sh > 26u;
      const unsigned qa = p2_codon(ref_a + woff, sh, two);
      const unsigned trow = tab_t + (qa << 7);  // 64 entries of 2 bytes per reference codon
      unsigned ka = sp_a + woff;                // the codon's word in species row 0 of the CTA, then row by row
#pragma unroll 1
      for (int kq = 0; kq < q_count; kq++) {
        float v4[4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          const int k = 4 * kq + kk;
          float v = 0.0f;
          if (k < n_real) {  // warp-uniform
            const unsigned qb = p2_codon(ka, sh, two);
            const unsigned e = lds_u16(trow + 2u * qb);
            v = lds_f32(tab_v + 4u * (e & 0x3ffu)) - lds_f32(sc_a + 16u * (unsigned)k + 4u * (e >> 10));  // src/score.c:422-425
            ka += prow32;
          }
          v4[kk] = v;
        }
        if (slow_warp) {  // some lane's rows hold an 'N' or 'X': those lanes test the six characters (src/score.c:394-404)
          unsigned fl_q = s_flag[0];
#pragma unroll
          for (int kk = 0; kk < 4; kk++) fl_q |= s_flag[1 + 4 * kq + kk];
          if ((fl_q & lane_bit) != 0u && valid) {
            const int i1 = c0s[frame + 1 + 3 * j], i2 = c0s[frame + 2 + 3 * j], i3 = c0s[frame + 3 + 3 * j];
            const unsigned a1 = cbase[i1], a2 = cbase[i2], a3 = cbase[i3];
            const unsigned nA = (a1 | a2 | a3) & CLS_N;
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
              if (4 * kq + kk < n_real) {
                const unsigned char* rk = cbase + (size_t)(1 + k_first + 4 * kq + kk) * cols;
                const unsigned b1 = rk[i1], b2 = rk[i2], b3 = rk[i3];
                if (nA | ((b1 | b2 | b3) & CLS_N) | (b1 & b2 & b3 & CLS_X)) v4[kk] = 0.0f;
              }
          }
        }
        *reinterpret_cast<float4*>(outg + ((size_t)j * ROW_BYTES + (size_t)kq * 512) / 4) = make_float4(v4[0], v4[1], v4[2], v4[3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (c) k_hss: the sequential scan of getHSS (src/score.c:888-961) over row digests.
// One WARP per (instance, strand, frame): the lanes fetch 32 row records at a time (coalesced 1 KB), rows without
// a positive entry are skipped by ballot, the others are replayed in row order with their fields broadcast by
// shuffles (every lane runs the same state machine; lane 0 writes).  grid = (x: item, y over ninst*6 warps).
// ---------------------------------------------------------------------------------------------
constexpr int HSS_WARPS = 4;
constexpr int HSS_WARP_MIN_SITES = 96;  // shorter frames: one thread per (instance, strand, frame) is the better fit (k_hss_thr)
constexpr int HSS_THR_MAX_SITES = 640;  // ... and frames up to this length when the block has thousands of scans (see hss_thr_tasks)

__global__ void __launch_bounds__(HSS_WARPS * 32)
    k_hss(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const RowRec* __restrict__ recs,
          float* __restrict__ res, HssDev* __restrict__ hss, int* __restrict__ hsscnt, int* __restrict__ ovf_counter) {
  const Item it = items[blockIdx.x];
  const BlockDev& bd = blocks[it.block];
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.y * HSS_WARPS + (threadIdx.x >> 5);
  if (idx >= it.ninst * 6 || !bd.hss_warp) return;
  const int inst_l = idx / 6, sf = idx % 6;
  const int strand = sf / 3, frame = sf % 3;
  const int sites = bd.sites[frame];
  const int inst = it.inst0 + inst_l;
  const uint4* rp = reinterpret_cast<const uint4*>(recs + it.rec_off[strand][frame] + (size_t)inst_l * sites);
  HssDev* out = hss + bd.hss_off[strand][frame];
  const unsigned FULL = 0xffffffffu;
  int nout = 0;
  float best = -1.0f;
  bool overflow = false;
  float cur = 0.0f;
  int segS = -1, segE = -1;
  const int nrows = sites - 1;  // the last row only holds the frame's final entry, which never survives (:893-900)
  uint4 n0 = make_uint4(0u, 0u, 0u, 0u), n1 = make_uint4(0u, 0u, 0u, 0u);  // records of the next 32 rows, fetched ahead
  if (lane < nrows) {
    n0 = rp[2 * lane];
    n1 = rp[2 * lane + 1];
  }
  for (int i0 = 0; i0 < nrows; i0 += 32) {
    const uint4 w0 = n0, w1 = n1;
    n0 = make_uint4(0u, 0u, 0u, 0u);
    n1 = make_uint4(0u, 0u, 0u, 0u);
    if (i0 + 32 + lane < nrows) {
      n0 = rp[2 * (i0 + 32 + lane)];
      n1 = rp[2 * (i0 + 32 + lane) + 1];
    }
    // RowRec: w0 = {Emax, vF, be0, be1}, w1 = {be2, jF | n << 16, bj0 | bj1 << 16, bj2 | pad << 16}
    // Rows that cannot change the scan's state are skipped 32 at a time: a row matters only if it has entries and either
    // lies beyond the current segment (flush, :897-949) or its maximum reaches cur - 1e-4 (a larger score or a tie, :953-959;
    // band entries never exceed the row maximum).  The state only changes at processed rows, so the test with the current
    // (cur, segE) is the test each skipped row would have seen.
    const bool has = (w1.y >> 16) != 0u;
    if (__any_sync(FULL, has && ((w1.y >> 16) & 0x8000u))) overflow = true;
    const float myE = __uint_as_float(w0.x);
    unsigned pend = __ballot_sync(FULL, has);
    while (pend) {
      const bool cand = has && ((cur > 0.0f && segE < i0 + lane) || !(myE - cur < -0.0001f));
      const unsigned mask = __ballot_sync(FULL, cand) & pend;
      if (!mask) break;
      const int src = __ffs(mask) - 1;
      pend &= ~((2u << src) - 1u);  // src and the rows before it are decided
      const int i = i0 + src;
      const float Emax = __uint_as_float(__shfl_sync(FULL, w0.x, src));
      const float vF = __uint_as_float(__shfl_sync(FULL, w0.y, src));
      const unsigned jn = __shfl_sync(FULL, w1.y, src);
      const int n = (int)(jn >> 16), jF = (int)(jn & 0xffffu);
      bool take;
      if (cur > 0.0f && segE < i) {  // flush (:897-949)
        if (segE - segS >= 2) {
          if (inst == 0 && lane == 0) {
            out[nout].startSite = segS;
            out[nout].endSite = segE;
            out[nout].score = cur;
          }
          nout++;
          best = fmaxf(best, cur);
        }
        take = true;
      } else {  // overlap with the current segment (:953-959)
        take = Emax > cur;
        if (!take) {
          const int nb = n & 0xff;
          const float be0 = __uint_as_float(__shfl_sync(FULL, w0.z, src)), be1 = __uint_as_float(__shfl_sync(FULL, w0.w, src)),
                      be2 = __uint_as_float(__shfl_sync(FULL, w1.x, src));
          const unsigned bj01 = __shfl_sync(FULL, w1.z, src), bj2 = __shfl_sync(FULL, w1.w, src);
          const float be[3] = {be0, be1, be2};
          const int bj[3] = {(int)(bj01 & 0xffffu), (int)(bj01 >> 16), (int)(bj2 & 0xffffu)};
#pragma unroll
          for (int m = 0; m < REC_SLOTS; m++) {
            const float d = be[m] - cur;
            if (m < nb && d >= -0.0001f && d <= 0.0001f && (bj[m] - i) >= (segE - segS)) take = true;
          }
        }
      }
      if (take) {
        cur = vF;
        segS = i;
        segE = jF;
      }
    }
  }
  if (sites > 0 && segE - segS >= 2) {  // forced flush on the frame's last entry
    if (inst == 0 && lane == 0) {
      out[nout].startSite = segS;
      out[nout].endSite = segE;
      out[nout].score = cur;
    }
    nout++;
    best = fmaxf(best, cur);
  }
  if (lane == 0) {
    if (overflow) {
      best = -2.0f;
      atomicAdd(ovf_counter, 1);
    }
    res[bd.res_off + (size_t)inst * 6 + sf] = best;
    if (inst == 0) hsscnt[bd.hsscnt_off + sf] = overflow ? -1 : nout;
  }
}

// Same scan, one thread per (instance, strand, frame): for short frames (many small blocks).
__global__ void __launch_bounds__(128)
    k_hss_thr(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const RowRec* __restrict__ recs,
          float* __restrict__ res, HssDev* __restrict__ hss, int* __restrict__ hsscnt, int* __restrict__ ovf_counter) {
  const Item it = items[blockIdx.x];
  const BlockDev& bd = blocks[it.block];
  const int idx = blockIdx.y * blockDim.x + threadIdx.x;
  if (idx >= it.ninst * 6 || bd.hss_warp) return;
  const int inst_l = idx / 6, sf = idx % 6;
  const int strand = sf / 3, frame = sf % 3;
  const int sites = bd.sites[frame];
  const int inst = it.inst0 + inst_l;
  const uint4* rp = reinterpret_cast<const uint4*>(recs + it.rec_off[strand][frame] + (size_t)inst_l * sites);
  HssDev* out = hss + bd.hss_off[strand][frame];
  int nout = 0;
  float best = -1.0f;
  bool overflow = false;
  float cur = 0.0f;
  int segS = -1, segE = -1;
  for (int i = 0; i + 1 < sites; i++) {  // the last row only holds the frame's final entry, which never survives (:893-900)
    const uint4 w0 = rp[2 * i], w1 = rp[2 * i + 1];
    RowRec r;
    *reinterpret_cast<uint4*>(&r) = w0;
    *(reinterpret_cast<uint4*>(&r) + 1) = w1;
    if (r.n == 0) continue;  // no positive entry in this row
    if (r.n & 0x8000) overflow = true;
    bool take;
    if (cur > 0.0f && segE < i) {  // flush (:897-949)
      if (segE - segS >= 2) {
        if (inst == 0) {
          out[nout].startSite = segS;
          out[nout].endSite = segE;
          out[nout].score = cur;
        }
        nout++;
        best = fmaxf(best, cur);
      }
      take = true;
    } else {  // overlap with the current segment (:953-959)
      take = r.Emax > cur;
      const int nb = r.n & 0xff;
      for (int m = 0; m < nb && !take; m++) {
        const float d = r.be[m] - cur;
        if (d >= -0.0001f && d <= 0.0001f && ((int)r.bj[m] - i) >= (segE - segS)) take = true;
      }
    }
    if (take) {
      cur = r.vF;
      segS = i;
      segE = r.jF;
    }
  }
  if (sites > 0 && segE - segS >= 2) {  // forced flush on the frame's last entry
    if (inst == 0) {
      out[nout].startSite = segS;
      out[nout].endSite = segE;
      out[nout].score = cur;
    }
    nout++;
    best = fmaxf(best, cur);
  }
  if (overflow) {
    best = -2.0f;
    atomicAdd(ovf_counter, 1);
  }
  res[bd.res_off + (size_t)inst * 6 + sf] = best;
  if (inst == 0) hsscnt[bd.hsscnt_off + sf] = overflow ? -1 : nout;
}

// ---------------------------------------------------------------------------------------------
// Exact fallback: getHSS on the materialised S entries of one frame.  One warp per
// (instance, strand, frame); lanes fetch 32 entries at a time, only positive ones are replayed
// (identically on every lane).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    k_hss_dense(const BlockDev* __restrict__ blocks, const Item* __restrict__ items, const float* __restrict__ dense,
                float* __restrict__ res, HssDev* __restrict__ hss, int* __restrict__ hsscnt) {
  const Item it = items[blockIdx.x];
  const BlockDev& bd = blocks[it.block];
  const int widx = (blockIdx.y * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (widx >= it.ninst * 6) return;
  const int inst_l = widx / 6, sf = widx % 6;
  const int strand = sf / 3, frame = sf % 3;
  const int sites = bd.sites[frame];
  const int inst = it.inst0 + inst_l;
  const float* S = dense + it.dense_off[strand][frame] + (size_t)inst_l * ((size_t)sites * (sites + 1) / 2);
  HssDev* out = hss + bd.hss_off[strand][frame];
  int nout = 0;
  float best = -1.0f, cur = 0.0f;
  int segS = -1, segE = -1;
  for (int i = 0; i < sites; i++) {
    const float* row = S + ((long long)i * sites - (long long)i * (i - 1) / 2) - i;
    for (int jb = i; jb < sites; jb += 32) {
      const int j = jb + lane;
      const float v = (j < sites) ? row[j] : 0.0f;
      const bool last = (i == sites - 1 && j == sites - 1);
      unsigned mask = __ballot_sync(0xffffffffu, (j < sites) && (v > 0.0f || last));
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float e = __shfl_sync(0xffffffffu, v, src);
        const int jj = jb + src;
        const bool lst = (i == sites - 1 && jj == sites - 1);
        if ((cur > 0.0f && segE < i) || lst) {
          if (segE - segS >= 2) {
            if (inst == 0 && lane == 0) {
              out[nout].startSite = segS;
              out[nout].endSite = segE;
              out[nout].score = cur;
            }
            nout++;
            best = fmaxf(best, cur);
          }
          cur = e;
          segS = i;
          segE = jj;
        } else {
          const float d = e - cur;
          if (e > cur || (d >= -0.0001f && d <= 0.0001f && (jj - i) >= (segE - segS))) {
            cur = e;
            segS = i;
            segE = jj;
          }
        }
      }
    }
  }
  if (lane == 0) {
    res[bd.res_off + (size_t)inst * 6 + sf] = best;
    if (inst == 0) hsscnt[bd.hsscnt_off + sf] = nout;
  }
}

// ---------------------------------------------------------------------------------------------
// (d) k_evolve: null alignments on the GPU, replacing simulateTree / tree2aln / sortAln on the host
// (src/treeSimulate.c:52-97, :254-283, src/misc.c:150-171; seq-gen HKY, no rate heterogeneity:
// seqgen/evolve.c:167-199, :291-308, :400-433).
//
// rng 0 (exact): every (block, sample) has a bit-exact MT19937 (seqgen/twister.c:73-146), seeded with the sample's
// seed and consumed in seq-gen's order: nodes in evolution order, one genrand_real1 per site.  A warp owns EVO_SPW
// consecutive samples: their generators are seeded side by side (init_genrand is a serial chain of 624 steps, lane q
// runs it for sample q), then the warp draws the samples one after the other with all 32 lanes.
// The 624-word state lives in shared memory and is regenerated by the 32 lanes in place (each lane reads its
// three source words before any lane writes; the recurrence only reaches words of earlier batches or old words
// of later ones).  SetState's `r > P[j]` on doubles (r = u * (1/4294967295.0)) is evaluated as `u > thr[j]`
// with integer thresholds computed on the host from the caller's cumulative matrices with the very same
// double expression, so the drawn states are identical.  State 4 (r above the last cumulative entry, undefined
// behaviour in the reference) is clamped to 3.
// rng 1: Philox4x32-10 keyed by the sample seed, counter = (node, site): same distribution, any order.
//
// Internal-node sequences go to a scratch area (one byte per site), tips are written as "ACGT" characters
// straight into the raw sample rows (input order), which k_pack then treats like host-provided samples.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned mt_temper(unsigned y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

// need: how many words of the new batch will ever be read (the sample's last batch is only regenerated that far: word kk of the
// new batch depends on the OLD words kk, kk+1, kk+397 or the new word kk-227 -- never on a later new word).
// (Measured and rejected: three chunks of 224 independent words, seven per lane, two barriers per chunk instead of two per 32
// words -- 10.6 against 9.1 ms of pack + evolve on 10 000 blocks of 10 x 120.)
#ifndef RC_EVO_TWIST_ROWS
#define RC_EVO_TWIST_ROWS 2  // rows of 32 words per turn of the regeneration (measured: 1 -> 8.37, 2 -> 8.14, 3 -> 8.16, 7 -> 10.6 ms of pack + evolve on 10 000 x 10x120)
#endif
__device__ __forceinline__ void mt_twist(unsigned* mt, int lane, int need = 624) {
#if RC_EVO_TWIST_ROWS > 1
  // U rows of 32 words per turn (U * 32 < 227 consecutive words are independent of each other): one pair of barriers per turn
  constexpr int U = RC_EVO_TWIST_ROWS;
  const unsigned ma = smem_u32(mt);
  const int lim = min(624, (need + 31) & ~31);
#pragma unroll 1
  for (int base = 0; base < lim; base += 32 * U) {
    unsigned a[U], b[U], c[U];
#pragma unroll
    for (int i = 0; i < U; i++) {
      const int kk = base + 32 * i + lane;
      if (kk < lim) {
        a[i] = lds_u32(ma + 4u * (unsigned)kk);
        b[i] = lds_u32(ma + 4u * (unsigned)(kk == 623 ? 0 : kk + 1));
        c[i] = lds_u32(ma + 4u * (unsigned)(kk < 227 ? kk + 397 : kk - 227));
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < U; i++) {
      const int kk = base + 32 * i + lane;
      if (kk < lim) {
        const unsigned y = (a[i] & 0x80000000u) | (b[i] & 0x7fffffffu);
        const unsigned v = c[i] ^ (y >> 1) ^ ((0u - (b[i] & 1u)) & 0x9908b0dfu);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ma + 4u * (unsigned)kk), "r"(v) : "memory");
      }
    }
    __syncwarp();
  }
#else
#pragma unroll 1
  for (int base = 0; base < need; base += 32) {
    const int kk = base + lane;
    unsigned a = 0, b = 0, c = 0;
    if (kk < 624) {
      a = mt[kk];
      b = mt[kk + 1 < 624 ? kk + 1 : 0];
      c = mt[kk + 397 < 624 ? kk + 397 : kk - 227];
    }
    __syncwarp();
    if (kk < 624) {
      const unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
      mt[kk] = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    __syncwarp();
  }
#endif
}

__device__ __forceinline__ void philox_round(unsigned (&c)[4], unsigned k0, unsigned k1) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0;
  c[1] = lo1;
  c[2] = n2;
  c[3] = lo0;
}
__device__ __forceinline__ unsigned philox_draw(unsigned seed, unsigned node, unsigned site) {
  unsigned c[4] = {site, node, 0x52434F44u, 0u};  // "RCOD"
  unsigned k0 = seed, k1 = 0x9E3779B9u ^ seed;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c[0];
}

constexpr int EVO_WARPS = 4;
#ifndef RC_EVO_SPW
#define RC_EVO_SPW 8
#endif
constexpr int EVO_SPW = RC_EVO_SPW;  // most samples per task: their MT19937 states are seeded side by side, one lane each (the host
                                     // takes fewer per task when a batch has too few samples to fill the GPU with tasks of eight)

// The next `cnt` (<= 128) outputs of the sample's generator, four consecutive ones per lane (lane l: outputs 4l .. 4l+3 of the
// pass); the state is regenerated when the pass crosses the end of the current 624-word batch (warp-uniform).
// left: outputs the sample still needs from the batches not yet generated.
__device__ __forceinline__ void evo_draw4(unsigned* mt, int& pos, int& left, int cnt, int lane, unsigned (&u)[4]) {
  const int off = 4 * lane, idx = pos + off;
  unsigned v[4];
#pragma unroll
  for (int t = 0; t < 4; t++) v[t] = (off + t < cnt && idx + t < 624) ? mt[idx + t] : 0u;
  if (pos + cnt > 624) {  // the pass crosses a batch boundary (warp-uniform)
    __syncwarp();
    mt_twist(mt, lane, min(624, left));
    left -= 624;
#pragma unroll
    for (int t = 0; t < 4; t++)
      if (off + t < cnt && idx + t >= 624) v[t] = mt[idx + t - 624];
    pos -= 624;
  }
  pos += cnt;
#pragma unroll
  for (int t = 0; t < 4; t++) u[t] = mt_temper(v[t]);
}

// One null alignment of a SMALL problem: at most 32 tree nodes and 128 * P columns (the production regime: blocks of a dozen
// species cut to ~200 columns).  A site's states along the tree never leave the lane that owns the site: the states of all
// nodes of a site are two bits each in one 64-bit register (a child looks its parent's up with a shift), the thresholds of
// ALL nodes sit in shared memory for the whole task, and nothing is exchanged between lanes except inside the generator --
// no sequence of an internal node is written anywhere, no per-node table reload, no per-node warp barrier.  The generator's
// outputs are consumed exactly as in the general path: node by node in evolution order, one per site.
// s_nd[n] = (parent + 1) | (row + 1) << 8; s_th[n * 4 + parent state] = the node's three thresholds (+ unused fourth).
constexpr int EVO_SMALL_NODES = 32;
template <int P>
__device__ __forceinline__ void evo_sample_small(int rng, int n_nodes, int cols, unsigned* mt, const unsigned* s_nd,
                                                 const uint4* s_th, unsigned seed, unsigned char* myraw, int lane) {
  unsigned long long sb[P][4];
#pragma unroll
  for (int p = 0; p < P; p++)
#pragma unroll
    for (int t = 0; t < 4; t++) sb[p][t] = 0ull;
  int pos = 624, left = n_nodes * cols;
  const int off = 4 * lane;
#pragma unroll 1
  for (int n = 0; n < n_nodes; n++) {
    const unsigned ndw = s_nd[n];
    const int parent = (int)(ndw & 0xffu) - 1, row = (int)(ndw >> 8) - 1;
    unsigned char* orow = row >= 0 ? myraw + (size_t)row * cols : nullptr;
    const bool row_aligned = (reinterpret_cast<size_t>(orow) & 3) == 0;  // warp-uniform
    const unsigned psh = parent >= 0 ? 2u * (unsigned)parent : 0u;
    const uint4* tn = s_th + 4 * n;
#pragma unroll
    for (int p = 0; p < P; p++) {
      const int c0 = 128 * p;
      if (P > 1 && c0 >= cols) break;
      const int cnt = min(128, cols - c0), site0 = c0 + off;
      unsigned u[4];
      if (rng == 0) {
        evo_draw4(mt, pos, left, cnt, lane, u);
      } else {
#pragma unroll
        for (int t = 0; t < 4; t++) u[t] = philox_draw(seed, (unsigned)n, (unsigned)(site0 + t));
      }
      if (off < cnt) {
        unsigned ch = 0u;
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const unsigned ps = parent >= 0 ? (unsigned)(sb[p][t] >> psh) & 3u : 0u;  // the root draws from row 0 of its table
          const uint4 tt = tn[ps];
          const unsigned state = (u[t] > tt.x) + (u[t] > tt.y) + (u[t] > tt.z);
          sb[p][t] |= (unsigned long long)state << (2 * n);
          ch |= ((0x54474341u >> (8 * state)) & 0xffu) << (8 * t);  // "ACGT"[state]
        }
        if (orow) {
          if (row_aligned && off + 3 < cnt) {
            *reinterpret_cast<unsigned*>(orow + site0) = ch;
          } else {
#pragma unroll
            for (int t = 0; t < 4; t++)
              if (off + t < cnt) orow[site0 + t] = (unsigned char)(ch >> (8 * t));
          }
        }
      }
    }
  }
}

// Persistent warps: warp `wg` of the grid works through the tasks wg, wg + n_warps, ...; a task is EVO_SPW consecutive
// samples of one block (evo_task0: prefix sums of the tasks per block).  The states of a task are seeded by EVO_SPW lanes in
// parallel into the warp's private slot of a global scratch (init_genrand is a serial chain of 624 steps: seeding one state
// per warp wastes 31 lanes, seeding 32 per warp in shared memory would leave room for two warps per SM); each sample then
// loads its state into the warp's 2.5 KB of shared memory and is drawn with all 32 lanes.
__global__ void __launch_bounds__(EVO_WARPS * 32, 4)
    k_evolve(const BlockDev* __restrict__ blocks, const EvoDev* __restrict__ evos, const int* __restrict__ evo_task0, int n_evos,
             const int* __restrict__ nodes, const unsigned* __restrict__ thr, const unsigned* __restrict__ seeds,
             unsigned char* __restrict__ seqs, unsigned char* __restrict__ raw, unsigned* __restrict__ mt_scratch, int spw) {
  __shared__ unsigned s_mt[EVO_WARPS][624];
  __shared__ uint4 s_thr[EVO_WARPS][4];
  __shared__ uint4 s_tha[EVO_WARPS][EVO_SMALL_NODES * 4];  // small problems: thresholds of all nodes
  __shared__ unsigned s_nda[EVO_WARPS][EVO_SMALL_NODES];   // ... and their (parent, row)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = blockIdx.x * EVO_WARPS + warp, n_warps = gridDim.x * EVO_WARPS;
  const int total_tasks = evo_task0[n_evos];
  unsigned* mt = s_mt[warp];
  unsigned* slot = mt_scratch + (size_t)wg * spw * 624;
#pragma unroll 1
  for (int task = wg; task < total_tasks; task += n_warps) {
    int lo = 0, hi = n_evos - 1;  // the block of the task: largest e with evo_task0[e] <= task
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (evo_task0[mid] <= task) lo = mid;
      else hi = mid - 1;
    }
    const EvoDev ev = evos[lo];
    const BlockDev& bd = blocks[ev.block];
    const int n_samp = bd.n_inst - 1;
    const int s_base = (task - evo_task0[lo]) * spw;
    const int n_mine = min(spw, n_samp - s_base);
    const int cols = bd.cols;
    const int* nd = nodes + ev.node_off;
    const unsigned* th = thr + ev.thr_off;
    const bool small = ev.n_nodes <= EVO_SMALL_NODES && cols <= 256 && bd.N <= 255;  // warp-uniform (evo_sample_small)
    if (small) {
      __syncwarp();  // the previous task's samples are done with the tables
      for (int i = lane; i < ev.n_nodes * 16; i += 32) reinterpret_cast<unsigned*>(&s_tha[warp][0])[i] = th[i];
      if (lane < ev.n_nodes) s_nda[warp][lane] = (unsigned)(nd[4 * lane] + 1) | ((unsigned)(nd[4 * lane + 1] + 1) << 8);
      __syncwarp();
    }
    if (ev.rng == 0) {
      // init_genrand (seqgen/twister.c:73-86): lane q seeds the state of sample s_base + q
      if (lane < n_mine) {
        unsigned x = seeds[ev.seed_off + s_base + lane];
        unsigned* m = slot + (size_t)lane * 624;
        for (int i = 0; i < 624; i++) {
          m[i] = x;
          x = 1812433253u * (x ^ (x >> 30)) + (unsigned)(i + 1);
        }
      }
      __syncwarp();
    }
#pragma unroll 1
    for (int q = 0; q < n_mine; q++) {
      const int sample = s_base + q;
      const unsigned seed = seeds[ev.seed_off + sample];
      const int cols4 = (cols + 3) & ~3;  // internal-node sequences: one byte per site (the state 0..3), rows padded to words
      unsigned char* myseq = seqs + ev.seq_off + (size_t)sample * ev.n_internal * cols4;
      unsigned char* myraw = raw + bd.raw_off + (size_t)sample * bd.N * cols;
      if (ev.rng == 0) {
        for (int i = lane; i < 624; i += 32) mt[i] = slot[(size_t)q * 624 + i];
        __syncwarp();
      }
      if (small) {
        if (cols <= 128) evo_sample_small<1>(ev.rng, ev.n_nodes, cols, mt, s_nda[warp], s_tha[warp], seed, myraw, lane);
        else evo_sample_small<2>(ev.rng, ev.n_nodes, cols, mt, s_nda[warp], s_tha[warp], seed, myraw, lane);
        __syncwarp();  // the state in shared memory is free for the next sample
        continue;
      }
      int pos = 624;  // next unread word of the current 624-word batch
      int left = ev.n_nodes * cols;  // outputs still to come from batches not yet generated
      for (int n = 0; n < ev.n_nodes; n++) {
        const int parent = nd[4 * n], row = nd[4 * n + 1], slot_n = nd[4 * n + 2];
        const unsigned* pseq = parent >= 0 ? reinterpret_cast<const unsigned*>(myseq + (size_t)nd[4 * parent + 2] * cols4) : nullptr;
        unsigned* oseq = slot_n >= 0 ? reinterpret_cast<unsigned*>(myseq + (size_t)slot_n * cols4) : nullptr;
        unsigned char* orow = row >= 0 ? myraw + (size_t)row * cols : nullptr;
        const bool row_aligned = (reinterpret_cast<size_t>(orow) & 3) == 0;  // warp-uniform
        // the node's thresholds [parent state][4] in shared memory: one 16-byte load per site below
        if (lane < 16) reinterpret_cast<unsigned*>(&s_thr[warp][0])[lane] = th[(size_t)n * 16 + lane];
        __syncwarp();
        // four consecutive sites per lane, 128 sites per pass: the draws of a pass are the next cnt outputs of the generator
        for (int c0 = 0; c0 < cols; c0 += 128) {
          const int cnt = min(128, cols - c0);
          const int off = 4 * lane, site0 = c0 + off;
          unsigned u[4];
          if (ev.rng == 0) {
            evo_draw4(mt, pos, left, cnt, lane, u);
          } else {
#pragma unroll
            for (int t = 0; t < 4; t++) u[t] = philox_draw(seed, (unsigned)n, (unsigned)(site0 + t));
          }
          if (off < cnt) {
            const unsigned pw = pseq ? pseq[site0 >> 2] : 0u;  // parent states of the four sites; the root draws from row 0 of its table
            unsigned st = 0u, ch = 0u;
#pragma unroll
            for (int t = 0; t < 4; t++) {
              const uint4 tt = s_thr[warp][(pw >> (8 * t)) & 3u];
              const unsigned state = (u[t] > tt.x) + (u[t] > tt.y) + (u[t] > tt.z);
              st |= state << (8 * t);
              ch |= ((0x54474341u >> (8 * state)) & 0xffu) << (8 * t);  // "ACGT"[state]
            }
            if (oseq) oseq[site0 >> 2] = st;  // (the padding bytes of the last word are never read as sites)
            if (orow) {
              if (row_aligned && off + 3 < cnt) {
                *reinterpret_cast<unsigned*>(orow + site0) = ch;
              } else {
#pragma unroll
                for (int t = 0; t < 4; t++)
                  if (off + t < cnt) orow[site0 + t] = (unsigned char)(ch >> (8 * t));
              }
            }
          }
        }
        __syncwarp();  // a child reads its parent's sequence written by other lanes; s_thr is rewritten for the next node
      }
      __syncwarp();  // the state in shared memory is free for the next sample
    }
  }
}

// ---------------------------------------------------------------------------------------------
// (f4) k_pair_rows: rows b of the pairwise matrices Sk[k][state][b][i] (src/score.c:496-535) of one strand of the native
// alignment -- the only part of Sk that backtrack() (src/score.c:558-797) reads when the --eps plots are drawn
// (src/postscript.c:303).  One thread per (requested row, species) walks the end codons i = b+2, b+5, ...; z and sigma are
// formed on the fly as in k_prep / k_sigma.  cls: class bytes of the N rows in the strand's own orientation (nucleotide code
// in bits 0-1); c0[x] = column of the x-th non-gap character of row 0.  out: [row][N][3][L+1], zero-filled by the host.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    k_pair_rows(const unsigned char* __restrict__ cls, const int* __restrict__ c0, const float* __restrict__ scores,
                const SigmaTables* __restrict__ tables, const int* __restrict__ b_list, int n_rows, int N, int cols, int L,
                Params prm, float* __restrict__ out) {
  __shared__ SigmaTables s_tab;
  __shared__ __align__(8) uint64_t s_bar;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
    mbar_expect_tx(&s_bar, (unsigned)sizeof(SigmaTables));
    bulk_g2s(&s_tab, tables, (unsigned)sizeof(SigmaTables), &s_bar);
  }
  __syncthreads();
  mbar_wait(&s_bar, 0);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = idx / (N - 1), k = idx % (N - 1) + 1;
  if (r >= n_rows) return;
  const int b = b_list[r];
  const unsigned char* row0 = cls;
  const unsigned char* rowk = cls + (size_t)k * cols;
  float* o0 = out + ((size_t)r * N + k) * 3 * (L + 1);
  float* o1 = o0 + (L + 1);
  float* o2 = o1 + (L + 1);
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;  // :500-504
  for (int x = b + 2; x <= L; x += 3) {
    const int c1 = c0[x - 2], c2 = c0[x - 1], c3 = c0[x];
    const int a = (x > 3) ? c0[x - 3] + 1 : 0;  // getBlock, src/misc.c:198-207
    int gk = 0;
    for (int col = a; col <= c3; col++) gk += (rowk[col] & CLS_GAP) ? 1 : 0;
    int diff = gk - ((c3 - a + 1) - 3);
    diff = diff < 0 ? -diff : diff;
    const int m = diff % 3;  // 1: z = +1, 2: z = -1 (src/misc.c:230-244)
    float n0, n1, n2;
    if (m == 0) {
      const unsigned a1 = row0[c1], a2 = row0[c2], a3 = row0[c3];
      const unsigned b1 = rowk[c1], b2 = rowk[c2], b3 = rowk[c3];
      const unsigned qa = ((a1 & 3u) << 4) | ((a2 & 3u) << 2) | (a3 & 3u);
      const unsigned qb = ((b1 & 3u) << 4) | ((b2 & 3u) << 2) | (b3 & 3u);
      float v;
      if (((a1 | a2 | a3 | b1 | b2 | b3) & CLS_N) | (b1 & b2 & b3 & CLS_X)) {
        v = 0.0f;  // src/score.c:394-404
      } else if (qa == qb) {
        v = 0.0f;  // :409
      } else {
        const int pepA = s_tab.transcode[qa], pepB = s_tab.transcode[qb];
        if (pepA < 0)
          v = prm.stop0;  // :414-416
        else if (pepB < 0)
          v = prm.stopk;  // :418-420
        else {
          const unsigned d = qa ^ qb;
          const int h = ((d & 0x30u) != 0) + ((d & 0x0cu) != 0) + ((d & 0x03u) != 0);
          v = s_tab.blosum[pepA * 24 + pepB] - scores[k * 4 + h];  // :422-425
        }
      }
      n0 = s0 + v;  // :506-510
      n1 = s1 + prm.omega;
      n2 = s2 + prm.omega;
    } else if (m == 1) {  // :512-521
      const float d0 = s0 + prm.Delta, d1 = s1 + prm.Delta, d2 = s2 + prm.Delta;
      const float w0 = s0 + prm.Omega, w1 = s1 + prm.Omega, w2 = s2 + prm.Omega;
      n0 = d0 > w2 ? d0 : w2;
      n1 = w0 > d1 ? w0 : d1;
      n2 = w1 > d2 ? w1 : d2;
    } else {  // :523-533
      const float d0 = s0 + prm.Delta, d1 = s1 + prm.Delta, d2 = s2 + prm.Delta;
      const float w0 = s0 + prm.Omega, w1 = s1 + prm.Omega, w2 = s2 + prm.Omega;
      n0 = d0 > w1 ? d0 : w1;
      n1 = d1 > w2 ? d1 : w2;
      n2 = d2 > w0 ? d2 : w0;
    }
    s0 = n0;
    s1 = n1;
    s2 = n2;
    o0[x] = n0;
    o1[x] = n1;
    o2[x] = n2;
  }
}

// ---------------------------------------------------------------------------------------------
// Issue-rate calibration: the instruction mix of the DP fast path (4 FADD : 1 FMNMX3 per cell) on
// register-resident chains, no memory traffic.  Gives the practical FP32/ALU issue ceiling of the
// device at its current clocks; bench.py reports the DP kernel against it.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_calib(float* __restrict__ out, int iters, float sg, float om) {
  float S0[8], S1[8], S2[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    S0[i] = threadIdx.x * 0.001f + i;
    S1[i] = S0[i] - 1.0f;
    S2[i] = S0[i] - 2.0f;
    acc[i] = 0.0f;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      S0[i] += sg;
      S1[i] += om;
      S2[i] += om;
      acc[i] += max3f(S0[i], S1[i], S2[i]);
    }
  }
  float r = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; i++) r += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace rc
