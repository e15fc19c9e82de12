// rc_device.cuh -- device-side data layout shared by the kernels and the host planner of
// libRNAcode_cuda (sm_100a only).  See DESIGN.md "Data layout in HBM".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rc {

constexpr int TILE = 16;          // codon steps per staged sigma/z tile
constexpr int REC_SLOTS = 3;      // tie-band slots in a row record
constexpr int DP_WARPS = 4;       // warps per DP CTA (each warp owns one task)
constexpr int REG_MAX_NK = 16;    // largest N-1 handled by the register-resident DP kernels
constexpr int SMP_WARPS = 4;      // warps per CTA of the sample-major DP kernel (8 when the sigma table limits the CTAs per SM)
constexpr int SMP_MAX_WARPS = 8;
#ifndef RC_SMP_SEG
#define RC_SMP_SEG 16
#endif
constexpr int SMP_SEG = RC_SMP_SEG;  // end codons per shared-memory stage of the streaming sample-major DP kernel

// class byte of one alignment character (k_pack): what calculateSigma / getBlock / revAln need
//   bits 0-1  ntMap[c]                 (forward strand code; anything but ACGTU -> 0, src/RNAcode.c:94-98)
//   bits 2-3  ntMap[revcomp(c)]        (reverse strand code; revAln complements upper-case ACGTU only,
//                                       src/rnaz_utils.c:327-333)
//   bit  4    c == 'N'                 (src/score.c:400-404)
//   bit  5    c == 'X'                 (the "XXX" guard, src/score.c:394-396)
//   bit  6    c == '-'                 (gap; getBlock src/misc.c:214-227, calculateSigma src/score.c:385)
constexpr unsigned CLS_N = 0x10, CLS_X = 0x20, CLS_GAP = 0x40;

// One scored row (start codon i of one frame of one strand of one alignment): everything the
// sequential part of getHSS (src/score.c:888-961) needs to know about the row.  32 bytes.
struct __align__(16) RowRec {
  float Emax;              // largest positive entry of the row (by the fresh lenient fold)
  float vF;                // value of the last entry the fresh fold accepted
  float be[REC_SLOTS];     // tie band: accepted entries within 1e-4 of Emax ...
  unsigned short jF;       // end codon of the last accepted entry
  unsigned short n;        // band entries in use | 0x8000 if the band overflowed; 0 = row has no positive entry
  unsigned short bj[REC_SLOTS];  // ... and their end codons
  unsigned short pad;
};
static_assert(sizeof(RowRec) == 32, "RowRec must be 32 bytes");

struct HssDev {  // one emitted high-scoring segment of the native alignment
  int startSite, endSite;
  float score;
};

struct Params {
  float Delta, Omega, omega, stop0, stopk;
};

// Per alignment block, resident for the life of a batch.
struct BlockDev {
  int N, cols, L, NK;       // NK = N-1 scored species
  int n_inst;               // 1 (native) + n_samples
  int hss_warp;             // getHSS scan: 1 = one warp per (instance, strand, frame) (k_hss), 0 = one thread (k_hss_thr)
  int inst_stride;          // bytes between consecutive instances in raw/cls (N*cols rounded up to 16)
  int zstride;              // u32 words per z tile: layout 0: NK rounded up to 4 (one word per species);
                            // layout 1: TILE (one word per step, 2 bits per species); layout 3: nchunk*TILE
  int layout;               // 0: sigma tile [k][TILE] (k_dp, any NK); 1: sigma tile [TILE][RS] (k_dp_reg, NK <= 16);
                            // 2: sample-major [group of 32 instances][step][RSB/4][lane][4] (k_dp_smp, short blocks);
                            // 3: sigma tile [chunk][TILE][RS] (k_dp_chain: species cut into nchunk chunks of <= nkw)
                            // 5: sample-major in nchunk species chunks of chunk_base (+1 for the first chunk_rem) QUADS:
                            //    [chunk][group of 32 instances][step][3 quads][lane][4] (k_dp_smp<., true>, one launch per chunk)
  int chain_tasks;          // layout 3: consecutive tasks one CTA of k_dp_chain works through
  int smp_seg;              // layouts 2 / 5: 1 = the frame's sigma table does not fit shared memory, streamed in segments (k_dp_smps)
  int smp_fused;            // layouts 2 / 5, resident table: 1 = the DP kernel builds its sigma table itself from class bytes
                            // (k_dp_smpf; no sigma scratch, no k_sigma_smp launch for this block)
  int smp_fold;             // k_dp_smpf: 1 = a last group of at most 16 instances runs several start-codon pairs side by side
  int smpf_allf;            // k_dp_smpf: 1 = one CTA per (strand, group) takes the three frames in turn (CtaDesc.sf = 6 + strand)
  int sig_p2;               // layouts 2 / 5 with a sigma table in HBM: 1 = the table is built from k_pack2's packed rows (k_sigma_p2)
                            // instead of class bytes (k_sigma_smp)
  int p2_words;             // k_dp_smpf: 32-bit words per packed row = ceil(L / 16)
  long long p2_off;         // k_dp_smpf: u32 offset of the block's packed rows (k_pack2):
                            // [group of 32 instances][strand][row][word w][lane]: 2-bit nucleotide codes of the row's characters at
                            // the reference's non-gap columns, 16 positions per word (position 16w + t of the strand's reading
                            // direction in bits 2t, 2t+1), groups padded with zero lanes
  long long p2f_off;        // k_dp_smpf: u32 offset of the rows' flag words [group][strand][row]: bit = lane whose row holds an
                            // 'N' or 'X' at some reference position (such codons take the byte-wise path)
  float fold_B;             // k_dp_smpf: half-width of the ambiguity zone of the getHSS fold in species-sum space (see RowFoldS)
  int nchunk, nkw;          // layout 3: chunks and species per chunk (template NK of k_dp_chain); chunk g holds
  int chunk_base, chunk_rem;  //   chunk_base + (g < chunk_rem) species starting at g*chunk_base + min(g, chunk_rem)
  int sig_tile;             // floats per sigma tile
  int sig_ks, sig_cs;       // strides (floats) of species / step inside a sigma tile
  int sites[3], ntiles[3];  // codon sites / tiles per frame
  long long raw_off, cls_off;   // byte offsets.  cls: [instance][inst_stride].  raw: sample i (instance 1+i) at raw_off + i*N*cols,
                                // exactly as the caller holds them (one contiguous copy per block)
  long long nat_off;            // byte offset in raw of the native rows (all natives of a batch are contiguous: one copy)
  long long cols0_off;          // int offset: [2][L+1], forward column (0-based) of position x (1-based) per strand
  long long scores_off;         // float offset: [2][N][4]
  long long z_off[2][3];        // u32 offset: [tile][zstride]
  long long res_off;            // float offset: [n_inst][6]   best emitted score per (strand, frame), -1 none, -2 overflow
  long long hss_off[2][3];      // HssDev offset, capacity sites/3+1 each
  long long hsscnt_off;         // int offset: [6]
  float fNK, rcpNK;
};

// A contiguous range of instances of one block whose sigma tiles and row records are in scratch.
struct Item {
  int block, inst0, ninst, pad;
  long long sigma_off[2][3];  // float offset: [inst_local][tile][NK][TILE]
  long long rec_off[2][3];    // RowRec offset: [inst_local][sites]
  long long dense_off[2][3];  // float offset (dense fallback only): [inst_local][sites*(sites+1)/2]
  long long part_off[2][3];   // float2 offset (layout 5 only): [group][pair][end codon][lane] partial species sums
};

// Null-alignment simulation (kernel d) of one block: the tree in seq-gen's evolution order.
struct EvoDev {
  int block;
  int n_nodes, n_internal;
  int rng;                 // 0: MT19937 exactly as seq-gen consumes it; 1: Philox4x32-10 counter-based
  long long node_off;      // int offset: [n_nodes][4] = parent, alignment row (-1: internal), slot of its sequence in scratch, 0
  long long thr_off;       // u32 offset: [n_nodes][4 parent states][4]: a draw u moves past cumulative entry j iff u > thr
  long long seed_off;      // u32 offset: [n_samples]
  long long seq_off;       // byte offset into the evolve scratch: [n_samples][n_internal][cols]
};

struct CtaDesc {
  int item;
  int sf;     // strand*3 + frame
  int task0;  // first task (warp) of this CTA inside the problem
};

}  // namespace rc
