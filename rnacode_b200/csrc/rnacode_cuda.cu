// rnacode_cuda.cu -- libRNAcode_cuda: host-side planner + C ABI (include/rnacode_cuda.h).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
//
// A batch is a list of alignment blocks.  Persistent per-batch device data: raw bytes, class bytes,
// position->column maps, z tiles, score tables, results.  Scratch (sigma tiles + row records) is
// sized per chunk; a chunk is a list of Items (block, instance range).  Per chunk: k_sigma -> k_dp
// (one launch per register-blocking class) -> k_hss.  No CPU fallback exists anywhere in this file.
#include "../../include/rnacode_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "rc_kernels.cuh"

using namespace rc;

#define RC_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ctx_fail(ctx, std::string(#call) + ": " + cudaGetErrorString(_e));                       \
      return RC_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

struct rc_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  unsigned char* d_lut = nullptr;
  std::string err;
  long force_dense = 0;
  long band_slots = REC_SLOTS;
  long scratch_mb = 2048;
  long no_smp = 0;
  long no_chain = 0;
  long smp_warps_forced = 0;
  long hss_thr_tasks = 3000;  // blocks with at least this many (instance, strand, frame) scans use one thread per scan for frames of up to
                              // HSS_THR_MAX_SITES codons (0: never).  10x500 n=1000: 0.89 -> 0.39 ms; 10x1200: 0.54 -> 0.37; 10x4806: 0.45 -> 1.24 (kept on warps)
  long reg_max_nk = 12;  // row-major alignments with more scored species take k_dp_chain (k_dp_reg<13..16> spills: 17x3000 14.7 vs 11.1 ms)
  long no_smps = 0;          // never stream the sigma table in segments (k_dp_smps)
  long smps_max_sites = 420;  // longest frame (codons) for k_dp_smps; beyond, the row-major k_dp_reg is faster (break-even ~1200 columns)
  long no_fold = 0;           // 1: a last group of at most 16 instances is scored like a full one (k_dp_smpf)
  long no_fused = 0;          // 1: never build the sigma table inside the sample-major DP kernel (k_dp_smpf)
  long no_sig_rows3 = 0;      // 1: layout-3 sigma tiles by the generic k_sigma instead of k_sigma_rows3
  long no_allf = 0;           // 1: k_dp_smpf with one CTA per frame instead of one per strand (three frames in turn)
  long reg_tu = -1;           // k_dp_regtu (three additions per cell, needs omega = -2^k): -1 = where the batch's frameshift density
                              // makes it the faster kernel, 0 = never, 1 = always
  long no_sig_p2 = 0;         // 1: sigma tables of the sample-major layouts always from class bytes (k_sigma_smp), never from packed rows
  long tail_max = 0;          // a sample-major block whose instance count leaves 1..tail_max instances in its last group of 32 scores
                              // those instances row-major (lanes = rows) instead of in a warp with that many live lanes (0: never)
  long smpc_max_sites = 0;    // longest frame (codons) for the STREAMED chunked sample-major route of wide alignments
                              // (measured slower than k_dp_chain: 50x800 11.3 vs 4.5 ms; kept as an experiment switch)
  int smem_optin = 0;
  int sm_count = 0;
  // device-memory cache: buffers of destroyed batches are kept and handed to the next batch, so that a
  // steady stream of rc_batch_create / rc_batch_destroy calls does not pay cudaMalloc / cudaFree every time
  std::vector<std::pair<void*, size_t>> free_blocks;
  std::vector<std::pair<void*, size_t>> live_blocks;
  size_t cached_bytes = 0;
};

static void ctx_fail(rc_ctx* ctx, const std::string& msg) {
  if (ctx) ctx->err = msg;
}

// Dynamic shared memory above 48 KB has to be allowed per kernel.  The attribute belongs to the (device, function) pair, not to
// the caller: two host threads that run batches on the same device must not race between "allow my size" and "launch", so
// every kernel is always allowed the device's opt-in maximum (less its static shared memory) -- the same value from every
// thread.
template <class K>
static cudaError_t allow_max_smem(rc_ctx* ctx, K kernel) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin - (int)fa.sharedSizeBytes);
}

static void* ctx_alloc(rc_ctx* ctx, size_t bytes) {
  bytes = (std::max<size_t>(bytes, 256) + 255) / 256 * 256;
  // best fit among cached blocks that are not wastefully large
  int best = -1;
  for (int i = 0; i < (int)ctx->free_blocks.size(); i++) {
    const size_t sz = ctx->free_blocks[i].second;
    if (sz >= bytes && sz <= bytes * 2 + (1u << 20) && (best < 0 || sz < ctx->free_blocks[best].second)) best = i;
  }
  void* p = nullptr;
  if (best >= 0) {
    p = ctx->free_blocks[best].first;
    ctx->live_blocks.push_back(ctx->free_blocks[best]);
    ctx->cached_bytes -= ctx->free_blocks[best].second;
    ctx->free_blocks.erase(ctx->free_blocks.begin() + best);
    return p;
  }
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    // release the cache and retry once
    cudaGetLastError();
    for (auto& fb : ctx->free_blocks) cudaFree(fb.first);
    ctx->free_blocks.clear();
    ctx->cached_bytes = 0;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  }
  ctx->live_blocks.push_back({p, bytes});
  return p;
}

static void ctx_free(rc_ctx* ctx, void* p) {
  if (!p) return;
  for (size_t i = 0; i < ctx->live_blocks.size(); i++)
    if (ctx->live_blocks[i].first == p) {
      ctx->free_blocks.push_back(ctx->live_blocks[i]);
      ctx->cached_bytes += ctx->live_blocks[i].second;
      ctx->live_blocks.erase(ctx->live_blocks.begin() + i);
      // keep the cache bounded: drop the largest blocks beyond 8 GiB
      while (ctx->cached_bytes > ((size_t)8 << 30) && !ctx->free_blocks.empty()) {
        size_t big = 0;
        for (size_t j = 1; j < ctx->free_blocks.size(); j++)
          if (ctx->free_blocks[j].second > ctx->free_blocks[big].second) big = j;
        cudaFree(ctx->free_blocks[big].first);
        ctx->cached_bytes -= ctx->free_blocks[big].second;
        ctx->free_blocks.erase(ctx->free_blocks.begin() + big);
      }
      return;
    }
  cudaFree(p);  // not ours (should not happen)
}

namespace {

// DP launch classes: 0..15 = k_dp_reg<NK = class+1>; 16 = k_dp<R = 2> (17 <= NK <= 24); 17 = k_dp<R = 1>;
// 18..33 = k_dp_smp<NK = class-17> (sample-major, short blocks);
// 34.. = k_dp_chain<NKW> with W warps: class = CHAIN_CLASS0 + (W-2)*CHAIN_NKW_SPAN + (NKW - CHAIN_NKW_MIN)
constexpr int SMP_CLASS0 = REG_MAX_NK + 2;
constexpr int CHAIN_CLASS0 = 2 * REG_MAX_NK + 2;
constexpr int CHAIN_NKW_MIN = 5, CHAIN_NKW_MAX = 12, CHAIN_NKW_SPAN = CHAIN_NKW_MAX - CHAIN_NKW_MIN + 1;
// then k_dp_smp<., true> (layout 5: wide alignments in short blocks), one class per number of species quads Q = ceil(NK/4)
constexpr int CHAIN_MAX_CHUNKS = 42;  // 499 scored species / 12; more than CHAIN_PASS_WARPS chunks run in several passes
constexpr int SMPC_CLASS0 = CHAIN_CLASS0 + (CHAIN_MAX_CHUNKS - 1) * CHAIN_NKW_SPAN;
constexpr int SMPC_Q_MIN = 4, SMPC_Q_MAX = 125;
// and the segmented (streaming) variants k_dp_smps of both sample-major kinds
constexpr int SMPS_CLASS0 = SMPC_CLASS0 + (SMPC_Q_MAX - SMPC_Q_MIN + 1);
constexpr int SMPCS_CLASS0 = SMPS_CLASS0 + REG_MAX_NK;
// and the fused variants k_dp_smpf (sigma table built inside the DP kernel) of the resident-table kinds
constexpr int SMPF_CLASS0 = SMPCS_CLASS0 + (SMPC_Q_MAX - SMPC_Q_MIN + 1);
constexpr int SMPCF_CLASS0 = SMPF_CLASS0 + REG_MAX_NK;
constexpr int N_CLASSES = SMPCF_CLASS0 + (SMPC_Q_MAX - SMPC_Q_MIN + 1);
constexpr size_t SMP_SMEM_MAX = 200 * 1024;  // sigma table + z words of one CTA of k_dp_smp
constexpr int SMP_MIN_INST = 16;             // fewer instances than this: the row-major kernels are the better fit

struct Chunk {
  size_t item0 = 0, nitems = 0;          // range in the batch's item array
  size_t cta0[N_CLASSES] = {}, ncta[N_CLASSES] = {};  // per class range in the CTA array
  int maxNK[N_CLASSES] = {}, maxZs[N_CLASSES] = {};
  size_t sigma_floats = 0, rec_count = 0, part_count = 0, max_smp_smem = 0, max_smps_smem = 0;
  size_t max_smpf_smem = 0;  // k_dp_smpf: shared memory of a CTA without the fold records
  int n_smp_unfused = 0;     // sample-major items that still need k_sigma_smp
  int n_sig_p2 = 0;          // ... of which k_sigma_p2 serves this many (packed rows); its grid and shared memory:
  int max_p2_chunks = 1;
  size_t max_p2_smem = 0;
  long long max_sigma_work = 0;  // largest ninst*2*(L-2) of an item, for the k_sigma grid
  int max_ninst = 0;
  int n_layout[6] = {0, 0, 0, 0, 0, 0};  // items per sigma layout
  int max_smp_quads = 0;  // most species quads of a sample-major item (k_sigma_smp spreads them over gridDim.z)
  int max_smp_npos = 0;   // most reference positions of a sample-major item (k_sigma_smp: chunks of SIG_PCH positions)
  int hss_warp_items = 0;  // items whose frames are long enough for the warp-per-task k_hss
};

struct EventPair {
  cudaEvent_t a, b;
  int stage;  // 0 pack (evolve + pack + prep), 1 sigma, 2 dp, 3 hss, 4 k_pack alone (nested in 0)
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr size_t MAX_ITEM_INST = 32768;  // instances per item (a multiple of 32): -n 100000 on tiny blocks must not overflow gridDim.y

int class_of(const BlockDev& bd) {
  if (bd.layout == 3) return CHAIN_CLASS0 + (bd.nchunk - 2) * CHAIN_NKW_SPAN + (bd.nkw - CHAIN_NKW_MIN);
  if (bd.layout == 5) return (bd.smp_seg ? SMPCS_CLASS0 : (bd.smp_fused ? SMPCF_CLASS0 : SMPC_CLASS0)) + ((bd.NK + 3) / 4 - SMPC_Q_MIN);
  if (bd.layout == 2) return (bd.smp_seg ? SMPS_CLASS0 : (bd.smp_fused ? SMPF_CLASS0 : SMP_CLASS0)) + bd.NK - 1;
  if (bd.layout == 1) return bd.NK - 1;
  return bd.NK <= 24 ? REG_MAX_NK : REG_MAX_NK + 1;
}
int class_R(int cl) { return cl == REG_MAX_NK + 1 ? 1 : 2; }

// k_dp_chain: alignments with more than CHAIN_PASS_WARPS species chunks run in several launches ("passes") of nearly
// equal width; this is the width of the widest pass
int chain_passes(int W) { return (W + CHAIN_PASS_WARPS - 1) / CHAIN_PASS_WARPS; }
int chain_pass_width(int W) { return (W + chain_passes(W) - 1) / chain_passes(W); }
// float2 entries per instance of the partial sums handed from one pass to the next (layout 3): one per (row group, tile from
// the group's first tile on, end codon, lane)
size_t chain_part_entries(int sites, int ntiles) {
  const size_t ng = (size_t)(sites + 63) / 64;
  return (ng * ntiles - 2 * ng * (ng - 1)) * TILE * 32;
}

// layout 5: (start-codon pair, end codon) entries of one frame: sum over pairs p of (sites - 2p)
size_t part_entries(int sites) {
  const size_t np = (size_t)(sites + 1) / 2;
  return np * sites - np * (np - 1);
}

// dynamic shared memory of a sample-major DP CTA without the fold state: the frame's sigma table (k_dp_smp) or two
// stages of SMP_SEG end codons (k_dp_smps), plus the frame's z words and the barriers
size_t smp_smem_bytes(const BlockDev& bd, int f, int layout, int seg) {
  const size_t rsb = layout == 5 ? 12 : (size_t)(bd.NK + 3) / 4 * 4;
  const size_t zb = ((size_t)bd.sites[f] * 4 + 15) / 16 * 16;
  if (seg) return 2 * (size_t)SMP_SEG * rsb * 32 * 4 + zb + 32;
  return (size_t)bd.sites[f] * rsb * 32 * 4 + zb + 16;
}

// k_dp_smpf: shared memory of a CTA (frame 0 is the longest) without the fold records
size_t smpf_smem_bytes(const BlockDev& bd, int layout) {
  const int nsp = layout == 5 ? 12 : (bd.NK + 3) / 4 * 4;
  return SmpfCfg<1>::off_rec(bd.sites[0], nsp * 32 * 4, nsp, (bd.L + 15) / 16);
}

// floats of sigma scratch for `ninst` instances of one (strand, frame) of a block
size_t sigma_floats_sf(const BlockDev& bd, int f, int ninst) {
  if ((bd.layout == 2 || bd.layout == 5) && bd.smp_fused) return 0;  // the table only ever exists in shared memory
  if (bd.layout == 2) {
    const size_t rsb = (size_t)(bd.NK + 3) / 4 * 4;
    return (size_t)((ninst + 31) / 32) * bd.sites[f] * rsb * 32;
  }
  if (bd.layout == 5) return (size_t)bd.nchunk * ((ninst + 31) / 32) * bd.sites[f] * 12 * 32;
  return (size_t)ninst * bd.ntiles[f] * bd.sig_tile;
}

// sigma / z layout of a block (see BlockDev)
void set_layout(BlockDev& bd, int layout) {
  bd.layout = layout;
  bd.nchunk = bd.nkw = bd.chunk_base = bd.chunk_rem = 0;
  if (layout == 3) {  // species chunks for k_dp_chain
    const int W = (bd.NK + CHAIN_NKW_MAX - 1) / CHAIN_NKW_MAX;
    bd.nchunk = W;
    bd.chunk_base = bd.NK / W;
    bd.chunk_rem = bd.NK % W;
    bd.nkw = bd.chunk_base + (bd.chunk_rem ? 1 : 0);
    const int rs = (bd.nkw + 1 + 3) / 4 * 4;
    bd.sig_tile = W * TILE * rs;
    bd.sig_ks = 1;
    bd.sig_cs = rs;
    bd.zstride = W * TILE;
  } else if (layout == 5) {  // species chunks counted in quads: Q quads over W = ceil(Q/3) chunks of 2-3 quads
    const int Q = (bd.NK + 3) / 4, W = (Q + 2) / 3;
    bd.nchunk = W;
    bd.chunk_base = Q / W;
    bd.chunk_rem = Q % W;
    bd.sig_tile = 0;
    bd.sig_ks = 0;
    bd.sig_cs = 0;
    bd.zstride = W * TILE;
  } else if (layout == 2) {  // z as in layout 1 (one word per step); sigma addressed explicitly (sigma_floats_sf)
    bd.sig_tile = 0;
    bd.sig_ks = 0;
    bd.sig_cs = 0;
    bd.zstride = TILE;
  } else if (layout == 1) {
    const int rs = (bd.NK + 1 + 3) / 4 * 4;  // RegCfg<NK>::RS
    bd.sig_tile = TILE * rs;
    bd.sig_ks = 1;
    bd.sig_cs = rs;
    bd.zstride = TILE;
  } else {
    bd.sig_tile = bd.NK * TILE;
    bd.sig_ks = TILE;
    bd.sig_cs = 1;
    bd.zstride = (int)align_up((size_t)bd.NK, 4);
  }
}

// what depends on the layout besides set_layout(): tasks per CTA of k_dp_chain, frame padding of k_dp_reg
void finish_layout(BlockDev& bd) {
  if (bd.layout == 3) {
    // tasks per CTA: enough tiles per CTA to amortise the W-1 tiles the warp pipeline needs to fill and drain
    // (a row group of a frame with T tiles has T - 4g tiles: about T/2 on average), but not more: long rows are
    // better balanced with one task per CTA
    const double avg_tiles = std::max(1.0, 0.5 * bd.ntiles[0]);
    const int wp = chain_pass_width(bd.nchunk);  // warps per CTA = depth of the pipeline
    bd.chain_tasks = (int)std::min<double>(CHAIN_MAX_TASKS, std::max(1.0, std::ceil(4.0 * (wp - 1) / avg_tiles)));
  }
  if (bd.layout == 1)  // k_dp_reg stages RC_REG_TILE end codons at a time: pad the frame to whole stages
    for (int f = 0; f < 3; f++) bd.ntiles[f] = (bd.ntiles[f] + RC_REG_TILE / TILE - 1) / (RC_REG_TILE / TILE) * (RC_REG_TILE / TILE);
}

}  // namespace

struct rc_batch {
  rc_ctx* ctx = nullptr;
  double fs_events = 0.0, fs_cells = 0.0;  // row-major register-kernel blocks: estimated (species, codon) pairs with a frameshift / all
  bool use_tu = false;                      // k_dp_regtu instead of k_dp_reg (decided per batch in rc_batch_create)
  int n_blocks = 0;  // blocks of the caller; b->blocks may hold "tail" blocks after them (see rc_batch_create)
  std::vector<int> tail_of;   // per caller block: index of its tail block in `blocks`, or -1
  std::vector<int> main_inst; // per caller block: instances [0, main_inst) use the block's own layout, the rest the tail block's
  std::vector<rc_block_desc> descs;
  std::vector<BlockDev> blocks;
  std::vector<Item> items;
  std::vector<CtaDesc> ctas;
  std::vector<Chunk> chunks;
  Params prm{};
  SigmaTables tables{};
  PairTables ptab{};
  // device, persistent
  BlockDev* d_blocks = nullptr;
  Item* d_items = nullptr;
  CtaDesc* d_ctas = nullptr;
  unsigned char *d_raw = nullptr, *d_cls = nullptr;
  unsigned *d_p2 = nullptr, *d_p2f = nullptr;  // packed rows (k_pack2) and their flag words, for the blocks scored by k_dp_smpf
  size_t p2_words = 0, p2f_words = 0;
  int max_n_inst = 1, max_fused_N = 1, max_fused_cols = 1;
  PairTables ptab2{};             // PairTables indexed by packed codons (first position in the low bits)
  PairTables* d_ptab2 = nullptr;
  int* d_cols0 = nullptr;
  float* d_scores = nullptr;
  unsigned* d_z = nullptr;
  float* d_res = nullptr;
  HssDev* d_hss = nullptr;
  int* d_hsscnt = nullptr;
  int* d_ovf = nullptr;
  SigmaTables* d_tables = nullptr;
  PairTables* d_ptab = nullptr;
  // null-alignment simulation (kernel d)
  std::vector<EvoDev> evos;
  std::vector<int> evo_nodes;
  std::vector<unsigned> evo_thr, evo_seeds;
  std::vector<int> evo_of_block;  // -1: samples come from the host
  size_t evo_seq_bytes = 0;
  EvoDev* d_evos = nullptr;
  int* d_evo_nodes = nullptr;
  unsigned *d_evo_thr = nullptr, *d_evo_seeds = nullptr;
  unsigned char* d_evo_seq = nullptr;
  int evo_spw = EVO_SPW;        // samples per k_evolve task in this batch
  std::vector<int> evo_task0;   // prefix sums of the k_evolve tasks (EVO_SPW samples each) per simulated block
  int* d_evo_task0 = nullptr;
  unsigned* d_evo_mt = nullptr;  // seeded generator states, one slot of EVO_SPW states per persistent warp of k_evolve
  size_t evo_mt_bytes = 0;
  int evo_max_samples = 0;
  // device, scratch
  float* d_sigma = nullptr;
  RowRec* d_recs = nullptr;
  float2* d_partial = nullptr;  // layout 5: partial species sums between chunk launches
  size_t part_count = 0;
  float* d_dense = nullptr;
  size_t dense_floats = 0;
  // sizes
  size_t raw_bytes = 0, cls_bytes = 0, nat_bytes = 0, cols0_ints = 0, scores_floats = 0, z_words = 0, res_floats = 0, hss_count = 0, hsscnt_ints = 0;
  size_t sigma_floats = 0, rec_count = 0;
  // host results
  std::vector<float> h_res;
  std::vector<HssDev> h_hss;
  std::vector<int> h_hsscnt;
  std::vector<float> h_scores;
  std::vector<unsigned char> h_nat;  // the native rows of all blocks, staged for one copy
  bool uploaded = false, ran = false, downloaded = false;
  rc_batch_stats stats{};
  std::vector<EventPair> events;
  size_t device_bytes = 0;
};

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" const char* rc_version(void) { return "libRNAcode_cuda 0.1 (sm_100a)"; }

extern "C" void rc_default_params(rc_params* p) {
  if (!p) return;
  p->Delta = -10.0f;  // src/RNAcode.c:68-72
  p->Omega = -4.0f;
  p->omega = -2.0f;
  p->stopPenalty_k = -8.0f;
  p->stopPenalty_0 = -9999.0f;
}

static void build_lut(unsigned char* lut) {
  for (int c = 0; c < 256; c++) {
    auto nt = [](int ch) -> unsigned {
      switch (ch) {
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': case 'U': case 'u': return 3;
        default: return 0;
      }
    };
    int rcmp = c;
    switch (c) {  // revAln, src/rnaz_utils.c:327-333
      case 'T': rcmp = 'A'; break;
      case 'U': rcmp = 'A'; break;
      case 'C': rcmp = 'G'; break;
      case 'G': rcmp = 'C'; break;
      case 'A': rcmp = 'T'; break;
      default: break;
    }
    unsigned v = nt(c) | (nt(rcmp) << 2);
    if (c == 'N') v |= CLS_N;
    if (c == 'X') v |= CLS_X;
    if (c == '-') v |= CLS_GAP;
    lut[c] = (unsigned char)v;
  }
}

extern "C" int rc_create(rc_ctx** out, int device) {
  if (!out) return RC_ERR_ARG;
  *out = nullptr;
  rc_ctx* ctx = new (std::nothrow) rc_ctx();
  if (!ctx) return RC_ERR_NOMEM;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    // no CPU fallback: the product fails loudly without a usable CUDA device
    delete ctx;
    return RC_ERR_CUDA;
  }
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return RC_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // experiment switches (same meaning as rc_set_option): RNACODE_CUDA_NO_SMPS, RNACODE_CUDA_SMPC_MAX_SITES
  if (const char* e = getenv("RNACODE_CUDA_NO_SMPS")) ctx->no_smps = atol(e) ? 1 : 0;
  if (const char* e = getenv("RNACODE_CUDA_SMPC_MAX_SITES")) ctx->smpc_max_sites = atol(e);
  if (const char* e = getenv("RNACODE_CUDA_SMPS_MAX_SITES")) ctx->smps_max_sites = atol(e);
  if (const char* e = getenv("RNACODE_CUDA_SMP_WARPS")) ctx->smp_warps_forced = atol(e);
  if (const char* e = getenv("RNACODE_CUDA_HSS_THR_TASKS")) ctx->hss_thr_tasks = atol(e);
  if (const char* e = getenv("RNACODE_CUDA_NO_FUSED")) ctx->no_fused = atol(e) ? 1 : 0;
  if (const char* e = getenv("RNACODE_CUDA_NO_SIG_P2")) ctx->no_sig_p2 = atol(e) ? 1 : 0;
  if (const char* e = getenv("RNACODE_CUDA_NO_SIG_ROWS3")) ctx->no_sig_rows3 = atol(e) ? 1 : 0;
  if (const char* e = getenv("RNACODE_CUDA_NO_ALLF")) ctx->no_allf = atol(e) ? 1 : 0;
  if (const char* e = getenv("RNACODE_CUDA_REG_TU")) ctx->reg_tu = std::max(-1L, std::min(1L, atol(e)));
  if (const char* e = getenv("RNACODE_CUDA_NO_FOLD")) ctx->no_fold = atol(e) ? 1 : 0;
  if (const char* e = getenv("RNACODE_CUDA_TAIL_MAX")) ctx->tail_max = std::max(0L, std::min(31L, atol(e)));
  if (const char* e = getenv("RNACODE_CUDA_REG_MAX_NK")) ctx->reg_max_nk = std::max(12L, std::min<long>(REG_MAX_NK, atol(e)));
  unsigned char lut[256];
  build_lut(lut);
  if (cudaMalloc(&ctx->d_lut, 256) != cudaSuccess ||
      cudaMemcpy(ctx->d_lut, lut, 256, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return RC_ERR_CUDA;
  }
  *out = ctx;
  return RC_OK;
}

extern "C" void rc_destroy(rc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->d_lut) cudaFree(ctx->d_lut);
  for (auto& fb : ctx->free_blocks) cudaFree(fb.first);
  for (auto& lb : ctx->live_blocks) cudaFree(lb.first);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

extern "C" const char* rc_last_error(const rc_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

extern "C" int rc_set_stream(rc_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RC_ERR_ARG;
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return RC_OK;
}

extern "C" int rc_device_count(int* count) {
  if (!count) return RC_ERR_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    *count = 0;
    return RC_ERR_CUDA;
  }
  *count = n;
  return n > 0 ? RC_OK : RC_ERR_CUDA;
}

extern "C" int rc_set_option(rc_ctx* ctx, const char* key, long value) {
  if (!ctx || !key) return RC_ERR_ARG;
  std::string k(key);
  if (k == "force_dense") ctx->force_dense = value ? 1 : 0;
  else if (k == "band_slots") {
    if (value < 1 || value > REC_SLOTS) { ctx_fail(ctx, "band_slots must be 1..3"); return RC_ERR_ARG; }
    ctx->band_slots = value;
  } else if (k == "no_smp") {
    ctx->no_smp = value ? 1 : 0;
  } else if (k == "no_chain") {
    ctx->no_chain = value ? 1 : 0;
  } else if (k == "no_smps") {
    ctx->no_smps = value ? 1 : 0;
  } else if (k == "hss_thr_tasks") {
    ctx->hss_thr_tasks = value;
  } else if (k == "reg_max_nk") {
    if (value < 12 || value > REG_MAX_NK) { ctx_fail(ctx, "reg_max_nk must be 12..16"); return RC_ERR_ARG; }
    ctx->reg_max_nk = value;
  } else if (k == "smps_max_sites") {
    ctx->smps_max_sites = value;
  } else if (k == "smpc_max_sites") {
    ctx->smpc_max_sites = value;
  } else if (k == "no_fold") {
    ctx->no_fold = value ? 1 : 0;
  } else if (k == "no_fused") {
    ctx->no_fused = value ? 1 : 0;
  } else if (k == "no_sig_p2") {
    ctx->no_sig_p2 = value ? 1 : 0;
  } else if (k == "no_sig_rows3") {
    ctx->no_sig_rows3 = value ? 1 : 0;
  } else if (k == "no_allf") {
    ctx->no_allf = value ? 1 : 0;
  } else if (k == "reg_tu") {
    ctx->reg_tu = std::max(-1L, std::min(1L, value));
  } else if (k == "tail_max") {
    if (value < 0 || value > 31) { ctx_fail(ctx, "tail_max must be 0..31"); return RC_ERR_ARG; }
    ctx->tail_max = value;
  } else if (k == "scratch_mb") {
    if (value < 1) { ctx_fail(ctx, "scratch_mb must be >= 1"); return RC_ERR_ARG; }
    ctx->scratch_mb = value;
  } else {
    ctx_fail(ctx, "unknown option " + k);
    return RC_ERR_ARG;
  }
  return RC_OK;
}

extern "C" int rc_calibrate_issue(rc_ctx* ctx, double* lane_ops_per_s) {
  if (!ctx || !lane_ops_per_s) return RC_ERR_ARG;
  RC_CUDA(cudaSetDevice(ctx->device));
  const int ctas = ctx->sm_count * 8, threads = 256, iters = 8192;
  float* d_out = nullptr;
  RC_CUDA(cudaMalloc((void**)&d_out, sizeof(float) * ctas * threads));
  cudaEvent_t a, b;
  RC_CUDA(cudaEventCreate(&a));
  RC_CUDA(cudaEventCreate(&b));
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {  // first repetition warms up
    RC_CUDA(cudaEventRecord(a, ctx->stream));
    k_calib<<<ctas, threads, 0, ctx->stream>>>(d_out, iters, 0.25f, -2.0f);
    RC_CUDA(cudaEventRecord(b, ctx->stream));
    RC_CUDA(cudaEventSynchronize(b));
    float ms = 0.0f;
    RC_CUDA(cudaEventElapsedTime(&ms, a, b));
    const double ops = (double)ctas * threads * iters * 8.0 * 5.0;  // 8 chains x (4 FADD + 1 FMNMX3)
    if (rep > 0) best = std::max(best, ops / (ms * 1e-3));
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_out);
  *lane_ops_per_s = best;
  return RC_OK;
}

// ------------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------------
static void free_batch_device(rc_batch* b) {
  rc_ctx* ctx = b->ctx;
  void* ptrs[] = {b->d_blocks, b->d_items, b->d_ctas, b->d_raw, b->d_cls, b->d_cols0, b->d_scores, b->d_z, b->d_res,
                  b->d_hss, b->d_hsscnt, b->d_ovf, b->d_tables, b->d_ptab, b->d_sigma, b->d_recs, b->d_dense, b->d_partial, b->d_p2, b->d_p2f, b->d_ptab2,
                  b->d_evos, b->d_evo_nodes, b->d_evo_thr, b->d_evo_seeds, b->d_evo_seq, b->d_evo_task0, b->d_evo_mt};
  for (void* p : ptrs) ctx_free(ctx, p);
  for (auto& e : b->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  b->events.clear();
}

// Append CTA descriptors for `items[i0..i0+n)` with row-group size 32*R.
static void build_ctas(const std::vector<BlockDev>& blocks, const std::vector<Item>& items, size_t i0, size_t n, int R,
                       int want_class, std::vector<CtaDesc>& out) {
  for (size_t i = i0; i < i0 + n; i++) {
    const Item& it = items[i];
    const BlockDev& bd = blocks[it.block];
    if (want_class >= 0 && class_of(bd) != want_class) continue;
    if (bd.layout == 2 && bd.smp_fused && bd.smpf_allf && want_class >= 0) {
      // k_dp_smpf: one CTA per (strand, group) works through the three frames (they share the tables and the packed rows)
      if (bd.sites[0] > 0)
        for (int strand = 0; strand < 2; strand++)
          for (int g = 0; g < (it.ninst + 31) / 32; g++) out.push_back(CtaDesc{(int)i, 6 + strand, g});
      continue;
    }
    if ((bd.layout == 2 || bd.layout == 5) && want_class >= 0) {
      for (int sf = 0; sf < 6; sf++) {
        if (bd.sites[sf % 3] <= 0) continue;
        for (int g = 0; g < (it.ninst + 31) / 32; g++) out.push_back(CtaDesc{(int)i, sf, g});
      }
      continue;
    }
    for (int sf = 0; sf < 6; sf++) {
      const int sites = bd.sites[sf % 3];
      if (sites <= 0) continue;
      const long long ngroups = (sites + 32 * R - 1) / (32 * R);
      const long long ntasks = ngroups * it.ninst;
      const int per_cta = (bd.layout == 3 && want_class >= 0) ? bd.chain_tasks : DP_WARPS;  // k_dp_chain: the CTA's warps share each task
      for (long long t = 0; t < ntasks; t += per_cta) out.push_back(CtaDesc{(int)i, sf, (int)t});
    }
  }
}

// BLOSUM as float + the standard genetic code in the reference's encoding (src/code.c:26-35): A,C,G,T = 0..3, amino acids in
// BLOSUM order ARNDCQEGHILKMFPSTWYV, stop = -1
static void fill_tables(SigmaTables& t, const int* blosum) {
  for (int i = 0; i < 576; i++) t.blosum[i] = (float)blosum[i];
  static const char* tcag_aa = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
  static const char* aa_order = "ARNDCQEGHILKMFPSTWYV";
  static const int tcag_to_acgt[4] = {3, 1, 0, 2};
  for (int i = 0; i < 64; i++) {
    int b1 = tcag_to_acgt[i / 16], b2 = tcag_to_acgt[(i / 4) % 4], b3 = tcag_to_acgt[i % 4];
    char aa = tcag_aa[i];
    t.transcode[b1 * 16 + b2 * 4 + b3] = (aa == '*') ? -1 : (signed char)(strchr(aa_order, aa) - aa_order);
  }
}

// calculateSigma's case analysis per (reference codon, species codon), see PairTables
static void fill_pair_tables(PairTables& pt, const SigmaTables& t, const Params& prm) {
  for (int i = 0; i < 576; i++) pt.val[i] = t.blosum[i];
  pt.val[PT_ZERO] = 0.0f;
  pt.val[PT_STOP0] = prm.stop0;
  pt.val[PT_STOPK] = prm.stopk;
  pt.val[579] = 0.0f;
  for (int qa = 0; qa < 64; qa++)
    for (int qb = 0; qb < 64; qb++) {
      const int pepA = t.transcode[qa], pepB = t.transcode[qb];
      const int d = qa ^ qb, h = ((d & 0x30) != 0) + ((d & 0x0c) != 0) + ((d & 0x03) != 0);
      unsigned e;
      if (h == 0) e = PT_ZERO;          // src/score.c:409, tested before the stop codons
      else if (pepA < 0) e = PT_STOP0;  // :414-416
      else if (pepB < 0) e = PT_STOPK;  // :418-420
      else e = (unsigned)(pepA * 24 + pepB) | ((unsigned)h << 10);
      pt.t[qa * 64 + qb] = (unsigned short)e;
    }
}

extern "C" int rc_batch_create(rc_ctx* ctx, const rc_block_desc* descs, int n_blocks, const rc_params* params,
                               const int* blosum, rc_batch** out) {
  if (!ctx || !out) return RC_ERR_ARG;
  *out = nullptr;
  if (!descs || n_blocks < 1 || !params || !blosum) {
    ctx_fail(ctx, "rc_batch_create: NULL argument or empty batch");
    return RC_ERR_ARG;
  }
  rc_batch* b = new (std::nothrow) rc_batch();
  if (!b) return RC_ERR_NOMEM;
  b->ctx = ctx;
  b->n_blocks = n_blocks;
  b->descs.assign(descs, descs + n_blocks);
  b->prm = Params{params->Delta, params->Omega, params->omega, params->stopPenalty_0, params->stopPenalty_k};
  fill_tables(b->tables, blosum);
  fill_pair_tables(b->ptab, b->tables, b->prm);
  b->ptab2 = b->ptab;
  for (unsigned qa = 0; qa < 64; qa++)
    for (unsigned qb = 0; qb < 64; qb++) b->ptab2.t[qa * 64 + qb] = b->ptab.t[pt_swap(qa) * 64 + pt_swap(qb)];

  b->blocks.resize(n_blocks);
  double cells = 0;
  for (int i = 0; i < n_blocks; i++) {
    const rc_block_desc& d = descs[i];
    if (d.N < 2 || d.N > 500 || d.cols < 1 || d.cols > 190000 || !d.rows || !d.scores_fwd || !d.scores_rev || d.n_samples < 0 ||
        false) {
      ctx_fail(ctx, "rc_batch_create: invalid block descriptor " + std::to_string(i));
      delete b;
      return RC_ERR_ARG;
    }
    BlockDev& bd = b->blocks[i];
    memset(&bd, 0, sizeof(bd));
    bd.N = d.N;
    bd.cols = d.cols;
    int L = 0;
    for (int c = 0; c < d.cols; c++) L += d.rows[c] != '-';  // getSeqLength, src/misc.c:272-289
    bd.L = L;
    bd.NK = d.N - 1;
    bd.n_inst = 1 + d.n_samples;
    b->max_n_inst = std::max(b->max_n_inst, bd.n_inst);
    bd.inst_stride = (int)align_up((size_t)d.N * d.cols, 16);
    bd.fNK = (float)bd.NK;
    bd.rcpNK = 1.0f / bd.fNK;
    bd.cls_off = (long long)b->cls_bytes;
    b->cls_bytes += (size_t)bd.inst_stride * bd.n_inst;
    bd.raw_off = (long long)b->raw_bytes;  // the samples, stride N*cols; the natives follow all samples (nat_off, below)
    b->raw_bytes += align_up((size_t)d.N * d.cols * d.n_samples, 16);
    bd.nat_off = (long long)b->nat_bytes;
    b->nat_bytes += (size_t)bd.inst_stride;
    bd.cols0_off = (long long)b->cols0_ints;
    b->cols0_ints += 2 * (size_t)(L + 1);
    bd.scores_off = (long long)b->scores_floats;
    b->scores_floats += 2 * (size_t)d.N * 4;
    double P = 0;
    for (int f = 0; f < 3; f++) {
      bd.sites[f] = L >= 3 ? (L - f) / 3 : 0;
      bd.ntiles[f] = (bd.sites[f] + TILE - 1) / TILE;
      P += (double)bd.sites[f] * (bd.sites[f] + 1) / 2;
    }
    cells += (double)bd.n_inst * 2.0 * bd.NK * P;
    // getHSS scan: long frames get a warp each (coalesced record fetch, skip by ballot) unless the block has so many scans
    // that one thread each already fills the GPU
    bd.hss_warp = bd.sites[0] >= HSS_WARP_MIN_SITES &&
                  !(ctx->hss_thr_tasks > 0 && (long)bd.n_inst * 6 >= ctx->hss_thr_tasks && bd.sites[0] <= HSS_THR_MAX_SITES);
    {
      // Which DP kernel.  Delta > 0 needs the general max(sum, Delta) (k_dp); the chunked kernels may carry
      // a dummy species, which is only neutral for omega <= 0.
      const int reg_max = (int)std::min<long>(REG_MAX_NK, ctx->reg_max_nk);  // widest alignment for the row-major register kernel
      const bool wide_ok = params->Delta <= 0.0f && params->omega <= 0.0f;
      const bool smp_ok = !ctx->no_smp && bd.n_inst >= SMP_MIN_INST && bd.sites[0] >= 1;  // enough instances to fill the lanes
      const size_t smem_cap = std::min<size_t>(SMP_SMEM_MAX, (size_t)ctx->smem_optin);
      int layout = 0, seg = 0;
      if (params->Delta <= 0.0f && bd.NK <= REG_MAX_NK && smp_ok && smp_smem_bytes(bd, 0, 2, 0) <= smem_cap) {
        layout = 2;  // short block, many instances: sample-major with the frame's sigma table resident
      } else if (params->Delta <= 0.0f && bd.NK <= REG_MAX_NK && smp_ok && !ctx->no_smps && bd.sites[0] <= ctx->smps_max_sites) {
        layout = 2;  // mid-length block: sample-major, sigma table streamed in segments
        seg = 1;
      } else if (params->Delta <= 0.0f && (bd.NK <= reg_max || (bd.NK <= REG_MAX_NK && !wide_ok))) {
        layout = 1;  // row-major, one warp holds all species in registers
      } else if (wide_ok && bd.NK > REG_MAX_NK && smp_ok &&
                 (smp_smem_bytes(bd, 0, 5, 0) <= smem_cap || (!ctx->no_smps && bd.sites[0] <= ctx->smpc_max_sites))) {
        layout = 5;  // wide alignment, short block, many instances: sample-major, one launch per species chunk
        seg = smp_smem_bytes(bd, 0, 5, 0) <= smem_cap ? 0 : 1;
      } else if (wide_ok && !ctx->no_chain && (bd.NK + CHAIN_NKW_MAX - 1) / CHAIN_NKW_MAX <= CHAIN_MAX_CHUNKS &&
                 (bd.NK + CHAIN_NKW_MAX - 1) / CHAIN_NKW_MAX >= 2) {
        layout = 3;  // wide alignment: species chunks pipelined through the warps of a CTA
      } else if (params->Delta <= 0.0f && bd.NK <= REG_MAX_NK) {
        layout = 1;
      }
      set_layout(bd, layout);
      bd.smp_seg = seg;
      finish_layout(bd);
      if (layout == 1) {
        // Frameshift density of the block, estimated from its gap runs: a run whose length is not a multiple of three shifts the
        // frame of one codon of its species (of every species when it sits in the reference row).  k_dp_regtu pays for every
        // frameshift, k_dp_reg for every cell.
        double ev = 0.0;
        for (int r = 0; r < d.N; r++) {
          const char* row = d.rows + (size_t)r * d.cols;
          int run = 0;
          for (int c = 0; c <= d.cols; c++) {
            if (c < d.cols && row[c] == '-') { run++; continue; }
            if (run % 3 != 0) ev += r == 0 ? (double)bd.NK : 1.0;
            run = 0;
          }
        }
        b->fs_events += ev * bd.n_inst;
        b->fs_cells += (double)bd.NK * (L / 3.0) * bd.n_inst;
      }
      // resident-table sample-major blocks build their sigma table inside the DP kernel when the staged rows fit as well
      // ... unless that costs a resident CTA (two CTAs of 8 warps per SM need <= 113 KB each): wide alignments in chunks have
      // a 100 KB table already and stay with k_sigma_smp + k_dp_smp
      const size_t sm_bytes = (size_t)228 * 1024;
      const bool two_fused = 2 * (smpf_smem_bytes(bd, layout) + (size_t)SMP_MAX_WARPS * 64 * sizeof(RowRec) + 1024) <= sm_bytes;
      const bool two_unfused = 2 * (smp_smem_bytes(bd, 0, layout, 0) + (size_t)SMP_WARPS * 64 * sizeof(RowRec) + 1024) <= sm_bytes;
      if ((layout == 2 || layout == 5) && !seg && !ctx->no_fused && !ctx->force_dense && (two_fused || !two_unfused) && bd.cols <= P2_MAX_COLS &&
          (layout == 5 || bd.NK <= 12) &&  // k_dp_smpf<13..16> would spill at the 128 registers two CTAs per SM allow
          smpf_smem_bytes(bd, layout) + (size_t)SMP_WARPS * 64 * sizeof(RowRec) <= (size_t)ctx->smem_optin) {
        bd.smp_fused = 1;
        bd.smpf_allf = (layout == 2 && !ctx->no_allf) ? 1 : 0;
        bd.smp_fold = ctx->no_fold ? 0 : 1;
        b->max_fused_N = std::max(b->max_fused_N, bd.N);
        b->max_fused_cols = std::max(b->max_fused_cols, bd.cols);
        bd.p2_words = (bd.L + 15) / 16;
        const size_t groups = (size_t)(bd.n_inst + 31) / 32;
        bd.p2_off = (long long)b->p2_words;
        b->p2_words += groups * 2 * bd.N * bd.p2_words * 32 + 64;  // + slack: the last row's codon look-ups read one word ahead
        bd.p2f_off = (long long)b->p2f_words;
        b->p2f_words += groups * 2 * bd.N;
      }
      // the other sample-major blocks keep their sigma table in HBM; it is built from the same packed rows (k_sigma_p2) when the
      // rows of a species chunk fit shared memory twice per SM, from class bytes (k_sigma_smp) otherwise
      if ((layout == 2 || layout == 5) && !bd.smp_fused && !ctx->no_sig_p2 && !ctx->force_dense && bd.cols <= P2_MAX_COLS && bd.L >= 3 &&
          SigP2Cfg::total((bd.L + 15) / 16, layout == 5 ? 12 : (bd.NK + 3) / 4 * 4) <= (size_t)100 * 1024) {
        bd.sig_p2 = 1;
        b->max_fused_N = std::max(b->max_fused_N, bd.N);
        b->max_fused_cols = std::max(b->max_fused_cols, bd.cols);
        bd.p2_words = (bd.L + 15) / 16;
        const size_t groups = (size_t)(bd.n_inst + 31) / 32;
        bd.p2_off = (long long)b->p2_words;
        b->p2_words += groups * 2 * bd.N * bd.p2_words * 32 + 64;
        bd.p2f_off = (long long)b->p2f_words;
        b->p2f_words += groups * 2 * bd.N;
      }
      if ((layout == 2 || layout == 5) && !seg) bd.smp_fold = ctx->no_fold ? 0 : 1;  // resident-table kernels fold a short last group
      if (layout == 2 || layout == 5) {
        // B of RowFoldS: (N-1) * 1.0002e-4 for the tolerance of getHSS's tie rule plus 2^-21 of the largest species sum a
        // row can reach (per end codon and species at most the largest sigma -- BLOSUM entry or stop penalty minus the
        // smallest expected score -- or a positive penalty) for the roundings of the two quotients and of the bound itself
        int bmax = 0;
        for (int q = 0; q < 576; q++) bmax = std::max(bmax, blosum[q]);
        double smax = 0.0;
        for (int k = 1; k < d.N; k++) {
          double lo = 0.0;
          for (int h = 1; h < 4; h++) lo = std::min(lo, (double)std::min(d.scores_fwd[4 * k + h], d.scores_rev[4 * k + h]));
          double g = std::max({(double)bmax, (double)params->stopPenalty_0, (double)params->stopPenalty_k}) - lo;
          g = std::max({g, (double)params->omega, (double)params->Omega, 0.0});
          smax += g;
        }
        smax *= std::max(1, bd.sites[0]);
        bd.fold_B = (float)(bd.NK * 1.0002e-4 + smax / 2097152.0);
      }
    }
    for (int s = 0; s < 2; s++)
      for (int f = 0; f < 3; f++) {
        bd.z_off[s][f] = (long long)b->z_words;
        b->z_words += (size_t)bd.ntiles[f] * bd.zstride;
        bd.hss_off[s][f] = (long long)b->hss_count;
        b->hss_count += (size_t)bd.sites[f] / 3 + 1;
      }
    bd.res_off = (long long)b->res_floats;
    b->res_floats += (size_t)bd.n_inst * 6;
    bd.hsscnt_off = (long long)b->hsscnt_ints;
    b->hsscnt_ints += 6;
  }
  {
    // k_dp_regtu (three additions per cell, deferred S1 / S2 bookkeeping at frameshifts) wins below about 0.6 % of the
    // (species, codon) pairs with a frameshift -- the real examples/genomic.maf has 0.17 %, SURVEY 8(d)'s generator 3 % -- and
    // needs omega = -2^k
    int ex = 0;
    const bool pow2 = params->omega < 0.0f && std::frexp(-params->omega, &ex) == 0.5f && ex >= -20 && ex <= 20;
    const double g = b->fs_cells > 0 ? b->fs_events / b->fs_cells : 1.0;
    b->use_tu = pow2 && (ctx->reg_tu == 1 || (ctx->reg_tu < 0 && g < 0.006));
  }
  b->stats.cells = cells;
  // Tail blocks.  The sample-major kernels put 32 instances into a warp; a block with 101 instances (RNAcode's default -n 100)
  // leaves 5 of them in a fourth warp that costs as much as a full one.  Those instances are scored by the row-major
  // kernels instead (lanes = rows of one instance): a "tail block" is a second BlockDev of the same alignment -- same
  // class bytes, maps, scores, results -- with a row-major sigma / z layout; the block's last instances become items of it.
  b->tail_of.assign(n_blocks, -1);
  b->main_inst.resize(n_blocks);
  for (int i = 0; i < n_blocks; i++) {
    const BlockDev bd = b->blocks[i];
    b->main_inst[i] = bd.n_inst;
    const int r = bd.n_inst % 32;
    if ((bd.layout != 2 && bd.layout != 5) || bd.n_inst < 64 || r == 0 || r > ctx->tail_max || bd.L < 3) continue;
    const int reg_max = (int)std::min<long>(REG_MAX_NK, ctx->reg_max_nk);
    int alt = -1;
    if (bd.NK <= reg_max) alt = 1;
    else if (params->omega <= 0.0f && !ctx->no_chain && (bd.NK + CHAIN_NKW_MAX - 1) / CHAIN_NKW_MAX >= 2 &&
             (bd.NK + CHAIN_NKW_MAX - 1) / CHAIN_NKW_MAX <= CHAIN_MAX_CHUNKS) alt = 3;
    if (alt < 0) continue;
    BlockDev vb = bd;
    for (int f = 0; f < 3; f++) vb.ntiles[f] = (vb.sites[f] + TILE - 1) / TILE;
    set_layout(vb, alt);
    vb.smp_seg = 0;
    finish_layout(vb);
    for (int s = 0; s < 2; s++)
      for (int f = 0; f < 3; f++) {
        vb.z_off[s][f] = (long long)b->z_words;
        b->z_words += (size_t)vb.ntiles[f] * vb.zstride;
      }
    b->tail_of[i] = (int)b->blocks.size();
    b->main_inst[i] = bd.n_inst - r;
    b->blocks.push_back(vb);
  }
  for (BlockDev& bd : b->blocks) bd.nat_off += (long long)b->raw_bytes;  // the natives follow the samples in d_raw
  b->stats.pack_chars = (double)b->cls_bytes;

  // chunking: fill items until the scratch budget is reached
  const size_t budget = (size_t)ctx->scratch_mb << 20;
  Chunk cur;
  size_t cur_bytes = 0;
  auto close_chunk = [&]() {
    if (cur.nitems == 0) return;
    b->chunks.push_back(cur);
    cur = Chunk();
    cur.item0 = b->items.size();
    cur_bytes = 0;
  };
  // segments of instances: a caller block's instances [0, main_inst) with its own layout, the rest with its tail block's
  struct Seg { int block, inst0, inst1; };
  std::vector<Seg> segs;
  for (int i = 0; i < n_blocks; i++) {
    segs.push_back(Seg{i, 0, b->main_inst[i]});
    if (b->tail_of[i] >= 0) segs.push_back(Seg{b->tail_of[i], b->main_inst[i], b->blocks[i].n_inst});
  }
  for (const Seg& sg : segs) {
    const int i = sg.block;
    const BlockDev& bd = b->blocks[i];
    if (bd.L < 3) continue;  // nothing to score (the reference skips such blocks, src/RNAcode.c:147-150)
    size_t sig_per_inst = 0, rec_per_inst = 0;
    for (int f = 0; f < 3; f++) {
      sig_per_inst += 2 * sigma_floats_sf(bd, f, 32) / 32;
      rec_per_inst += 2 * (size_t)bd.sites[f];
    }
    size_t part_per_inst = 0;  // layout 5: partial species sums handed from one chunk's launch to the next (float2 per lane)
    if (bd.layout == 5)
      for (int f = 0; f < 3; f++) part_per_inst += 2 * part_entries(bd.sites[f]);
    if (bd.layout == 3 && chain_passes(bd.nchunk) > 1)
      for (int f = 0; f < 3; f++) part_per_inst += 2 * chain_part_entries(bd.sites[f], bd.ntiles[f]);
    const size_t bytes_per_inst = sig_per_inst * sizeof(float) + rec_per_inst * sizeof(RowRec) + part_per_inst * sizeof(float2);
    int inst = sg.inst0;
    while (inst < sg.inst1) {
      size_t room = budget > cur_bytes ? (budget - cur_bytes) / bytes_per_inst : 0;
      room = std::min<size_t>(room, MAX_ITEM_INST);  // grids carry (instances of an item) * 6 / warps in gridDim.y (<= 65535)
      if (room == 0) {
        if (cur.nitems == 0) room = 1;  // a single instance always goes through
        else { close_chunk(); continue; }
      }
      int take = (int)std::min<size_t>(room, (size_t)(sg.inst1 - inst));
      if ((bd.layout == 2 || bd.layout == 5) && take < sg.inst1 - inst) {  // instance groups of 32 must not straddle chunks
        if (take >= 32) take = take / 32 * 32;
        else if (cur.nitems == 0) take = std::min(32, sg.inst1 - inst);
        else { close_chunk(); continue; }
      }
      Item it;
      memset(&it, 0, sizeof(it));
      it.block = i;
      it.inst0 = inst;
      it.ninst = take;
      for (int s = 0; s < 2; s++)
        for (int f = 0; f < 3; f++) {
          it.sigma_off[s][f] = (long long)cur.sigma_floats;
          cur.sigma_floats += sigma_floats_sf(bd, f, take);
          it.rec_off[s][f] = (long long)cur.rec_count;
          cur.rec_count += (size_t)take * bd.sites[f];
          if (bd.layout == 5) {
            it.part_off[s][f] = (long long)cur.part_count;
            cur.part_count += (size_t)((take + 31) / 32) * part_entries(bd.sites[f]) * 32;
          }
          if (bd.layout == 3 && chain_passes(bd.nchunk) > 1) {
            it.part_off[s][f] = (long long)cur.part_count;
            cur.part_count += (size_t)take * chain_part_entries(bd.sites[f], bd.ntiles[f]);
          }
        }
      const int cl = class_of(bd);
      cur.maxNK[cl] = std::max(cur.maxNK[cl], bd.NK);
      if ((bd.layout == 2 || bd.layout == 5) && bd.smp_fused) {
        cur.max_smpf_smem = std::max(cur.max_smpf_smem, smpf_smem_bytes(bd, bd.layout));
      } else if (bd.layout == 2 || bd.layout == 5) {
        cur.n_smp_unfused++;
        if (bd.sig_p2) {
          cur.n_sig_p2++;
          cur.max_p2_chunks = std::max(cur.max_p2_chunks, bd.layout == 5 ? bd.nchunk : 1);
          cur.max_p2_smem = std::max(cur.max_p2_smem, SigP2Cfg::total(bd.p2_words, bd.layout == 5 ? 12 : (bd.NK + 3) / 4 * 4));
        }
        if (!bd.smp_seg) cur.max_smp_smem = std::max(cur.max_smp_smem, smp_smem_bytes(bd, 0, bd.layout, 0));
        if (bd.smp_seg) cur.max_smps_smem = std::max(cur.max_smps_smem, smp_smem_bytes(bd, 0, bd.layout, 1));
        cur.max_smp_quads = std::max(cur.max_smp_quads, (bd.NK + 3) / 4);
        cur.max_smp_npos = std::max(cur.max_smp_npos, bd.L - 2);
      }
      cur.maxZs[cl] = std::max(cur.maxZs[cl], bd.zstride);
      if (bd.layout != 2 && bd.layout != 5)  // grid of k_sigma / k_sigma_rows (sample-major items have kernels of their own)
        cur.max_sigma_work = std::max(cur.max_sigma_work, (long long)take * 2 * (bd.L - 2 + 3 * RC_REG_TILE));
      cur.max_ninst = std::max(cur.max_ninst, take);
      if (bd.hss_warp) cur.hss_warp_items++;
      cur.n_layout[bd.layout]++;
      b->items.push_back(it);
      cur.nitems++;
      cur_bytes += (size_t)take * bytes_per_inst;
      inst += take;
    }
  }
  close_chunk();
  for (Chunk& ch : b->chunks) {
    for (int cl = 0; cl < N_CLASSES; cl++) {
      ch.cta0[cl] = b->ctas.size();
      if (ch.maxNK[cl] > 0) build_ctas(b->blocks, b->items, ch.item0, ch.nitems, class_R(cl), cl, b->ctas);
      ch.ncta[cl] = b->ctas.size() - ch.cta0[cl];
    }
    b->sigma_floats = std::max(b->sigma_floats, ch.sigma_floats);
    b->rec_count = std::max(b->rec_count, ch.rec_count);
    b->part_count = std::max(b->part_count, ch.part_count);
  }

  // device allocations
  if (cudaSetDevice(ctx->device) != cudaSuccess) { delete b; return RC_ERR_CUDA; }
  size_t total = 0;
  auto dalloc = [&](void** p, size_t bytes) -> bool {
    bytes = std::max<size_t>(bytes, 256);
    total += bytes;
    *p = ctx_alloc(ctx, bytes);
    return *p != nullptr;
  };
  bool ok = dalloc((void**)&b->d_blocks, sizeof(BlockDev) * b->blocks.size()) &&
            dalloc((void**)&b->d_items, sizeof(Item) * b->items.size()) &&
            dalloc((void**)&b->d_ctas, sizeof(CtaDesc) * b->ctas.size()) &&
            dalloc((void**)&b->d_raw, b->raw_bytes + b->nat_bytes + 64) && dalloc((void**)&b->d_cls, b->cls_bytes + 64) && dalloc((void**)&b->d_cols0, sizeof(int) * b->cols0_ints) &&
            dalloc((void**)&b->d_scores, sizeof(float) * b->scores_floats) &&
            dalloc((void**)&b->d_z, sizeof(unsigned) * b->z_words) && dalloc((void**)&b->d_res, sizeof(float) * b->res_floats) &&
            dalloc((void**)&b->d_hss, sizeof(HssDev) * b->hss_count) &&
            dalloc((void**)&b->d_hsscnt, sizeof(int) * b->hsscnt_ints) && dalloc((void**)&b->d_ovf, sizeof(int)) &&
            dalloc((void**)&b->d_tables, sizeof(SigmaTables)) && dalloc((void**)&b->d_ptab, sizeof(PairTables)) && dalloc((void**)&b->d_sigma, sizeof(float) * b->sigma_floats) &&
            dalloc((void**)&b->d_recs, sizeof(RowRec) * b->rec_count) &&
            (b->part_count == 0 || dalloc((void**)&b->d_partial, sizeof(float2) * b->part_count)) &&
            (b->p2_words == 0 || (dalloc((void**)&b->d_p2, sizeof(unsigned) * b->p2_words) &&
                                  dalloc((void**)&b->d_p2f, sizeof(unsigned) * b->p2f_words) &&
                                  dalloc((void**)&b->d_ptab2, sizeof(PairTables))));
  if (!ok) {
    ctx_fail(ctx, std::string("rc_batch_create: device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    free_batch_device(b);
    delete b;
    return RC_ERR_NOMEM;
  }
  b->device_bytes = total;
  b->evo_of_block.assign(n_blocks, -1);
  b->h_res.assign(b->res_floats, -1.0f);
  b->h_hss.resize(b->hss_count);
  b->h_hsscnt.assign(b->hsscnt_ints, 0);
  *out = b;
  return RC_OK;
}

extern "C" void rc_batch_destroy(rc_batch* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  free_batch_device(b);
  delete b;
}

// ------------------------------------------------------------------------------------------------
// null-alignment simulation set-up
// ------------------------------------------------------------------------------------------------
// SetState (seqgen/evolve.c:167-175) compares r = u * (1.0/4294967295.0) (genrand_real1, seqgen/twister.c:162-166)
// with a cumulative probability P in double.  u -> r is monotone, so `r > P` <=> `u > T(P)` with
// T(P) = max{u : (double)u * (1.0/4294967295.0) <= P}, found by bisection with the very same expression.
static unsigned threshold_of(double P) {
  const double c = 1.0 / 4294967295.0;
  volatile double r0 = 0.0 * c;
  if (!(r0 <= P)) return 0u;  // P < 0 (not a probability): treat like 0
  // start from the real-number estimate P / c and correct it with the very expression seq-gen evaluates
  double est = P * 4294967295.0;
  unsigned long long u = est >= 4294967295.0 ? 0xffffffffull : (est <= 0.0 ? 0ull : (unsigned long long)est);
  for (;;) {  // largest u with (double)u * c <= P
    volatile double r = (double)u * c;
    if (r <= P) break;
    if (u == 0) return 0u;
    u--;
  }
  while (u < 0xffffffffull) {
    volatile double r = (double)(u + 1) * c;
    if (!(r <= P)) break;
    u++;
  }
  return (unsigned)u;
}

// deferred: (offset into evo_thr, cumulative probabilities, count) of thresholds still to be computed -- rc_batch_set_evolve_many
// converts the tables of all its blocks on several host threads
struct ThrJob { size_t off; const double* cum; size_t n; };
static int set_evolve_impl(rc_batch* b, int block, const rc_tree_desc* tree, const unsigned int* seeds, int rng,
                           std::vector<ThrJob>* deferred);

extern "C" int rc_batch_set_evolve(rc_batch* b, int block, const rc_tree_desc* tree, const unsigned int* seeds, int rng) {
  return set_evolve_impl(b, block, tree, seeds, rng, nullptr);
}

static int set_evolve_impl(rc_batch* b, int block, const rc_tree_desc* tree, const unsigned int* seeds, int rng,
                           std::vector<ThrJob>* deferred) {
  if (!b) return RC_ERR_ARG;
  rc_ctx* ctx = b->ctx;
  if (block < 0 || block >= b->n_blocks || !tree || !seeds || tree->n_nodes < 2 || !tree->parent || !tree->row || !tree->cum ||
      (rng != RC_RNG_MT19937 && rng != RC_RNG_PHILOX)) {
    ctx_fail(ctx, "rc_batch_set_evolve: bad argument");
    return RC_ERR_ARG;
  }
  if (b->uploaded) {
    ctx_fail(ctx, "rc_batch_set_evolve after rc_batch_upload");
    return RC_ERR_STATE;
  }
  if (b->evo_of_block[block] >= 0) {
    ctx_fail(ctx, "rc_batch_set_evolve: block already has a tree");
    return RC_ERR_STATE;
  }
  const BlockDev& bd = b->blocks[block];
  const int n_samples = bd.n_inst - 1;
  std::vector<char> seen(bd.N, 0);
  EvoDev ev;
  memset(&ev, 0, sizeof(ev));
  ev.block = block;
  ev.n_nodes = tree->n_nodes;
  ev.rng = rng;
  ev.node_off = (long long)b->evo_nodes.size();
  ev.thr_off = (long long)b->evo_thr.size();
  ev.seed_off = (long long)b->evo_seeds.size();
  int n_internal = 0;
  for (int n = 0; n < tree->n_nodes; n++) {
    const int parent = tree->parent[n], row = tree->row[n];
    if ((n == 0) != (parent < 0) || parent >= n || row >= bd.N || (row >= 0 && seen[row]) ||
        (parent >= 0 && tree->row[parent] >= 0)) {
      ctx_fail(ctx, "rc_batch_set_evolve: nodes must be in evolution order (parents first, root at 0), tips map to distinct rows");
      return RC_ERR_ARG;
    }
    if (row >= 0) seen[row] = 1;
    b->evo_nodes.push_back(parent);
    b->evo_nodes.push_back(row);
    b->evo_nodes.push_back(row >= 0 ? -1 : n_internal++);
    b->evo_nodes.push_back(0);
    if (!deferred)
      for (int k = 0; k < 16; k++) b->evo_thr.push_back(threshold_of(tree->cum[(size_t)n * 16 + k]));
  }
  if (deferred) {
    deferred->push_back(ThrJob{b->evo_thr.size(), tree->cum, (size_t)tree->n_nodes * 16});
    b->evo_thr.resize(b->evo_thr.size() + (size_t)tree->n_nodes * 16);
  }
  for (int r = 0; r < bd.N; r++)
    if (!seen[r]) {
      ctx_fail(ctx, "rc_batch_set_evolve: alignment row " + std::to_string(r) + " is not a tip of the tree");
      return RC_ERR_ARG;
    }
  ev.n_internal = n_internal;
  for (int i = 0; i < n_samples; i++) b->evo_seeds.push_back(seeds[i]);
  ev.seq_off = (long long)b->evo_seq_bytes;
  b->evo_seq_bytes += (size_t)n_samples * n_internal * ((bd.cols + 3) & ~3);  // rows of the internal nodes padded to words
  b->evo_max_samples = std::max(b->evo_max_samples, n_samples);
  b->evo_of_block[block] = (int)b->evos.size();
  b->evos.push_back(ev);
  return RC_OK;
}

extern "C" int rc_batch_set_evolve_many(rc_batch* b, int first, int n, const rc_tree_desc* trees, const unsigned int* const* seeds,
                                        int rng) {
  if (!b) return RC_ERR_ARG;
  if (first < 0 || n < 0 || first + n > b->n_blocks || (n > 0 && (!trees || !seeds))) {
    ctx_fail(b->ctx, "rc_batch_set_evolve_many: bad argument");
    return RC_ERR_ARG;
  }
  std::vector<ThrJob> jobs;
  jobs.reserve(n);
  for (int i = 0; i < n; i++) {
    const int r = set_evolve_impl(b, first + i, &trees[i], seeds[i], rng, &jobs);
    if (r != RC_OK) return r;
  }
  // the integer thresholds of all branches (16 per node, a few double operations each): 3 M of them for 10 000 ten-species blocks,
  // spread over the host's cores
  size_t total = 0;
  for (const ThrJob& j : jobs) total += j.n;
  unsigned* thr = b->evo_thr.data();
  auto work = [&](size_t j0, size_t j1) {
    for (size_t q = j0; q < j1; q++)
      for (size_t k = 0; k < jobs[q].n; k++) thr[jobs[q].off + k] = threshold_of(jobs[q].cum[k]);
  };
  const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  const size_t nt = total >= 65536 ? std::min<size_t>(hw, jobs.size()) : 1;
  if (nt <= 1) {
    work(0, jobs.size());
  } else {
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++) th.emplace_back(work, jobs.size() * t / nt, jobs.size() * (t + 1) / nt);
    for (std::thread& t : th) t.join();
  }
  return RC_OK;
}

extern "C" int rc_batch_get_sample_rows(rc_batch* b, int block, int sample, char* rows) {
  if (!b || block < 0 || block >= b->n_blocks || !rows) return RC_ERR_ARG;
  rc_ctx* ctx = b->ctx;
  const BlockDev& bd = b->blocks[block];
  if (sample < 0 || sample >= bd.n_inst - 1) return RC_ERR_ARG;
  if (!b->ran) {
    ctx_fail(ctx, "rc_batch_get_sample_rows before rc_batch_run");
    return RC_ERR_STATE;
  }
  RC_CUDA(cudaSetDevice(ctx->device));
  RC_CUDA(cudaStreamSynchronize(ctx->stream));
  RC_CUDA(cudaMemcpy(rows, b->d_raw + bd.raw_off + (size_t)sample * bd.N * bd.cols, (size_t)bd.N * bd.cols,
                     cudaMemcpyDeviceToHost));
  return RC_OK;
}

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------
extern "C" int rc_batch_upload(rc_batch* b) {
  if (!b) return RC_ERR_ARG;
  rc_ctx* ctx = b->ctx;
  RC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  size_t h2d = 0;
  b->h_scores.resize(b->scores_floats);
  b->h_nat.assign(b->nat_bytes, 0);
  for (int i = 0; i < b->n_blocks; i++) {
    const rc_block_desc& d = b->descs[i];
    const BlockDev& bd = b->blocks[i];
    const size_t rowbytes = (size_t)d.N * d.cols;
    memcpy(&b->h_nat[bd.nat_off - (long long)b->raw_bytes], d.rows, rowbytes);
    h2d += rowbytes;
    if (d.n_samples > 0 && b->evo_of_block[i] < 0) {
      if (!d.samples) {
        ctx_fail(ctx, "block " + std::to_string(i) + " has n_samples > 0 but neither samples nor rc_batch_set_evolve");
        return RC_ERR_ARG;
      }
      RC_CUDA(cudaMemcpyAsync(b->d_raw + bd.raw_off, d.samples, rowbytes * d.n_samples, cudaMemcpyHostToDevice, st));
      h2d += rowbytes * d.n_samples;
    }
    memcpy(&b->h_scores[bd.scores_off], d.scores_fwd, sizeof(float) * d.N * 4);
    memcpy(&b->h_scores[bd.scores_off + (size_t)d.N * 4], d.scores_rev, sizeof(float) * d.N * 4);
  }
  RC_CUDA(cudaMemcpyAsync(b->d_raw + b->raw_bytes, b->h_nat.data(), b->nat_bytes, cudaMemcpyHostToDevice, st));
  RC_CUDA(cudaMemcpyAsync(b->d_scores, b->h_scores.data(), sizeof(float) * b->scores_floats, cudaMemcpyHostToDevice, st));
  RC_CUDA(cudaMemcpyAsync(b->d_blocks, b->blocks.data(), sizeof(BlockDev) * b->blocks.size(), cudaMemcpyHostToDevice, st));
  if (!b->items.empty())
    RC_CUDA(cudaMemcpyAsync(b->d_items, b->items.data(), sizeof(Item) * b->items.size(), cudaMemcpyHostToDevice, st));
  if (!b->ctas.empty())
    RC_CUDA(cudaMemcpyAsync(b->d_ctas, b->ctas.data(), sizeof(CtaDesc) * b->ctas.size(), cudaMemcpyHostToDevice, st));
  RC_CUDA(cudaMemcpyAsync(b->d_tables, &b->tables, sizeof(SigmaTables), cudaMemcpyHostToDevice, st));
  RC_CUDA(cudaMemcpyAsync(b->d_ptab, &b->ptab, sizeof(PairTables), cudaMemcpyHostToDevice, st));
  if (b->d_ptab2) RC_CUDA(cudaMemcpyAsync(b->d_ptab2, &b->ptab2, sizeof(PairTables), cudaMemcpyHostToDevice, st));
  if (!b->evos.empty()) {
    void* ptrs[] = {b->d_evos, b->d_evo_nodes, b->d_evo_thr, b->d_evo_seeds, b->d_evo_seq, b->d_evo_task0};
    for (void* p : ptrs) ctx_free(ctx, p);
    // samples per task of k_evolve: eight (their generators are seeded side by side) unless the batch has so few samples that
    // tasks of eight would leave most of the GPU's warps without work
    long long total_samples = 0;
    for (const EvoDev& e : b->evos) total_samples += b->blocks[e.block].n_inst - 1;
    b->evo_spw = (int)std::max<long long>(1, std::min<long long>(EVO_SPW, total_samples / ((long long)ctx->sm_count * 32)));
    b->evo_task0.assign(1, 0);
    for (const EvoDev& e : b->evos)
      b->evo_task0.push_back(b->evo_task0.back() + (b->blocks[e.block].n_inst - 1 + b->evo_spw - 1) / b->evo_spw);
    b->d_evo_task0 = (int*)ctx_alloc(ctx, sizeof(int) * b->evo_task0.size());
    b->d_evos = (EvoDev*)ctx_alloc(ctx, sizeof(EvoDev) * b->evos.size());
    b->d_evo_nodes = (int*)ctx_alloc(ctx, sizeof(int) * b->evo_nodes.size());
    b->d_evo_thr = (unsigned*)ctx_alloc(ctx, sizeof(unsigned) * b->evo_thr.size());
    b->d_evo_seeds = (unsigned*)ctx_alloc(ctx, sizeof(unsigned) * b->evo_seeds.size());
    b->d_evo_seq = (unsigned char*)ctx_alloc(ctx, b->evo_seq_bytes);
    if (!b->d_evos || !b->d_evo_nodes || !b->d_evo_thr || !b->d_evo_seeds || !b->d_evo_seq || !b->d_evo_task0) {
      ctx_fail(ctx, "device allocation failed (evolve tables)");
      return RC_ERR_NOMEM;
    }
    RC_CUDA(cudaMemcpyAsync(b->d_evos, b->evos.data(), sizeof(EvoDev) * b->evos.size(), cudaMemcpyHostToDevice, st));
    RC_CUDA(cudaMemcpyAsync(b->d_evo_nodes, b->evo_nodes.data(), sizeof(int) * b->evo_nodes.size(), cudaMemcpyHostToDevice, st));
    RC_CUDA(cudaMemcpyAsync(b->d_evo_thr, b->evo_thr.data(), sizeof(unsigned) * b->evo_thr.size(), cudaMemcpyHostToDevice, st));
    RC_CUDA(cudaMemcpyAsync(b->d_evo_seeds, b->evo_seeds.data(), sizeof(unsigned) * b->evo_seeds.size(), cudaMemcpyHostToDevice, st));
    RC_CUDA(cudaMemcpyAsync(b->d_evo_task0, b->evo_task0.data(), sizeof(int) * b->evo_task0.size(), cudaMemcpyHostToDevice, st));
    h2d += sizeof(EvoDev) * b->evos.size() + sizeof(int) * b->evo_nodes.size() +
           sizeof(unsigned) * (b->evo_thr.size() + b->evo_seeds.size());
  }
  h2d += sizeof(float) * b->scores_floats + sizeof(BlockDev) * b->blocks.size() + sizeof(Item) * b->items.size() +
         sizeof(CtaDesc) * b->ctas.size() + sizeof(SigmaTables);
  // descriptors may go away after this call returns
  RC_CUDA(cudaStreamSynchronize(st));
  b->stats.h2d_bytes = h2d;
  b->uploaded = true;
  b->ran = false;
  b->downloaded = false;
  return RC_OK;
}

// ------------------------------------------------------------------------------------------------
// run
// ------------------------------------------------------------------------------------------------
static int ev_begin(rc_batch* b, int stage) {
  EventPair ep;
  ep.stage = stage;
  if (cudaEventCreate(&ep.a) != cudaSuccess || cudaEventCreate(&ep.b) != cudaSuccess) return -1;
  cudaEventRecord(ep.a, b->ctx->stream);
  b->events.push_back(ep);
  return (int)b->events.size() - 1;
}
static void ev_end(rc_batch* b, int idx) {
  if (idx >= 0) cudaEventRecord(b->events[idx].b, b->ctx->stream);
}

template <int R, bool DENSE>
static int launch_dp(rc_batch* b, const CtaDesc* d_ctas, size_t ncta, int maxNK, int maxZs) {
  rc_ctx* ctx = b->ctx;
  if (ncta == 0) return RC_OK;
  int nst = 2;
  size_t per_warp = DpSmem<R>::per_warp(maxNK, maxZs, nst);
  if (per_warp > (size_t)ctx->smem_optin) {  // the reference's maximum of 500 rows only fits with a single-stage ring
    nst = 1;
    per_warp = DpSmem<R>::per_warp(maxNK, maxZs, nst);
  }
  const int nw = (int)std::min<size_t>(DP_WARPS, (size_t)ctx->smem_optin / per_warp);
  if (nw < 1) {
    ctx_fail(ctx, "alignment has too many rows for the shared-memory resident DP state (N-1 = " + std::to_string(maxNK) + ")");
    return RC_ERR_ARG;
  }
  const size_t smem = nw * per_warp;
  RC_CUDA(allow_max_smem(ctx, k_dp<R, DENSE>));
  k_dp<R, DENSE><<<(unsigned)ncta, nw * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma, b->d_z,
                                                                      b->d_recs, b->d_dense, b->prm, (int)ctx->band_slots,
                                                                      maxNK, maxZs, nst);
  RC_CUDA(cudaGetLastError());
  b->stats.launches++;
  b->stats.dp_launches++;
  return RC_OK;
}

template <int NK>
static int launch_dp_reg_nk(rc_batch* b, const CtaDesc* d_ctas, size_t ncta) {
  rc_ctx* ctx = b->ctx;
  // k_dp_regtu: the three-addition form with deferred S1 / S2 bookkeeping, chosen per batch (rc_batch_create)
  if (b->use_tu && NK <= 12)
    k_dp_regtu<(NK <= 12 ? NK : 1)><<<(unsigned)ncta, DP_WARPS * 32, 0, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma,
                                                                                     b->d_recs, b->prm, (int)ctx->band_slots);
  else
    k_dp_reg<NK><<<(unsigned)ncta, DP_WARPS * 32, 0, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma, b->d_recs,
                                                                   b->prm, (int)ctx->band_slots);
  RC_CUDA(cudaGetLastError());
  b->stats.launches++;
  b->stats.dp_launches++;
  return RC_OK;
}

static int launch_dp_reg(rc_batch* b, int NK, const CtaDesc* d_ctas, size_t ncta) {
  if (ncta == 0) return RC_OK;
  switch (NK) {
    case 1: return launch_dp_reg_nk<1>(b, d_ctas, ncta);
    case 2: return launch_dp_reg_nk<2>(b, d_ctas, ncta);
    case 3: return launch_dp_reg_nk<3>(b, d_ctas, ncta);
    case 4: return launch_dp_reg_nk<4>(b, d_ctas, ncta);
    case 5: return launch_dp_reg_nk<5>(b, d_ctas, ncta);
    case 6: return launch_dp_reg_nk<6>(b, d_ctas, ncta);
    case 7: return launch_dp_reg_nk<7>(b, d_ctas, ncta);
    case 8: return launch_dp_reg_nk<8>(b, d_ctas, ncta);
    case 9: return launch_dp_reg_nk<9>(b, d_ctas, ncta);
    case 10: return launch_dp_reg_nk<10>(b, d_ctas, ncta);
    case 11: return launch_dp_reg_nk<11>(b, d_ctas, ncta);
    case 12: return launch_dp_reg_nk<12>(b, d_ctas, ncta);
    case 13: return launch_dp_reg_nk<13>(b, d_ctas, ncta);
    case 14: return launch_dp_reg_nk<14>(b, d_ctas, ncta);
    case 15: return launch_dp_reg_nk<15>(b, d_ctas, ncta);
    case 16: return launch_dp_reg_nk<16>(b, d_ctas, ncta);
    default: ctx_fail(b->ctx, "internal: k_dp_reg NK out of range"); return RC_ERR_STATE;
  }
}

// Warps per CTA of k_dp_smp: the CTA's sigma table decides how many CTAs fit an SM; with few of them, more warps
// share each table (the launch that owns the getHSS digest also needs 2 KB of fold state per warp).
static int smp_warps(rc_ctx* ctx, size_t smem_table, bool with_fold) {
  if (ctx->smp_warps_forced >= 1 && ctx->smp_warps_forced <= SMP_MAX_WARPS) return (int)ctx->smp_warps_forced;  // experiments
  int best = SMP_WARPS, best_warps = 0;
  for (int nw = SMP_WARPS; nw <= SMP_MAX_WARPS; nw += 2) {
    const size_t per_cta = smem_table + (with_fold ? (size_t)nw * 64 * sizeof(RowRec) : 0) + 1024;
    if (per_cta > (size_t)ctx->smem_optin + 1024) break;
    const int ctas = (int)std::min<size_t>(16 / nw, ((size_t)228 * 1024) / per_cta);  // 128 registers per thread: 16 warps per SM
    if (ctas * nw > best_warps) {
      best_warps = ctas * nw;
      best = nw;
    }
  }
  return best;
}

// layout 5: one launch per species chunk (1-3 quads each), chunk g continuing the partial sums of chunk g-1
template <int NK>
static int launch_dp_smpc_nk(rc_batch* b, int chunk, bool last, bool seg, const CtaDesc* d_ctas, size_t ncta, size_t smem) {
  rc_ctx* ctx = b->ctx;
  if (!seg) smem += SMP_PF_BYTES;  // prefetch ring of the incoming partial sums (k_dp_smp<., true>)
  const int nw = smp_warps(ctx, smem, last);
  if (last) smem += (size_t)nw * 64 * sizeof(RowRec);
  if (seg) {
    RC_CUDA(allow_max_smem(ctx, k_dp_smps<NK, true>));
    k_dp_smps<NK, true><<<(unsigned)ncta, nw * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma, b->d_z,
                                                                             b->d_recs, b->prm, (int)ctx->band_slots, chunk,
                                                                             b->d_partial);
  } else {
    if (last) {
      RC_CUDA(allow_max_smem(ctx, k_dp_smp<NK, true, true>));
      k_dp_smp<NK, true, true><<<(unsigned)ncta, nw * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma, b->d_z,
                                                                                    b->d_recs, b->prm, (int)ctx->band_slots, chunk,
                                                                                    b->d_partial);
    } else {
      RC_CUDA(allow_max_smem(ctx, k_dp_smp<NK, true, false>));
      k_dp_smp<NK, true, false><<<(unsigned)ncta, nw * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma, b->d_z,
                                                                                     b->d_recs, b->prm, (int)ctx->band_slots, chunk,
                                                                                     b->d_partial);
    }
  }
  RC_CUDA(cudaGetLastError());
  b->stats.launches++;
  b->stats.dp_launches++;
  return RC_OK;
}

static int launch_dp_smpc(rc_batch* b, int Q, bool seg, const CtaDesc* d_ctas, size_t ncta, size_t smem) {
  if (ncta == 0) return RC_OK;
  const int W = (Q + 2) / 3, base = Q / W, rem = Q % W;
  for (int g = 0; g < W; g++) {
    const int quads = base + (g < rem ? 1 : 0);
    int rc;
    switch (quads) {
      case 1: rc = launch_dp_smpc_nk<4>(b, g, g == W - 1, seg, d_ctas, ncta, smem); break;
      case 2: rc = launch_dp_smpc_nk<8>(b, g, g == W - 1, seg, d_ctas, ncta, smem); break;
      case 3: rc = launch_dp_smpc_nk<12>(b, g, g == W - 1, seg, d_ctas, ncta, smem); break;
      default: ctx_fail(b->ctx, "internal: k_dp_smp chunk size out of range"); return RC_ERR_STATE;
    }
    if (rc != RC_OK) return rc;
  }
  return RC_OK;
}

template <int NK>
static int launch_dp_smp_nk(rc_batch* b, bool seg, const CtaDesc* d_ctas, size_t ncta, size_t smem) {
  rc_ctx* ctx = b->ctx;
  const int nw = smp_warps(ctx, smem, true);
  smem += (size_t)nw * 64 * sizeof(RowRec);
  if (seg) {
    RC_CUDA(allow_max_smem(ctx, k_dp_smps<NK, false>));
    k_dp_smps<NK, false><<<(unsigned)ncta, nw * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma, b->d_z,
                                                                              b->d_recs, b->prm, (int)ctx->band_slots, 0, nullptr);
  } else {
    RC_CUDA(allow_max_smem(ctx, k_dp_smp<NK, false>));
    k_dp_smp<NK, false><<<(unsigned)ncta, nw * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma, b->d_z,
                                                                           b->d_recs, b->prm, (int)ctx->band_slots, 0, nullptr);
  }
  RC_CUDA(cudaGetLastError());
  b->stats.launches++;
  b->stats.dp_launches++;
  return RC_OK;
}

static int launch_dp_smp(rc_batch* b, int NK, bool seg, const CtaDesc* d_ctas, size_t ncta, size_t smem) {
  if (ncta == 0) return RC_OK;
  switch (NK) {
    case 1: return launch_dp_smp_nk<1>(b, seg, d_ctas, ncta, smem);
    case 2: return launch_dp_smp_nk<2>(b, seg, d_ctas, ncta, smem);
    case 3: return launch_dp_smp_nk<3>(b, seg, d_ctas, ncta, smem);
    case 4: return launch_dp_smp_nk<4>(b, seg, d_ctas, ncta, smem);
    case 5: return launch_dp_smp_nk<5>(b, seg, d_ctas, ncta, smem);
    case 6: return launch_dp_smp_nk<6>(b, seg, d_ctas, ncta, smem);
    case 7: return launch_dp_smp_nk<7>(b, seg, d_ctas, ncta, smem);
    case 8: return launch_dp_smp_nk<8>(b, seg, d_ctas, ncta, smem);
    case 9: return launch_dp_smp_nk<9>(b, seg, d_ctas, ncta, smem);
    case 10: return launch_dp_smp_nk<10>(b, seg, d_ctas, ncta, smem);
    case 11: return launch_dp_smp_nk<11>(b, seg, d_ctas, ncta, smem);
    case 12: return launch_dp_smp_nk<12>(b, seg, d_ctas, ncta, smem);
    case 13: return launch_dp_smp_nk<13>(b, seg, d_ctas, ncta, smem);
    case 14: return launch_dp_smp_nk<14>(b, seg, d_ctas, ncta, smem);
    case 15: return launch_dp_smp_nk<15>(b, seg, d_ctas, ncta, smem);
    case 16: return launch_dp_smp_nk<16>(b, seg, d_ctas, ncta, smem);
    default: ctx_fail(b->ctx, "internal: k_dp_smp NK out of range"); return RC_ERR_STATE;
  }
}

// k_dp_smpf: the fused kernels (sigma table built in shared memory by the DP CTA itself)
template <int NK, bool CHAINED>
static int launch_dp_smpf_nk(rc_batch* b, int chunk, bool last, const CtaDesc* d_ctas, size_t ncta, size_t smem) {
  rc_ctx* ctx = b->ctx;
  (void)last;
  const int nw = smp_warps(ctx, smem, true);
  smem += (size_t)nw * 64 * sizeof(RowRec);  // the carve-up always has the records at the end (SmpfCfg::off_rec)
  RC_CUDA(allow_max_smem(ctx, k_dp_smpf<NK, CHAINED>));
  k_dp_smpf<NK, CHAINED><<<(unsigned)ncta, nw * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_p2, b->d_p2f, b->d_cls,
                                                                         b->d_cols0, b->d_scores, b->d_ptab2, b->d_z, b->d_recs,
                                                                         b->prm, (int)ctx->band_slots, chunk, b->d_partial);
  RC_CUDA(cudaGetLastError());
  b->stats.launches++;
  b->stats.dp_launches++;
  return RC_OK;
}

static int launch_dp_smpf(rc_batch* b, int NK, const CtaDesc* d_ctas, size_t ncta, size_t smem) {
  if (ncta == 0) return RC_OK;
  switch (NK) {
    case 1: return launch_dp_smpf_nk<1, false>(b, 0, true, d_ctas, ncta, smem);
    case 2: return launch_dp_smpf_nk<2, false>(b, 0, true, d_ctas, ncta, smem);
    case 3: return launch_dp_smpf_nk<3, false>(b, 0, true, d_ctas, ncta, smem);
    case 4: return launch_dp_smpf_nk<4, false>(b, 0, true, d_ctas, ncta, smem);
    case 5: return launch_dp_smpf_nk<5, false>(b, 0, true, d_ctas, ncta, smem);
    case 6: return launch_dp_smpf_nk<6, false>(b, 0, true, d_ctas, ncta, smem);
    case 7: return launch_dp_smpf_nk<7, false>(b, 0, true, d_ctas, ncta, smem);
    case 8: return launch_dp_smpf_nk<8, false>(b, 0, true, d_ctas, ncta, smem);
    case 9: return launch_dp_smpf_nk<9, false>(b, 0, true, d_ctas, ncta, smem);
    case 10: return launch_dp_smpf_nk<10, false>(b, 0, true, d_ctas, ncta, smem);
    case 11: return launch_dp_smpf_nk<11, false>(b, 0, true, d_ctas, ncta, smem);
    case 12: return launch_dp_smpf_nk<12, false>(b, 0, true, d_ctas, ncta, smem);
    case 13: return launch_dp_smpf_nk<13, false>(b, 0, true, d_ctas, ncta, smem);
    case 14: return launch_dp_smpf_nk<14, false>(b, 0, true, d_ctas, ncta, smem);
    case 15: return launch_dp_smpf_nk<15, false>(b, 0, true, d_ctas, ncta, smem);
    case 16: return launch_dp_smpf_nk<16, false>(b, 0, true, d_ctas, ncta, smem);
    default: ctx_fail(b->ctx, "internal: k_dp_smpf NK out of range"); return RC_ERR_STATE;
  }
}

// layout 5, fused: one launch per species chunk (1-3 quads each), as launch_dp_smpc
static int launch_dp_smpcf(rc_batch* b, int Q, const CtaDesc* d_ctas, size_t ncta, size_t smem) {
  if (ncta == 0) return RC_OK;
  const int W = (Q + 2) / 3, base = Q / W, rem = Q % W;
  for (int g = 0; g < W; g++) {
    const int quads = base + (g < rem ? 1 : 0);
    int rc;
    switch (quads) {
      case 1: rc = launch_dp_smpf_nk<4, true>(b, g, g == W - 1, d_ctas, ncta, smem); break;
      case 2: rc = launch_dp_smpf_nk<8, true>(b, g, g == W - 1, d_ctas, ncta, smem); break;
      case 3: rc = launch_dp_smpf_nk<12, true>(b, g, g == W - 1, d_ctas, ncta, smem); break;
      default: ctx_fail(b->ctx, "internal: k_dp_smpf chunk size out of range"); return RC_ERR_STATE;
    }
    if (rc != RC_OK) return rc;
  }
  return RC_OK;
}

template <int NKW>
static int launch_dp_chain_nk(rc_batch* b, int W, const CtaDesc* d_ctas, size_t ncta) {
  rc_ctx* ctx = b->ctx;
  // W chunks in chain_passes(W) launches of nearly equal width; each continues the partial sums of the one before
  const int np = chain_passes(W), base = W / np, rem = W % np;
  int c_lo = 0;
  for (int pass = 0; pass < np; pass++) {
    const int wp = base + (pass < rem ? 1 : 0);
    const size_t smem = ChainCfg<NKW>::smem_bytes(wp);
    if (smem > (size_t)ctx->smem_optin) {
      ctx_fail(ctx, "internal: k_dp_chain shared memory exceeds the device limit");
      return RC_ERR_STATE;
    }
#define RC_LAUNCH_CHAIN(...)                                                                                              \
  do {                                                                                                                    \
    RC_CUDA(allow_max_smem(ctx, k_dp_chain<__VA_ARGS__>));                                                                \
    k_dp_chain<__VA_ARGS__><<<(unsigned)ncta, wp * 32, smem, ctx->stream>>>(b->d_blocks, b->d_items, d_ctas, b->d_sigma,  \
                                                                            b->d_recs, b->prm, (int)ctx->band_slots, c_lo, \
                                                                            np == 1 ? nullptr : b->d_partial);            \
  } while (0)
    if (np == 1) RC_LAUNCH_CHAIN(NKW, false, true, true);
    else if (pass == 0) RC_LAUNCH_CHAIN(NKW, true, true, false);
    else if (pass == np - 1) RC_LAUNCH_CHAIN(NKW, true, false, true);
    else RC_LAUNCH_CHAIN(NKW, true, false, false);
#undef RC_LAUNCH_CHAIN
    RC_CUDA(cudaGetLastError());
    b->stats.launches++;
    b->stats.dp_launches++;
    c_lo += wp;
  }
  return RC_OK;
}

static int launch_dp_chain(rc_batch* b, int NKW, int W, const CtaDesc* d_ctas, size_t ncta) {
  if (ncta == 0) return RC_OK;
  switch (NKW) {
    case 5: return launch_dp_chain_nk<5>(b, W, d_ctas, ncta);
    case 6: return launch_dp_chain_nk<6>(b, W, d_ctas, ncta);
    case 7: return launch_dp_chain_nk<7>(b, W, d_ctas, ncta);
    case 8: return launch_dp_chain_nk<8>(b, W, d_ctas, ncta);
    case 9: return launch_dp_chain_nk<9>(b, W, d_ctas, ncta);
    case 10: return launch_dp_chain_nk<10>(b, W, d_ctas, ncta);
    case 11: return launch_dp_chain_nk<11>(b, W, d_ctas, ncta);
    case 12: return launch_dp_chain_nk<12>(b, W, d_ctas, ncta);
    default: ctx_fail(b->ctx, "internal: k_dp_chain NKW out of range"); return RC_ERR_STATE;
  }
}

// Dense (exact fallback) scoring of a list of single-instance items.  Always runs k_dp<1, DENSE> on
// layout-0 sigma / z tiles that are rebuilt for the item in scratch, whatever layout the block uses.
static int run_dense_items(rc_batch* b, const std::vector<Item>& src_items) {
  rc_ctx* ctx = b->ctx;
  cudaStream_t st = ctx->stream;
  BlockDev* d_blk = nullptr;
  Item* d_item = nullptr;
  CtaDesc* d_cta = nullptr;
  unsigned* d_zs = nullptr;
  size_t cta_cap = 0, zs_cap = 0;
  int rcode = RC_OK;
  auto cleanup = [&]() { cudaFree(d_blk); cudaFree(d_item); cudaFree(d_cta); cudaFree(d_zs); };
#define RC_CUDA_D(call)                                                         \
  do {                                                                          \
    cudaError_t _e = (call);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ctx_fail(ctx, std::string(#call) + ": " + cudaGetErrorString(_e));        \
      cleanup();                                                                \
      return RC_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)
  RC_CUDA_D(cudaMalloc((void**)&d_blk, sizeof(BlockDev)));
  RC_CUDA_D(cudaMalloc((void**)&d_item, sizeof(Item)));
  for (const Item& src : src_items) {
    BlockDev bd = b->blocks[src.block];
    set_layout(bd, 0);
    for (int f = 0; f < 3; f++) bd.ntiles[f] = (bd.sites[f] + TILE - 1) / TILE;  // without the padding of layout 1
    size_t zw = 0;
    for (int s = 0; s < 2; s++)
      for (int f = 0; f < 3; f++) {
        bd.z_off[s][f] = (long long)zw;
        zw += (size_t)bd.ntiles[f] * bd.zstride;
      }
    Item it = src;
    it.block = 0;
    size_t sig = 0, dn = 0;
    for (int s = 0; s < 2; s++)
      for (int f = 0; f < 3; f++) {
        it.sigma_off[s][f] = (long long)sig;
        sig += sigma_floats_sf(bd, f, it.ninst);
        it.rec_off[s][f] = 0;
        it.dense_off[s][f] = (long long)dn;
        dn += (size_t)it.ninst * ((size_t)bd.sites[f] * (bd.sites[f] + 1) / 2);
      }
    if (sig > b->sigma_floats) {  // layout 0 tiles can be larger than the padded layout-1 ones never; guard anyway
      ctx_free(ctx, b->d_sigma);
      b->d_sigma = (float*)ctx_alloc(ctx, sig * sizeof(float));
      if (!b->d_sigma) { cleanup(); ctx_fail(ctx, "device allocation failed (sigma scratch)"); return RC_ERR_NOMEM; }
      b->sigma_floats = sig;
    }
    if (dn > b->dense_floats) {
      ctx_free(ctx, b->d_dense);
      b->dense_floats = 0;
      b->d_dense = (float*)ctx_alloc(ctx, dn * sizeof(float));
      if (!b->d_dense) { cleanup(); ctx_fail(ctx, "device allocation failed (dense S scratch)"); return RC_ERR_NOMEM; }
      b->dense_floats = dn;
    }
    if (zw > zs_cap) {
      cudaFree(d_zs);
      d_zs = nullptr;
      RC_CUDA_D(cudaMalloc((void**)&d_zs, std::max<size_t>(zw * sizeof(unsigned), 256)));
      zs_cap = zw;
    }
    std::vector<BlockDev> one_blk{bd};
    std::vector<Item> one{it};
    std::vector<CtaDesc> ctas;
    build_ctas(one_blk, one, 0, 1, 1, -1, ctas);
    if (ctas.size() > cta_cap) {
      cudaFree(d_cta);
      d_cta = nullptr;
      RC_CUDA_D(cudaMalloc((void**)&d_cta, sizeof(CtaDesc) * ctas.size()));
      cta_cap = ctas.size();
    }
    RC_CUDA_D(cudaMemcpyAsync(d_blk, &bd, sizeof(BlockDev), cudaMemcpyHostToDevice, st));
    RC_CUDA_D(cudaMemcpyAsync(d_item, &it, sizeof(Item), cudaMemcpyHostToDevice, st));
    RC_CUDA_D(cudaMemcpyAsync(d_cta, ctas.data(), sizeof(CtaDesc) * ctas.size(), cudaMemcpyHostToDevice, st));
    k_prep<0><<<1, 256, 0, st>>>(d_blk, b->d_cls, b->d_cols0, d_zs);
    RC_CUDA_D(cudaGetLastError());
    const long long work = (long long)it.ninst * 2 * (bd.L - 2);
    dim3 gs(1, (unsigned)std::min<long long>((work + 255) / 256, 4096));
    k_sigma<<<gs, 256, 0, st>>>(d_blk, d_item, b->d_cls, b->d_cols0, b->d_scores, b->d_tables, d_zs, b->d_sigma, b->prm, 0);
    RC_CUDA_D(cudaGetLastError());
    // the DP kernel reads blocks / items / z through the batch pointers: point them at the private copies
    BlockDev* saved_blocks = b->d_blocks;
    Item* saved_items = b->d_items;
    unsigned* saved_z = b->d_z;
    b->d_blocks = d_blk;
    b->d_items = d_item;
    b->d_z = d_zs;
    rcode = launch_dp<1, true>(b, d_cta, ctas.size(), bd.NK, bd.zstride);
    b->d_blocks = saved_blocks;
    b->d_items = saved_items;
    b->d_z = saved_z;
    if (rcode != RC_OK) {
      cleanup();
      return rcode;
    }
    dim3 gh(1, (unsigned)((it.ninst * 6 * 32 + 127) / 128));
    k_hss_dense<<<gh, 128, 0, st>>>(d_blk, d_item, b->d_dense, b->d_res, b->d_hss, b->d_hsscnt);
    RC_CUDA_D(cudaGetLastError());
    b->stats.launches += 3;
    RC_CUDA_D(cudaStreamSynchronize(st));  // the private descriptors are rewritten for the next item
    b->stats.dense_fallbacks += it.ninst;
  }
  cleanup();
#undef RC_CUDA_D
  return RC_OK;
}

extern "C" int rc_batch_run(rc_batch* b) {
  if (!b) return RC_ERR_ARG;
  rc_ctx* ctx = b->ctx;
  if (!b->uploaded) {
    ctx_fail(ctx, "rc_batch_run before rc_batch_upload");
    return RC_ERR_STATE;
  }
  RC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  for (auto& e : b->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  b->events.clear();
  b->stats.launches = 0;
  b->stats.dp_launches = 0;
  b->stats.dense_fallbacks = 0;

  int maxchunks = 1;
  for (const BlockDev& bd : b->blocks) maxchunks = std::max(maxchunks, (int)(((size_t)bd.inst_stride >> 4) * bd.n_inst / 256 + 1));
  int ev = ev_begin(b, 0);
  if (!b->evos.empty()) {
    const int total_tasks = b->evo_task0.back();
    const int want = (total_tasks + EVO_WARPS - 1) / EVO_WARPS;
    const int gx = std::max(1, std::min(want, ctx->sm_count * 16));  // persistent warps: 16 CTAs of 4 warps per SM at most
    const size_t need = (size_t)gx * EVO_WARPS * b->evo_spw * 624 * sizeof(unsigned);
    if (need > b->evo_mt_bytes) {
      ctx_free(ctx, b->d_evo_mt);
      b->d_evo_mt = (unsigned*)ctx_alloc(ctx, need);
      if (!b->d_evo_mt) { ctx_fail(ctx, "device allocation failed (generator states)"); return RC_ERR_NOMEM; }
      b->evo_mt_bytes = need;
    }
    k_evolve<<<gx, EVO_WARPS * 32, 0, st>>>(b->d_blocks, b->d_evos, b->d_evo_task0, (int)b->evos.size(), b->d_evo_nodes,
                                            b->d_evo_thr, b->d_evo_seeds, b->d_evo_seq, b->d_raw, b->d_evo_mt, b->evo_spw);
    RC_CUDA(cudaGetLastError());
    b->stats.launches++;
  }
  {
    dim3 g((unsigned)b->n_blocks, (unsigned)std::min(maxchunks, 2048));
    const int evk = ev_begin(b, 4);  // k_pack alone (HBM roofline of subsystem (a))
    k_pack<<<g, 256, 0, st>>>(b->d_blocks, b->d_raw, b->d_cls, ctx->d_lut);
    RC_CUDA(cudaGetLastError());
    ev_end(b, evk);
    k_prep<1><<<b->n_blocks, 256, 0, st>>>(b->d_blocks, b->d_cls, b->d_cols0, b->d_z);
    RC_CUDA(cudaGetLastError());
    if (b->p2_words > 0) {  // packed rows for the blocks whose DP kernel builds its own sigma table (needs cols0)
      RC_CUDA(cudaMemsetAsync(b->d_p2f, 0, sizeof(unsigned) * b->p2f_words, st));
      dim3 g2((unsigned)b->n_blocks, (unsigned)std::min(64, (b->max_n_inst + 31) / 32));
      const int max_L = (b->max_fused_cols + 3) / 4 * 4;
      const int pitch_max = b->max_fused_cols + 16;
      const int stage = std::max(32 * pitch_max, std::min(P2_STAGE_BYTES, 32 * pitch_max * b->max_fused_N));
      const size_t p2_smem = (size_t)2 * max_L * sizeof(int) + (size_t)stage;
      RC_CUDA(allow_max_smem(ctx, k_pack2));
      k_pack2<<<g2, 256, p2_smem, st>>>(b->d_blocks, b->d_cls, b->d_cols0, b->d_p2, b->d_p2f, max_L, stage);
      RC_CUDA(cudaGetLastError());
      b->stats.launches++;
    }
    // z words: a long block would keep a single CTA busy for ~0.2 ms; spread each block over several CTAs
    size_t maxz = 1;
    for (const BlockDev& bd : b->blocks) maxz = std::max(maxz, (size_t)bd.ntiles[0] * bd.zstride);
    dim3 gz((unsigned)b->blocks.size(), (unsigned)std::min<size_t>((maxz + 255) / 256, 64));  // tail blocks have z words of their own
    k_prep<2><<<gz, 256, 0, st>>>(b->d_blocks, b->d_cls, b->d_cols0, b->d_z);
    RC_CUDA(cudaGetLastError());
    b->stats.launches += 3;
  }
  ev_end(b, ev);
  RC_CUDA(cudaMemsetAsync(b->d_ovf, 0, sizeof(int), st));
  // results of blocks that are never scored (L < 3): no HSS, every sample maximum -1
  RC_CUDA(cudaMemsetAsync(b->d_hsscnt, 0, sizeof(int) * b->hsscnt_ints, st));

  if (ctx->force_dense) {
    // test hook: everything through the exact dense path, one instance at a time
    std::vector<Item> singles;
    for (const Item& it : b->items)
      for (int i = 0; i < it.ninst; i++) {
        Item s = it;
        s.inst0 = it.inst0 + i;
        s.ninst = 1;
        singles.push_back(s);
      }
    int rcode = run_dense_items(b, singles);
    if (rcode != RC_OK) return rcode;
    b->ran = true;
    b->downloaded = false;
    return RC_OK;
  }

  for (const Chunk& ch : b->chunks) {
    ev = ev_begin(b, 1);
    {
      dim3 g((unsigned)ch.nitems, (unsigned)std::min<long long>((ch.max_sigma_work + 255) / 256, 8192));
      if (ch.n_layout[3] > 0 && !ctx->no_sig_rows3) {  // layout 3: whole step rows per thread
        k_sigma_rows3<<<g, 256, 0, st>>>(b->d_blocks, b->d_items + ch.item0, b->d_cls, b->d_cols0, b->d_scores, b->d_ptab, b->d_z,
                                         b->d_sigma);
        RC_CUDA(cudaGetLastError());
        b->stats.launches++;
      }
      if (ch.n_layout[0] + (ctx->no_sig_rows3 ? ch.n_layout[3] : 0) > 0) {
        k_sigma<<<g, 256, 0, st>>>(b->d_blocks, b->d_items + ch.item0, b->d_cls, b->d_cols0, b->d_scores, b->d_tables,
                                   b->d_z, b->d_sigma, b->prm, ctx->no_sig_rows3 ? 0 : 1);
        RC_CUDA(cudaGetLastError());
        b->stats.launches++;
      }
      if (ch.n_layout[1] > 0) {
        k_sigma_rows<<<g, 256, 0, st>>>(b->d_blocks, b->d_items + ch.item0, b->d_cls, b->d_cols0, b->d_scores, b->d_ptab,
                                        b->d_z, b->d_sigma, b->prm);
        RC_CUDA(cudaGetLastError());
        b->stats.launches++;
      }
      if (ch.n_sig_p2 > 0) {  // sample-major sigma tables built from the packed rows
        dim3 g3((unsigned)ch.nitems, (unsigned)((ch.max_ninst + 31) / 32) * 2, (unsigned)ch.max_p2_chunks);
        RC_CUDA(allow_max_smem(ctx, k_sigma_p2));
        k_sigma_p2<<<g3, 256, ch.max_p2_smem, st>>>(b->d_blocks, b->d_items + ch.item0, b->d_p2, b->d_p2f, b->d_cls, b->d_cols0,
                                                    b->d_scores, b->d_ptab2, b->d_sigma);
        RC_CUDA(cudaGetLastError());
        b->stats.launches++;
      }
      if (ch.n_smp_unfused > ch.n_sig_p2) {  // ... and those built from class bytes
        const int nqz = std::min(16, std::max(1, ch.max_smp_quads / 3));
        const int npc = (ch.max_smp_npos + SIG_PCH - 1) / SIG_PCH;
        dim3 g2((unsigned)ch.nitems, (unsigned)((ch.max_ninst + 31) / 32), (unsigned)(npc * 2 * nqz));
        constexpr int SIGMA_SMP_DYN_SMEM = 4 * 32 * SIG_PITCH;  // species staging; static + dynamic exceed 48 KB
        RC_CUDA(allow_max_smem(ctx, k_sigma_smp));
        k_sigma_smp<<<g2, 256, SIGMA_SMP_DYN_SMEM, st>>>(b->d_blocks, b->d_items + ch.item0, b->d_cls, b->d_cols0, b->d_scores,
                                                         b->d_ptab, b->d_sigma, nqz);
        RC_CUDA(cudaGetLastError());
        b->stats.launches++;
      }
    }
    ev_end(b, ev);
    ev = ev_begin(b, 2);
    {
      // CtaDesc.item indexes the batch-wide item array
      for (int cl = 0; cl < N_CLASSES; cl++) {
        if (ch.ncta[cl] == 0) continue;
        int rcode;
        if (cl >= SMPCF_CLASS0)
          rcode = launch_dp_smpcf(b, SMPC_Q_MIN + (cl - SMPCF_CLASS0), b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.max_smpf_smem);
        else if (cl >= SMPF_CLASS0)
          rcode = launch_dp_smpf(b, cl - SMPF_CLASS0 + 1, b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.max_smpf_smem);
        else if (cl >= SMPCS_CLASS0)
          rcode = launch_dp_smpc(b, SMPC_Q_MIN + (cl - SMPCS_CLASS0), true, b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.max_smps_smem);
        else if (cl >= SMPS_CLASS0)
          rcode = launch_dp_smp(b, cl - SMPS_CLASS0 + 1, true, b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.max_smps_smem);
        else if (cl >= SMPC_CLASS0)
          rcode = launch_dp_smpc(b, SMPC_Q_MIN + (cl - SMPC_CLASS0), false, b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.max_smp_smem);
        else if (cl >= CHAIN_CLASS0)
          rcode = launch_dp_chain(b, CHAIN_NKW_MIN + (cl - CHAIN_CLASS0) % CHAIN_NKW_SPAN, 2 + (cl - CHAIN_CLASS0) / CHAIN_NKW_SPAN,
                                  b->d_ctas + ch.cta0[cl], ch.ncta[cl]);
        else if (cl >= SMP_CLASS0) rcode = launch_dp_smp(b, cl - SMP_CLASS0 + 1, false, b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.max_smp_smem);
        else if (cl < REG_MAX_NK) rcode = launch_dp_reg(b, cl + 1, b->d_ctas + ch.cta0[cl], ch.ncta[cl]);
        else if (cl == REG_MAX_NK) rcode = launch_dp<2, false>(b, b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.maxNK[cl], ch.maxZs[cl]);
        else rcode = launch_dp<1, false>(b, b->d_ctas + ch.cta0[cl], ch.ncta[cl], ch.maxNK[cl], ch.maxZs[cl]);
        if (rcode != RC_OK) return rcode;
      }
    }
    ev_end(b, ev);
    ev = ev_begin(b, 3);
    {
      // long frames: one warp per (instance, strand, frame); short frames: one thread (each kernel skips the other's items)
      if (ch.hss_warp_items > 0) {
        dim3 g((unsigned)ch.nitems, (unsigned)((ch.max_ninst * 6 + HSS_WARPS - 1) / HSS_WARPS));
        k_hss<<<g, HSS_WARPS * 32, 0, st>>>(b->d_blocks, b->d_items + ch.item0, b->d_recs, b->d_res, b->d_hss, b->d_hsscnt, b->d_ovf);
        RC_CUDA(cudaGetLastError());
        b->stats.launches++;
      }
      if (ch.hss_warp_items < (int)ch.nitems) {
        dim3 g((unsigned)ch.nitems, (unsigned)((ch.max_ninst * 6 + 127) / 128));
        k_hss_thr<<<g, 128, 0, st>>>(b->d_blocks, b->d_items + ch.item0, b->d_recs, b->d_res, b->d_hss, b->d_hsscnt, b->d_ovf);
        RC_CUDA(cudaGetLastError());
        b->stats.launches++;
      }
    }
    ev_end(b, ev);
  }

  // band overflow -> exact dense re-scoring of the affected alignments (rare)
  int ovf = 0;
  RC_CUDA(cudaMemcpyAsync(&ovf, b->d_ovf, sizeof(int), cudaMemcpyDeviceToHost, st));
  RC_CUDA(cudaStreamSynchronize(st));
  if (ovf > 0) {
    RC_CUDA(cudaMemcpy(b->h_res.data(), b->d_res, sizeof(float) * b->res_floats, cudaMemcpyDeviceToHost));
    std::vector<Item> singles;
    for (int i = 0; i < b->n_blocks; i++) {
      const BlockDev& bd = b->blocks[i];
      if (bd.L < 3) continue;
      for (int inst = 0; inst < bd.n_inst; inst++) {
        bool hit = false;
        for (int sf = 0; sf < 6; sf++)
          if (bd.sites[sf % 3] > 0) hit |= b->h_res[bd.res_off + (size_t)inst * 6 + sf] == -2.0f;
        if (hit) {
          Item s;
          memset(&s, 0, sizeof(s));
          s.block = i;
          s.inst0 = inst;
          s.ninst = 1;
          singles.push_back(s);
        }
      }
    }
    int rcode = run_dense_items(b, singles);
    if (rcode != RC_OK) return rcode;
  }
  // stage timings
  b->stats.ms_pack = b->stats.ms_sigma = b->stats.ms_dp = b->stats.ms_hss = b->stats.ms_pack_kernel = 0.0f;
  for (auto& e : b->events) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
      if (e.stage == 4) b->stats.ms_pack_kernel += ms;
      else if (e.stage == 0) b->stats.ms_pack += ms;
      else if (e.stage == 1) b->stats.ms_sigma += ms;
      else if (e.stage == 2) b->stats.ms_dp += ms;
      else b->stats.ms_hss += ms;
    }
  }
  b->ran = true;
  b->downloaded = false;
  return RC_OK;
}

// ------------------------------------------------------------------------------------------------
// download + result access
// ------------------------------------------------------------------------------------------------
extern "C" int rc_batch_download(rc_batch* b) {
  if (!b) return RC_ERR_ARG;
  rc_ctx* ctx = b->ctx;
  if (!b->ran) {
    ctx_fail(ctx, "rc_batch_download before rc_batch_run");
    return RC_ERR_STATE;
  }
  RC_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  RC_CUDA(cudaMemcpyAsync(b->h_res.data(), b->d_res, sizeof(float) * b->res_floats, cudaMemcpyDeviceToHost, st));
  RC_CUDA(cudaMemcpyAsync(b->h_hss.data(), b->d_hss, sizeof(HssDev) * b->hss_count, cudaMemcpyDeviceToHost, st));
  RC_CUDA(cudaMemcpyAsync(b->h_hsscnt.data(), b->d_hsscnt, sizeof(int) * b->hsscnt_ints, cudaMemcpyDeviceToHost, st));
  RC_CUDA(cudaStreamSynchronize(st));
  b->stats.d2h_bytes = sizeof(float) * b->res_floats + sizeof(HssDev) * b->hss_count + sizeof(int) * b->hsscnt_ints;
  b->downloaded = true;
  return RC_OK;
}

extern "C" int rc_batch_native_hss(rc_batch* b, int block, rc_hss* out, int max_hss, int* n_hss) {
  if (!b || block < 0 || block >= b->n_blocks || !n_hss) return RC_ERR_ARG;
  if (!b->downloaded) {
    ctx_fail(b->ctx, "rc_batch_native_hss before rc_batch_download");
    return RC_ERR_STATE;
  }
  const BlockDev& bd = b->blocks[block];
  int n = 0;
  if (bd.L >= 3) {
    for (int s = 0; s < 2; s++)      // '+' first, then '-' (src/score.c:1107-1127)
      for (int f = 0; f < 3; f++) {  // frames in order (src/score.c:880)
        const int cnt = b->h_hsscnt[bd.hsscnt_off + s * 3 + f];
        if (cnt < 0) {
          ctx_fail(b->ctx, "internal: unresolved band overflow");
          return RC_ERR_STATE;
        }
        for (int i = 0; i < cnt; i++) {
          const HssDev& h = b->h_hss[bd.hss_off[s][f] + i];
          if (out && n < max_hss) {
            out[n].strand = s ? '-' : '+';
            out[n].frame = f;
            out[n].startSite = h.startSite;
            out[n].endSite = h.endSite;
            out[n].score = h.score;
          }
          n++;
        }
      }
  }
  *n_hss = n;
  return (out == nullptr || n <= max_hss) ? RC_OK : RC_ERR_CAPACITY;
}

extern "C" int rc_batch_max_scores(rc_batch* b, int block, double* max_scores) {
  if (!b || block < 0 || block >= b->n_blocks || !max_scores) return RC_ERR_ARG;
  if (!b->downloaded) {
    ctx_fail(b->ctx, "rc_batch_max_scores before rc_batch_download");
    return RC_ERR_STATE;
  }
  const BlockDev& bd = b->blocks[block];
  for (int inst = 1; inst < bd.n_inst; inst++) {
    float best = -1.0f;  // results[0].score of an empty list (src/score.c:1129-1134, :1044)
    if (bd.L >= 3)
      for (int sf = 0; sf < 6; sf++)
        if (bd.sites[sf % 3] > 0) best = std::max(best, b->h_res[bd.res_off + (size_t)inst * 6 + sf]);
    max_scores[inst - 1] = (double)best;
  }
  return RC_OK;
}

extern "C" int rc_batch_max_scores_all(rc_batch* b, double* max_scores, size_t n_out) {
  if (!b || !max_scores) return RC_ERR_ARG;
  size_t need = 0;
  for (int i = 0; i < b->n_blocks; i++) need += (size_t)b->blocks[i].n_inst - 1;
  if (need != n_out) {
    ctx_fail(b->ctx, "rc_batch_max_scores_all: n_out must be the sum of the blocks' n_samples");
    return RC_ERR_ARG;
  }
  size_t pos = 0;
  for (int i = 0; i < b->n_blocks; i++) {
    const int r = rc_batch_max_scores(b, i, max_scores + pos);
    if (r != RC_OK) return r;
    pos += (size_t)b->blocks[i].n_inst - 1;
  }
  return RC_OK;
}

extern "C" int rc_batch_get_stats(rc_batch* b, rc_batch_stats* stats) {
  if (!b || !stats) return RC_ERR_ARG;
  b->stats.device_bytes = b->device_bytes + b->dense_floats * sizeof(float);
  *stats = b->stats;
  return RC_OK;
}

// ------------------------------------------------------------------------------------------------
// one block at a time
// ------------------------------------------------------------------------------------------------
extern "C" int rc_score_aln(rc_ctx* ctx, const rc_block_desc* block, const rc_params* params, const int* blosum,
                            rc_hss* out, int max_hss, int* n_hss) {
  if (!ctx || !block || !n_hss) return RC_ERR_ARG;
  rc_block_desc d = *block;
  d.n_samples = 0;
  d.samples = nullptr;
  rc_batch* b = nullptr;
  int r = rc_batch_create(ctx, &d, 1, params, blosum, &b);
  if (r != RC_OK) return r;
  if ((r = rc_batch_upload(b)) == RC_OK && (r = rc_batch_run(b)) == RC_OK && (r = rc_batch_download(b)) == RC_OK)
    r = rc_batch_native_hss(b, 0, out, max_hss, n_hss);
  rc_batch_destroy(b);
  return r;
}

// (f4) rows of the pairwise matrices for backtrack(), see k_pair_rows
#define RC_CUDA_D(call)                                                         \
  do {                                                                          \
    cudaError_t _e = (call);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ctx_fail(ctx, std::string(#call) + ": " + cudaGetErrorString(_e));        \
      cleanup();                                                                \
      return RC_ERR_CUDA;                                                       \
    }                                                                           \
  } while (0)
extern "C" int rc_pair_rows(rc_ctx* ctx, const rc_block_desc* block, const rc_params* params, const int* blosum, int strand,
                            int n_rows, const int* b, float* out) {
  if (!ctx) return RC_ERR_ARG;
  if (!block || !params || !blosum || !b || !out || n_rows < 1 || strand < 0 || strand > 1 || block->N < 2 || block->N > 500 ||
      block->cols < 1 || !block->rows || !block->scores_fwd || !block->scores_rev) {
    ctx_fail(ctx, "rc_pair_rows: invalid argument");
    return RC_ERR_ARG;
  }
  const int N = block->N, cols = block->cols;
  unsigned char lut[256];
  build_lut(lut);
  // class bytes in the strand's own orientation (revAln: reversed and complemented), nucleotide code in bits 0-1
  std::vector<unsigned char> cls((size_t)N * cols);
  for (int k = 0; k < N; k++)
    for (int c = 0; c < cols; c++) {
      const unsigned char v = lut[(unsigned char)block->rows[(size_t)k * cols + (strand ? cols - 1 - c : c)]];
      cls[(size_t)k * cols + c] = strand ? (unsigned char)((v & ~3u) | ((v >> 2) & 3u)) : v;
    }
  std::vector<int> c0(1, -1);
  for (int c = 0; c < cols; c++)
    if (!(cls[c] & CLS_GAP)) c0.push_back(c);
  const int L = (int)c0.size() - 1;
  for (int r = 0; r < n_rows; r++)
    if (b[r] < 1 || b[r] > L) {
      ctx_fail(ctx, "rc_pair_rows: start position outside 1..L");
      return RC_ERR_ARG;
    }
  SigmaTables tab{};
  fill_tables(tab, blosum);
  const Params prm{params->Delta, params->Omega, params->omega, params->stopPenalty_0, params->stopPenalty_k};
  const size_t out_floats = (size_t)n_rows * N * 3 * (L + 1);
  unsigned char* d_cls = nullptr;
  int *d_c0 = nullptr, *d_b = nullptr;
  float *d_sc = nullptr, *d_out = nullptr;
  SigmaTables* d_tab = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_cls); cudaFree(d_c0); cudaFree(d_b); cudaFree(d_sc); cudaFree(d_out); cudaFree(d_tab);
  };
  cudaStream_t st = ctx->stream;
  RC_CUDA_D(cudaSetDevice(ctx->device));
  RC_CUDA_D(cudaMalloc((void**)&d_cls, cls.size()));
  RC_CUDA_D(cudaMalloc((void**)&d_c0, sizeof(int) * c0.size()));
  RC_CUDA_D(cudaMalloc((void**)&d_b, sizeof(int) * n_rows));
  RC_CUDA_D(cudaMalloc((void**)&d_sc, sizeof(float) * 4 * N));
  RC_CUDA_D(cudaMalloc((void**)&d_out, sizeof(float) * out_floats));
  RC_CUDA_D(cudaMalloc((void**)&d_tab, sizeof(SigmaTables)));
  RC_CUDA_D(cudaMemcpyAsync(d_cls, cls.data(), cls.size(), cudaMemcpyHostToDevice, st));
  RC_CUDA_D(cudaMemcpyAsync(d_c0, c0.data(), sizeof(int) * c0.size(), cudaMemcpyHostToDevice, st));
  RC_CUDA_D(cudaMemcpyAsync(d_b, b, sizeof(int) * n_rows, cudaMemcpyHostToDevice, st));
  RC_CUDA_D(cudaMemcpyAsync(d_sc, strand ? block->scores_rev : block->scores_fwd, sizeof(float) * 4 * N, cudaMemcpyHostToDevice, st));
  RC_CUDA_D(cudaMemcpyAsync(d_tab, &tab, sizeof(SigmaTables), cudaMemcpyHostToDevice, st));
  RC_CUDA_D(cudaMemsetAsync(d_out, 0, sizeof(float) * out_floats, st));
  const int threads = n_rows * (N - 1);
  k_pair_rows<<<(threads + 127) / 128, 128, 0, st>>>(d_cls, d_c0, d_sc, d_tab, d_b, n_rows, N, cols, L, prm, d_out);
  RC_CUDA_D(cudaGetLastError());
  RC_CUDA_D(cudaMemcpyAsync(out, d_out, sizeof(float) * out_floats, cudaMemcpyDeviceToHost, st));
  RC_CUDA_D(cudaStreamSynchronize(st));
  cleanup();
  return RC_OK;
}
#undef RC_CUDA_D

extern "C" int rc_score_samples_evolve(rc_ctx* ctx, const rc_block_desc* block, const rc_tree_desc* tree,
                                       const unsigned int* seeds, int rng, const rc_params* params, const int* blosum,
                                       double* max_scores) {
  if (!ctx || !block || !max_scores) return RC_ERR_ARG;
  rc_block_desc d = *block;
  d.samples = nullptr;
  rc_batch* b = nullptr;
  int r = rc_batch_create(ctx, &d, 1, params, blosum, &b);
  if (r != RC_OK) return r;
  if ((r = rc_batch_set_evolve(b, 0, tree, seeds, rng)) == RC_OK && (r = rc_batch_upload(b)) == RC_OK &&
      (r = rc_batch_run(b)) == RC_OK && (r = rc_batch_download(b)) == RC_OK)
    r = rc_batch_max_scores(b, 0, max_scores);
  rc_batch_destroy(b);
  return r;
}

extern "C" int rc_score_samples(rc_ctx* ctx, const rc_block_desc* block, const rc_params* params, const int* blosum,
                                double* max_scores) {
  if (!ctx || !block || !max_scores) return RC_ERR_ARG;
  rc_batch* b = nullptr;
  int r = rc_batch_create(ctx, block, 1, params, blosum, &b);
  if (r != RC_OK) return r;
  if ((r = rc_batch_upload(b)) == RC_OK && (r = rc_batch_run(b)) == RC_OK && (r = rc_batch_download(b)) == RC_OK)
    r = rc_batch_max_scores(b, 0, max_scores);
  rc_batch_destroy(b);
  return r;
}
