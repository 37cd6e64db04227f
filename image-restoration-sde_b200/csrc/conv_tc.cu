// tcgen05 "tap-GEMM" convolution for sm_100a: bf16 NHWC activations, bf16 weights, fp32 accumulate in
// TMEM.
//
//   D[pixel, n] = sum_{tap} sum_{c} A[pixel + (dh,dw)_tap, c] * Wt[phase][tap][n][c]
//
// One CTA computes a 128-pixel x BN-channel tile.  The 128 pixels are a BW x BH patch of one image
// (BW*BH = 128), so that for every tap the A operand is ONE TMA box {64 ch, BW, BH} of the NHWC
// tensor shifted by (dw,dh): zero padding is the TMA out-of-bounds fill, im2col never exists in
// memory.  The box lands in shared memory as 128 rows x 128 B with the 128B swizzle, which is exactly
// the canonical K-major UMMA operand layout; the weights tile ({64 ch, BN} of [tap][Cout][Cin]) likewise.
// This one kernel serves: 3x3 (9 taps), 1x1 (1 tap), 4x4/s2 (16 taps over 4 space-to-depth planes),
// nearest-x2-upsample + 3x3 (4 output phases x 4 taps with pre-summed weights).
//
// One persistent kernel (conv_tc_persist_kernel<BN, MODE, CG, EKT>), 320 threads: warp 0 = TMEM alloc + TMA producer, warp 1 =
// MMA issuer (one elected lane), warps 2..9 = epilogue (tcgen05.ld -> bias / (scale+1)*x+shift / SiLU / q-softmax /
// +residual -> bf16 -> swizzled staging -> TMA bulk store at the channel offset of the destination concat buffer).
// Pipeline: STAGES-deep smem ring with full/empty mbarriers (tcgen05.commit releases slots), double-buffered TMEM
// accumulators so the epilogue of tile i overlaps the main loop of tile i+1.  Variants: MODE 0 per-tap boxes, MODE 2 ROWS
// (3x3, row-window reuse), MODE 3 fp32x3 (3 x kind::tf32 on split fp32 operands), CG = 2 CTA pairs (cta_group::2), EKT =
// epilogue features compiled in.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace irsde {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

struct TcParams {
  int B, H, W;          // pixel grid enumerated by the tiles (== A tensor spatial size)
  int Cin, Cout;
  int ntaps, kchunks;
  int BW, BH, logBW;
  int tilesW, tilesH;
  int nphases;
  int Hout, Wout, os;   // output spatial size and output stride (2 for the upsample phases)
  bf16* out;
  int out_pitch;
  float* out_nchw;      // if set: write fp32 NCHW cropped to (cropH, cropW) instead of the NHWC view
  int cropH, cropW;
  int tma_store;        // persistent kernel: bf16 NHWC output through smem + TMA bulk store
  int qsm;              // LinearAttention q: softmax over each 32-channel head * 32^-0.5 for output channels < 128
  int w_per_image;      // weights tensor is [B][Cout][Cin]: third TMA coordinate = image index
  float* out_f32;       // MODE 3 (fp32x3): fp32 NHWC output view (channel offset applied), pitch = out_pitch
  const float* res_f32; // MODE 3: fp32 residual view
  int lo_plane_off;     // MODE 3: plane offset of the "lo" split of the A tensor ([hi planes][lo planes])
  // tile decode without integer divisions (they cost ~25 instructions + a MUFU.RCP each, five per tile and warp: on layers
  // with one k-iteration per tile the kernel was ISSUE-bound on them, ncu r02): q = umulhi(x, mg) with mg = ceil(2^32 / d)
  unsigned mg_tiles_m, mg_tiles_n, mg_tilesW, mg_tilesH, mg_B, mg_pairs_m;  // 0: divisor is 1 (or fast path not provable): see fdiv
  int epi_kind;         // EK_* combination the epilogue is specialised for (EK_GENERIC: every feature checked at run time)
  int n_fast;           // tile order: N tiles of one M tile are consecutive (1x1 convs with several N tiles, see tc_conv_create)
  int tiles_n;
  int rows_a_bytes;     // ROWS mode: bytes of one A box ((BH+2) x BW x 128)
  unsigned long long* dbg;  // optional [grid][8] cycle counters (IRSDE_TC_DEBUG=1)
  const bf16* res;
  int res_pitch;
  const float* bias;
  const float* mult_vec;
  const float* ss;
  const int* t_ptr;
  int ss_S, ss_off, ss_img_stride, silu;
  TcTap taps[16];
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (fails the launch) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0) {
      long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000LL) {  // ~2 s
        printf("irsde conv_tc: mbarrier wait timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// SM100 shared-memory matrix descriptor: K-major operand, 128B swizzle, rows of 128 bytes,
// 8-row core groups 1024 B apart (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, uint32_t sbo_bytes = 1024) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)(0) << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                             // version = 1 (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N.
__device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// kind::tf32 instruction descriptor: D=f32, A=B=tf32 (32-bit containers, 10-bit mantissa), both K-major, M x N; K = 8.
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
// The registers are tied to the wait ("+r") so the compiler cannot schedule their consumers above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t v[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

// ---- CTA-pair (cta_group::2) variants: per-tap BN = 256 layers, opt-in / auto (g_pair) ----------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Executed by both CTAs of the pair: the data lands in the issuing CTA's shared memory, the transaction bytes are
// counted on the LEADER's barrier (clearing the CTA-rank bit of the shared::cluster address selects the even CTA).
constexpr uint32_t PAIR_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_5d_pair(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar) & PAIR_PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar) & PAIR_PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// one commit arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3)
               : "memory");
}
// arrive on the LEADER's copy of a barrier from either CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}

constexpr int A_STAGE_BYTES = 128 * 128;  // 128 pixel rows x 64 bf16

// =====================================================================================================
// Persistent kernel: one CTA per SM loops over output tiles.  The TMEM accumulator is
// double buffered (2 x BN columns) so the epilogue of tile i overlaps the MMA main loop of tile i+1,
// the smem ring is as deep as 192 KB allows (BN=256: 4, BN=128: 6, BN<=64: 8 stages) to cover TMA
// latency, and the prologue (barrier init, TMEM alloc, descriptor prefetch) is paid once per SM.
// =====================================================================================================
//
// ROWS mode (default for 3x3 stride-1 convs with BN <= 128): a stage is ONE TMA box of (BH+2) image rows x BW pixels
// for a given horizontal shift dw, plus the three weight tiles of the taps (dh=-1,0,1; dw).  The three vertical taps
// read 1024-byte-ALIGNED row windows of the same box (window dh starts (dh+1)*BW pixel rows in), so the A operand
// is fetched 3x instead of 9x per 64-channel chunk with ordinary aligned UMMA descriptors, and one full/empty
// barrier round covers 12 MMAs.
constexpr int TF3_SEG = 2;  // fp32x3: k-iterations (of 32 channels, 12 MMAs each) accumulated in TMEM before a flush to registers
constexpr int ROWS_A_SLOT = 20 * 1024;  // (8+2) rows x 16 px x 128 B (or 18 x 8 px for narrow images)

// MODE 3 "fp32x3" (the fp32-accurate tensor-core mode, SURVEY 7 "Parity vs precision"): operands are fp32 tensors that were
// split into hi = rn_tf32(x) and lo = rn_tf32(x - hi) (|x - hi - lo| <= 2^-22 |x|); a K chunk is 32 fp32 channels = the same
// 128-byte swizzled row as 64 bf16 channels; per chunk the MMA warp issues A_hi*W_hi + A_lo*W_hi + A_hi*W_lo as three
// groups of four tcgen05.mma.kind::tf32 (K=8) into the SAME fp32 TMEM accumulator (the dropped lo*lo term is 2^-22 relative).
// The epilogue keeps everything in fp32 (IEEE SiLU, fp32 residual, fp32 NHWC / NCHW stores).
// CG = 2: CTA pair (cta_group::2) - a cluster of two CTAs computes a 256-pixel x 256-channel tile with M = 256 MMAs issued by
// the leader; each CTA stages its own 128 pixels of A and HALF of the weight tile (32 KB instead of 48 KB per k-iteration
// and SM: the BN = 256 layers were L2->smem feed bound, MMA warp ~40 % of the time on the `full` barrier).
template <int BN, int MODE, int CG = 1>
struct TcCfgP {
  static constexpr bool ROWS = MODE == 2;
  static constexpr bool TF3 = MODE == 3;
  static constexpr int B_STAGE_BYTES = BN * 128 / CG;
  // G consecutive k-iterations (64-channel chunks) share one full/empty barrier round, which amortises the
  // mbarrier wait + tcgen05.commit of the single MMA-issuing thread over 4*G MMAs (matters for narrow N tiles)
  static constexpr int G = 1;  // measured: G=2 trades issue overhead for coarser prefetch granularity; no net gain
  // Epilogue staging for the TMA stores: one 2 KB tile per epilogue warp.  (Two tiles per warp - so that a warp need not
  // wait for its previous bulk store to finish reading the tile - were measured in round 2, same-box ABAB: no change;
  // the store-heavy 1x1 convs are bound by L2 traffic (A re-read per N tile + weight tile re-fetched per output tile),
  // not by the store latency.  STG_BUFS = 2 still works where shared memory allows it.)
  static constexpr int STG_BUFS = TF3 ? 0 : 1;
  static constexpr int STAGES = CG == 2 ? 6 : TF3 ? (BN == 128 ? 3 : (BN == 64 ? 4 : 5))
                                : ROWS ? (BN == 128 ? 3 : (BN == 64 ? 4 : 5))
                                       : (BN == 256 ? 4 : (BN == 128 ? 6 : 8));
  static constexpr int BSUB = ROWS ? 3 : (TF3 ? 2 : 1);  // weight tiles per stage
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;  // per k-iteration (mode 0)
  static constexpr int A_BYTES = TF3 ? STAGES * 2 * A_STAGE_BYTES
                                     : (ROWS ? STAGES * ROWS_A_SLOT : STAGES * G * A_STAGE_BYTES);
  static constexpr int EPI_FLOATS = 3 * BN;
  static constexpr int STG_BYTES = STG_BUFS * 8 * 2048;  // per epilogue warp: STG_BUFS 32 rows x 64 B staging tiles for TMA stores
  static constexpr int SMEM_BYTES = 1024 + A_BYTES + STAGES * G * BSUB * B_STAGE_BYTES + STG_BYTES + EPI_FLOATS * 4 + 512;
  static_assert(SMEM_BYTES <= 232448, "over the 227 KB per-CTA shared memory limit");
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
};

// One elected lane of a converged warp (the compiler recognises elect.sync and emits straight-line uniform
// datapath code for the UTMA / UTCHMMA instructions under it, instead of per-active-thread ELECT loops).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct OutMaps {
  CUtensorMap m[4];  // one per upsample phase (only m[0] otherwise)
};
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct TileCoord {
  int b, phase, py, px, h0, w0, n0;
};
// x / d for 0 <= x with x * d < 2^32 (host-checked in tc_fill_magic; mg == 0 selects the plain division)
__device__ __forceinline__ int fdiv(int x, int d, unsigned mg) { return mg ? (int)__umulhi((unsigned)x, mg) : x / d; }
enum { EK_QSM = 1, EK_AFF = 2, EK_SILU = 4, EK_RES = 8, EK_GENERIC = 16 };

__device__ __forceinline__ void decode_m(const TcParams& P, int m_idx, TileCoord& t) {
  int q = fdiv(m_idx, P.tilesW, P.mg_tilesW);
  const int tw_i = m_idx - q * P.tilesW; m_idx = q;
  q = fdiv(m_idx, P.tilesH, P.mg_tilesH);
  const int th_i = m_idx - q * P.tilesH; m_idx = q;
  q = fdiv(m_idx, P.B, P.mg_B);
  t.b = m_idx - q * P.B;
  t.phase = q;
  t.py = t.phase >> 1; t.px = t.phase & 1;
  t.h0 = th_i * P.BH; t.w0 = tw_i * P.BW;
}
__device__ __forceinline__ TileCoord decode_tile(const TcParams& P, int tile, int tiles_m, int BN) {
  TileCoord t;
  int m_idx;
  if (P.n_fast) { m_idx = fdiv(tile, P.tiles_n, P.mg_tiles_n); t.n0 = (tile - m_idx * P.tiles_n) * BN; }
  else { const int q = fdiv(tile, tiles_m, P.mg_tiles_m); m_idx = tile - q * tiles_m; t.n0 = q * BN; }
  decode_m(P, m_idx, t);
  return t;
}

// CTA pair: pair tile `pt` covers the m-tiles 2*mp and 2*mp+1 (one per CTA rank) of one N tile.  An odd tail leaves the
// second CTA without pixels: it still runs the whole protocol on an out-of-range image index (TMA loads zero-fill,
// TMA stores are clipped).
__device__ __forceinline__ TileCoord decode_pair_tile(const TcParams& P, int pt, int rank, int tiles_m, int BN) {
  const int pairs_m = (tiles_m + 1) >> 1;
  const int q = fdiv(pt, pairs_m, P.mg_pairs_m);
  const int mp = pt - q * pairs_m;
  const int m_idx = 2 * mp + rank;
  TileCoord t;
  t.n0 = q * BN;
  if (m_idx >= tiles_m) {
    t.b = P.B; t.phase = 0; t.py = 0; t.px = 0; t.h0 = 0; t.w0 = 0;
    return t;
  }
  decode_m(P, m_idx, t);
  return t;
}

#define DBG_WAIT(ctr, bar, par)                 \
  do {                                          \
    if (P.dbg) {                                \
      long long _t0 = clock64();                \
      mbar_wait(bar, par);                      \
      ctr += clock64() - _t0;                   \
    } else {                                    \
      mbar_wait(bar, par);                      \
    }                                           \
  } while (0)

// EKT: EK_* features the epilogue is compiled for.  Only the per-tap BN <= 128 instantiations that serve the 1x1 convs with one
// or two k-iterations per tile (res_conv: 0, to_qkv: EK_QSM, to_out: EK_AFF) are specialised - there the kernel is issue-bound
// on epilogue instructions; everything else runs EK_GENERIC (an in-kernel switch over all kinds made ptxas spill 3-4 KB).
template <int BN, int MODE, int CG = 1, int EKT = 16 /* EK_GENERIC */>
__global__ void __launch_bounds__(320, 1) conv_tc_persist_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_b,
                                                                 const __grid_constant__ OutMaps map_o,
                                                                 const __grid_constant__ TcParams P, int tiles_m,
                                                                 int num_tiles) {
  using Cfg = TcCfgP<BN, MODE, CG>;
  constexpr bool ROWS = Cfg::ROWS;
  constexpr bool TF3 = Cfg::TF3;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(CG == 1 || (CG == 2 && MODE == 0 && BN == 256), "the CTA-pair variant exists for per-tap mode, BN = 256");
  pdl_trigger();  // the next kernel may be scheduled; this one's prologue below touches no global memory
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                       // A ring
  uint8_t* smem_b = smem + Cfg::A_BYTES;
  uint8_t* smem_stg = smem_b + STAGES * Cfg::G * Cfg::BSUB * Cfg::B_STAGE_BYTES;  // 1024-aligned
  float* s_epi = reinterpret_cast<float*>(smem_stg + Cfg::STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_epi + Cfg::EPI_FLOATS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;       // [2] accumulator ready
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KI = P.ntaps * P.kchunks;
  // CTA pair (CG == 2): blockIdx.x = 2 * cluster + rank; `num_tiles` counts pair tiles; both CTAs walk the same pair tiles
  const int crank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int tile_first = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_coord = [&](int tile) -> TileCoord {
    if constexpr (CG == 2) return decode_pair_tile(P, tile, crank, tiles_m, BN);
    else return decode_tile(P, tile, tiles_m, BN);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8 * CG);  // one arrive per epilogue warp (of both CTAs of a pair, on the leader)
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_b);
    }
    __syncwarp();
    if constexpr (CG == 2) {  // collective over the pair: the same warp of both CTAs, same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(Cfg::TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, TMEM and descriptors are ready: from here on global memory written by the previous kernel is read
  // (activations, the device step counter behind P.t_ptr) and buffers it may still be reading are overwritten
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer (whole warp loops; one elected lane issues) =================
    {
      int it = 0;
      long long w_prod = 0, t_start = clock64();
      for (int tile = tile_first; tile < num_tiles; tile += tile_step) {
        const TileCoord t = tile_coord(tile);
        if constexpr (TF3) {
          for (int ki = 0; ki < KI; ++ki, ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) DBG_WAIT(w_prod, &empty_bar[s], ((it / STAGES) - 1) & 1);
            if (elect_one()) {
              const int tap = ki / P.kchunks, kc = ki - tap * P.kchunks;
              const int dh = P.taps[tap].dh + (P.nphases == 4 ? t.py : 0);
              const int dw = P.taps[tap].dw + (P.nphases == 4 ? t.px : 0);
              const int wt = (t.phase * P.ntaps + tap) * 2;
              mbar_expect_tx(&full_bar[s], 2 * A_STAGE_BYTES + 2 * Cfg::B_STAGE_BYTES);
              tma_load_5d(smem_a + (s * 2 + 0) * A_STAGE_BYTES, &map_a, &full_bar[s], kc * 32, t.w0 + dw, t.h0 + dh, t.b, P.taps[tap].plane);
              tma_load_5d(smem_a + (s * 2 + 1) * A_STAGE_BYTES, &map_a, &full_bar[s], kc * 32, t.w0 + dw, t.h0 + dh, t.b,
                          P.taps[tap].plane + P.lo_plane_off);
              tma_load_3d(smem_b + (s * 2 + 0) * Cfg::B_STAGE_BYTES, &map_b, &full_bar[s], kc * 32, t.n0, wt);
              tma_load_3d(smem_b + (s * 2 + 1) * Cfg::B_STAGE_BYTES, &map_b, &full_bar[s], kc * 32, t.n0, wt + 1);
            }
            __syncwarp();
          }
        } else if constexpr (ROWS) {
          for (int kc = 0; kc < P.kchunks; ++kc) {
            for (int dwi = 0; dwi < 3; ++dwi, ++it) {
              const int s = it % STAGES;
              if (it >= STAGES) DBG_WAIT(w_prod, &empty_bar[s], ((it / STAGES) - 1) & 1);
              if (elect_one()) {
                mbar_expect_tx(&full_bar[s], P.rows_a_bytes + 3 * Cfg::B_STAGE_BYTES);
                tma_load_5d(smem_a + s * ROWS_A_SLOT, &map_a, &full_bar[s], kc * 64, t.w0 + dwi - 1, t.h0 - 1, t.b, 0);
#pragma unroll
                for (int j = 0; j < 3; ++j)  // taps (dh = j-1, dw = dwi-1): index (dh+1)*3 + (dw+1)
                  tma_load_3d(smem_b + (s * 3 + j) * Cfg::B_STAGE_BYTES, &map_b, &full_bar[s], kc * 64, t.n0, j * 3 + dwi);
              }
              __syncwarp();
            }
          }
        } else {
          constexpr int G = Cfg::G;
          for (int g0 = 0; g0 < KI; g0 += G, ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) DBG_WAIT(w_prod, &empty_bar[s], ((it / STAGES) - 1) & 1);
            if (elect_one()) {
              const int n = (KI - g0) < G ? (KI - g0) : G;
              if constexpr (CG == 2) {
                // the leader's barrier collects both CTAs' bytes: own 128 pixels of A + own half (128 channels) of B, each
                if (crank == 0) mbar_expect_tx(&full_bar[s], 2 * n * Cfg::STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < G; ++j) {
                  if (j < n) {
                    const int ki = g0 + j, tap = ki / P.kchunks, kc = ki - tap * P.kchunks;
                    const int dh = P.taps[tap].dh + (P.nphases == 4 ? t.py : 0);
                    const int dw = P.taps[tap].dw + (P.nphases == 4 ? t.px : 0);
                    tma_load_5d_pair(smem_a + (s * G + j) * A_STAGE_BYTES, &map_a, &full_bar[s], kc * 64, t.w0 + dw, t.h0 + dh,
                                     t.b, P.taps[tap].plane);
                    tma_load_3d_pair(smem_b + (s * G + j) * Cfg::B_STAGE_BYTES, &map_b, &full_bar[s], kc * 64,
                                     t.n0 + crank * (BN / 2), t.phase * P.ntaps + tap);
                  }
                }
              } else {
                mbar_expect_tx(&full_bar[s], n * Cfg::STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < G; ++j) {
                  if (j < n) {
                    const int ki = g0 + j, tap = ki / P.kchunks, kc = ki - tap * P.kchunks;
                    const int dh = P.taps[tap].dh + (P.nphases == 4 ? t.py : 0);
                    const int dw = P.taps[tap].dw + (P.nphases == 4 ? t.px : 0);
                    tma_load_5d(smem_a + (s * G + j) * A_STAGE_BYTES, &map_a, &full_bar[s], kc * 64, t.w0 + dw, t.h0 + dh, t.b,
                                P.taps[tap].plane);
                    tma_load_3d(smem_b + (s * G + j) * Cfg::B_STAGE_BYTES, &map_b, &full_bar[s], kc * 64, t.n0,
                                t.phase * P.ntaps + tap + (P.w_per_image ? t.b : 0));
                  }
                }
              }
            }
            __syncwarp();
          }
        }
      }
      if (P.dbg && lane == 0) { P.dbg[blockIdx.x * 8 + 0] = (unsigned long long)w_prod; P.dbg[blockIdx.x * 8 + 1] = (unsigned long long)(clock64() - t_start); }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp loops; one elected lane issues) =================
    {
      const uint32_t idesc = TF3 ? make_idesc_tf32(128, BN) : make_idesc_bf16(128 * CG, BN);  // pair: M = 256 over both CTAs' lanes
      int it = 0, lt = 0;
      long long w_full = 0, w_tempty = 0, t_start = clock64();
      if constexpr (TF3) {
        // fp32x3: the K loop of a tile is cut into SEGMENTS of TF3_SEG k-iterations; each segment accumulates from zero in
        // one of the two TMEM buffers and is handed to the epilogue warps, which sum the segments in fp32 registers with
        // round-to-nearest adds.  Reason (measured, profiles/ + tests/test_gpu_fp32x3.py): the tensor core's accumulator
        // update truncates, so a single TMEM accumulation over n MMAs drifts by ~n * 2^-26 relative - 3e-5 per layer at
        // K = 13 824 (5 184 MMAs), 70x the error of the fp32 FMA path.  24 MMAs per segment (TF3_SEG = 2) keep it at the fp32 FMA path's own level;
        // TF3_SEG = 8 measured 9x that level at nf=64 (2.3e-5 vs 2.6e-6 relative RMS at the last layer).
        int st = 0;  // segments issued by this CTA (TMEM buffer = st & 1)
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          for (int k0 = 0; k0 < KI; k0 += TF3_SEG, ++st) {
            const int acc = st & 1;
            if (st >= 2) DBG_WAIT(w_tempty, &tempty_bar[acc], ((st >> 1) - 1) & 1);  // epilogue drained this buffer
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
            const int k1 = k0 + TF3_SEG < KI ? k0 + TF3_SEG : KI;
            for (int ki = k0; ki < k1; ++ki, ++it) {
              const int s = it % STAGES;
              DBG_WAIT(w_full, &full_bar[s], (it / STAGES) & 1);
              tc_fence_after();
              if (elect_one()) {
                const uint64_t a_hi = make_sw128_desc(smem_u32(smem_a + (s * 2 + 0) * A_STAGE_BYTES));
                const uint64_t a_lo = make_sw128_desc(smem_u32(smem_a + (s * 2 + 1) * A_STAGE_BYTES));
                const uint64_t b_hi = make_sw128_desc(smem_u32(smem_b + (s * 2 + 0) * Cfg::B_STAGE_BYTES));
                const uint64_t b_lo = make_sw128_desc(smem_u32(smem_b + (s * 2 + 1) * Cfg::B_STAGE_BYTES));
                // the two small cross terms first, the dominant hi*hi product last
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 4 x (K=8 tf32) per 32-channel chunk: +32 B inside the swizzle atom
                  umma_tf32(tmem_d, a_lo + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc, (ki > k0 || k > 0) ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_d, a_hi + (uint64_t)(k * 2), b_lo + (uint64_t)(k * 2), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_d, a_hi + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc, 1u);
                umma_commit(&empty_bar[s]);
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit(&tfull_bar[acc]);
            __syncwarp();
          }
        }
      }
      for (int tile = tile_first; tile < num_tiles && !TF3 && (CG == 1 || crank == 0); tile += tile_step, ++lt) {  // pair: leader only
        const int acc = lt & 1;
        if (lt >= 2) DBG_WAIT(w_tempty, &tempty_bar[acc], ((lt >> 1) - 1) & 1);  // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        if constexpr (ROWS) {
          const int nst = 3 * P.kchunks;
          for (int si = 0; si < nst; ++si, ++it) {
            const int s = it % STAGES;
            DBG_WAIT(w_full, &full_bar[s], (it / STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a0 = smem_u32(smem_a + s * ROWS_A_SLOT);
#pragma unroll
              for (int j = 0; j < 3; ++j) {  // vertical tap dh = j-1: row window starting j*BW pixel rows into the box
                const uint64_t adesc = make_sw128_desc(a0 + (uint32_t)(j * P.BW * 128));
                const uint64_t bdesc = make_sw128_desc(smem_u32(smem_b + (s * 3 + j) * Cfg::B_STAGE_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (si > 0 || j > 0 || k > 0) ? 1u : 0u);
              }
              umma_commit(&empty_bar[s]);
            }
            __syncwarp();
          }
        } else {
          constexpr int G = Cfg::G;
          for (int g0 = 0; g0 < KI; g0 += G, ++it) {
            const int s = it % STAGES;
            DBG_WAIT(w_full, &full_bar[s], (it / STAGES) & 1);
            tc_fence_after();
            if (elect_one()) {
              const int n = (KI - g0) < G ? (KI - g0) : G;
#pragma unroll
              for (int j = 0; j < G; ++j) {
                if (j < n) {
                  const uint64_t adesc = make_sw128_desc(smem_u32(smem_a + (s * G + j) * A_STAGE_BYTES));
                  const uint64_t bdesc = make_sw128_desc(smem_u32(smem_b + (s * G + j) * Cfg::B_STAGE_BYTES));
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if constexpr (CG == 2)
                      umma_bf16_pair(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (g0 + j > 0 || k > 0) ? 1u : 0u);
                    else
                      umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (g0 + j > 0 || k > 0) ? 1u : 0u);
                  }
                }
              }
              if constexpr (CG == 2) umma_commit_pair(&empty_bar[s]);  // frees stage s in both CTAs
              else umma_commit(&empty_bar[s]);
            }
            __syncwarp();
          }
        }
        if (elect_one()) {
          if constexpr (CG == 2) umma_commit_pair(&tfull_bar[acc]);  // both epilogues
          else umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
      if (P.dbg && lane == 0) {
        P.dbg[blockIdx.x * 8 + 2] = (unsigned long long)w_full;
        P.dbg[blockIdx.x * 8 + 3] = (unsigned long long)w_tempty;
        P.dbg[blockIdx.x * 8 + 4] = (unsigned long long)(clock64() - t_start);
      }
    }
  } else {
    // ================= epilogue (warps 2..9) =================
    // Two warps per TMEM lane quadrant (warp & 3), taking alternate 32-column chunks; the tcgen05.ld of the next
    // chunk is in flight while the current one is converted, staged (64B-swizzled) and bulk-stored by TMA.
    const int et = threadIdx.x - 64;  // 0..255
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = quad * 32 + lane;
    const int trow = P.t_ptr ? *P.t_ptr : 0;
    const bool affine = (P.ss != nullptr) || (P.bias != nullptr) || (P.mult_vec != nullptr);
    uint8_t* const stg_base = smem_stg + (warp - 2) * 2048 * (Cfg::STG_BUFS > 0 ? Cfg::STG_BUFS : 1);
    int stg_i = 0;  // staging tile used by the next bulk store (alternates when STG_BUFS == 2)
    long long w_tfull = 0, t_start = clock64();
    int lt = 0, epi_n0 = -1, epi_b = -1;
    int seg_st = 0;  // fp32x3: K segments consumed by this warp (TMEM buffer = seg_st & 1)
    for (int tile = tile_first; tile < num_tiles; tile += tile_step, ++lt) {
      const TileCoord t = tile_coord(tile);
      const int acc = lt & 1;
      // The per-channel (mult, add) table depends on the N tile (and on the image only with per-image timesteps):
      // rebuild it - two barriers and a dependent global load on the epilogue's critical path - only when that changes.
      const int key_b = P.ss_img_stride ? t.b : 0;
      if (affine && (t.n0 != epi_n0 || key_b != epi_b)) {
        epi_n0 = t.n0; epi_b = key_b;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's reads of s_epi are done
        const int tb = (CG == 2 && t.b >= P.B) ? 0 : t.b;  // pair tail: a CTA without pixels still needs a valid table row
        const float* ssrow = P.ss ? P.ss + (long long)(trow + tb * P.ss_img_stride) * P.ss_S + P.ss_off : nullptr;
        for (int j = et; j < BN; j += 256) {
          const int n = t.n0 + j;
          float mult = 1.f, add = 0.f, bias = 0.f;
          if (n < P.Cout) {
            if (ssrow) { mult = ssrow[n] + 1.0f; add = ssrow[P.Cout + n]; }
            if (P.mult_vec) mult *= P.mult_vec[n];
            if (P.bias) bias = P.bias[n];
          }
          s_epi[j] = mult; s_epi[BN + j] = bias * mult + add;  // (x + bias) * mult + add == x * mult + (bias * mult + add)
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const int hh = t.h0 + (m >> P.logBW), ww = t.w0 + (m & (P.BW - 1));
      const bool pvalid = hh < P.H && ww < P.W && (CG == 1 || t.b < P.B);
      const long long opix = ((long long)t.b * P.Hout + (hh * P.os + t.py)) * P.Wout + (ww * P.os + t.px);
      bf16* orow = P.out + opix * P.out_pitch + t.n0;
      const bf16* rrow = P.res ? P.res + opix * P.res_pitch + t.n0 : nullptr;

      // KC: std::integral_constant<int, KIND>.  KIND = the EK_* features this conv has, fixed per launch: feature blocks
      // outside KIND do not exist in the instantiation (the compiler had turned several of them into predicated code that
      // was issued for every chunk), EK_GENERIC keeps every run-time check (fp32 NCHW output, no-TMA-store fallback).
      auto process = [&](auto KC, uint32_t* v, const int c0) __attribute__((always_inline)) {
        constexpr int KIND = decltype(KC)::value;
        constexpr bool GEN = KIND == EK_GENERIC;
        if constexpr (TF3) {  // fp32 in, fp32 out: IEEE arithmetic, no fast-math approximations
          if (affine) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), s_epi[c0 + j], s_epi[BN + c0 + j]));
          }
          if (P.silu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = __uint_as_float(v[j]);
              v[j] = __float_as_uint(x / (1.0f + expf(-x)));
            }
          }
          if (P.out_nchw) {
            if (pvalid && hh < P.cropH && ww < P.cropW) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int n = t.n0 + c0 + j;
                if (n < P.Cout) P.out_nchw[(((long long)t.b * P.Cout + n) * P.cropH + hh) * P.cropW + ww] = __uint_as_float(v[j]);
              }
            }
            return;
          }
          if (!pvalid) return;
          float* of = P.out_f32 + opix * P.out_pitch + t.n0 + c0;
          const float* rf = P.res_f32 ? P.res_f32 + opix * P.res_pitch + t.n0 + c0 : nullptr;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (t.n0 + c0 + g * 4 < P.Cout) {
              float4 o = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                                     __uint_as_float(v[g * 4 + 3]));
              if (rf) {
                const float4 r = *reinterpret_cast<const float4*>(rf + g * 4);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
              }
              *reinterpret_cast<float4*>(of + g * 4) = o;
            }
          }
          return;
        }
        if constexpr (GEN || (KIND & EK_QSM)) if ((!GEN || P.qsm) && t.n0 + c0 < 128) {  // one chunk == one attention head of this pixel's q (module_util.py:168,171)
          float mx = __uint_as_float(v[0]);
#pragma unroll
          for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          float sm = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float e = __expf(__uint_as_float(v[j]) - mx);
            v[j] = __float_as_uint(e);
            sm += e;
          }
          const float inv = __fdividef(0.17677669529663687f, sm);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * inv);
        }
        if constexpr (GEN || (KIND & EK_AFF)) if (!GEN || affine) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(v[j]), s_epi[c0 + j], s_epi[BN + c0 + j]));
        }
        if constexpr (GEN || (KIND & EK_SILU)) if (!GEN || P.silu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            // SiLU(x) = x * (0.5 + 0.5 tanh(x/2)): one MUFU op (tanh.approx, rel. error 2^-11 << bf16 rounding) instead of
            // ex2 + rcp; on the narrow layers the epilogue is MUFU-bound (128 x BN outputs per tile, 16 MUFU/clk/SM)
            const float x = __uint_as_float(v[j]);
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * x));
            v[j] = __float_as_uint(x * fmaf(th, 0.5f, 0.5f));
          }
        }
        if constexpr (GEN) if (P.out_nchw) {
          if (pvalid && hh < P.cropH && ww < P.cropW) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = t.n0 + c0 + j;
              if (n < P.Cout) P.out_nchw[(((long long)t.b * P.Cout + n) * P.cropH + hh) * P.cropW + ww] = __uint_as_float(v[j]);
            }
          }
          return;
        }
        uint4 ov[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cg = c0 + g * 8;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
          if constexpr (GEN || (KIND & EK_RES)) if (rrow && pvalid && t.n0 + cg < P.Cout) {
            uint4 r = *reinterpret_cast<const uint4*>(rrow + cg);
            const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              f[2 * j] += __low2float(r2[j]);
              f[2 * j + 1] += __high2float(r2[j]);
            }
          }
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&ov[g]);
#pragma unroll
          for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        }
        if (!GEN || P.tma_store) {
          // bf16 NHWC through shared memory + TMA store: each lane owns one pixel row of 32 channels (64 B); the
          // staging tile uses the 64B swizzle so the 16-byte st.shared are conflict free and the bulk store writes
          // full, coalesced lines (out-of-range pixels / channels are clipped by TMA).
          uint8_t* stg = stg_base + stg_i * 2048;
          if (lane == 0) {  // the bulk store that last used THIS tile has finished reading it
            if constexpr (Cfg::STG_BUFS == 2) bulk_wait_read1(); else bulk_wait_read0();
          }
          if constexpr (Cfg::STG_BUFS == 2) stg_i ^= 1;
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = ov[g];
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int m0 = quad * 32;
            tma_store_4d(&map_o.m[t.phase], stg, t.n0 + c0, t.w0 + (m0 & (P.BW - 1)), t.h0 + (m0 >> P.logBW), t.b);
            bulk_commit();
          }
        } else if (pvalid) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (t.n0 + c0 + g * 8 < P.Cout) *reinterpret_cast<uint4*>(orow + c0 + g * 8) = ov[g];
        }
      };
      typedef std::integral_constant<int, EK_GENERIC> KGen;

      constexpr int NCH = (BN / 32 + 1) / 2;  // chunks per warp
      if constexpr (TF3) {
        // sum the tile's K segments in registers (IEEE round-to-nearest adds), then run the epilogue on the sums
        float accf[NCH][32];
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
          for (int j = 0; j < 32; ++j) accf[i][j] = 0.f;
        for (int k0 = 0; k0 < KI; k0 += TF3_SEG, ++seg_st) {
          const int sacc = seg_st & 1;
          DBG_WAIT(w_tfull, &tfull_bar[sacc], (seg_st >> 1) & 1);
          tc_fence_after();
          const uint32_t tmem_seg = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sacc * BN);
#pragma unroll
          for (int i = 0; i < NCH; ++i) {
            const int c0 = (2 * i + half) * 32;
            if (c0 < BN) {
              uint32_t v1[32];
              tmem_ld32(tmem_seg + (uint32_t)c0, v1);
#pragma unroll
              for (int j = 0; j < 32; ++j) accf[i][j] = __fadd_rn(accf[i][j], __uint_as_float(v1[j]));
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[sacc]);
        }
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c0 = (2 * i + half) * 32;
          if (c0 < BN && t.n0 + c0 < P.Cout) process(KGen{}, reinterpret_cast<uint32_t*>(accf[i]), c0);
        }
        continue;
      }
      DBG_WAIT(w_tfull, &tfull_bar[acc], (lt >> 1) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
      auto run_chunks = [&](auto KC) __attribute__((always_inline)) {
        uint32_t v[2][32];
        if (half * 32 < BN) tmem_ld32_async(tmem_acc + (uint32_t)(half * 32), v[0]);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          const int c0 = (2 * i + half) * 32;
          if (c0 < BN) {
            tmem_ld_wait(v[i & 1]);
            const int c1 = (2 * (i + 1) + half) * 32;
            if (i + 1 < NCH && c1 < BN) tmem_ld32_async(tmem_acc + (uint32_t)c1, v[(i + 1) & 1]);
            if (t.n0 + c0 < P.Cout) process(KC, v[i & 1], c0);  // chunks entirely past Cout carry no output
          }
        }
      };
      run_chunks(std::integral_constant<int, EKT>{});  // EKT: the kernel instantiation's epilogue specialisation (tc_conv_launch)
      // every tcgen05.ld of this warp has completed: hand the accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_leader(&tempty_bar[acc]);  // the leader's MMA warp owns both accumulators' reuse
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
    if (lane == 0) bulk_wait_all();  // staging smem must stay valid until the last bulk store has read it
    if (P.dbg && threadIdx.x == 64) {
      P.dbg[blockIdx.x * 8 + 5] = (unsigned long long)w_tfull;
      P.dbg[blockIdx.x * 8 + 6] = (unsigned long long)(clock64() - t_start);
      P.dbg[blockIdx.x * 8 + 7] = (unsigned long long)lt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) {
    cluster_sync_all();  // the leader's MMAs read the peer's shared memory and write its TMEM: nobody leaves early
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    return;
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
  }
}

int g_num_sms = 148;
bool g_tma_store = true;
bool g_rows = true;
int g_epi_spec = 3;  // specialised epilogues: 1 = the per-tap 1x1 layers, 2 = + the ROWS 3x3 ResBlock convs, 3 = + the BN = 256
                     // CTA-pair ResBlock convs and biased per-tap convs; IRSDE_TC_EPI_SPEC=0: generic only
bool g_pair = true;    // cta_group::2 tiles (CTA pairs, M = 256) for the per-tap BN = 256 3x3 layers; IRSDE_TC_PAIR=0 disables.
                       // Same-box ABAB (round 2, config 2): 720.0 / 719.5 -> 710.4 / 710.5 ms per chain; the Cout >= 256 3x3
                       // layers go from 1.18-1.30 to 1.23-1.36 PFLOP/s (84-93 % of the sustained bf16 peak)

}  // namespace

struct TcConvDesc {
  CUtensorMap map_a, map_b;
  OutMaps map_o;
  TcParams P;
  int BN;
  dim3 grid;
  int tiles_m, num_tiles;
  int mode;  // 0 per-tap boxes, 2 ROWS (row-window reuse), 3 fp32x3
  bool pair = false;   // CTA-pair (cta_group::2) variant
  int pair_tiles = 0;  // ceil(tiles_m / 2) * N tiles
};

bool tc_init(std::string* err) {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    if (err) *err = "cuTensorMapEncodeTiled not available from the driver";
    return false;
  }
  g_encode = (EncodeTiledFn)fn;
  cudaFuncSetAttribute(conv_tc_persist_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<32, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<32, 2>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<64, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<64, 2>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 2>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<64, 2, 1, EK_AFF | EK_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<64, 2>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<64, 2, 1, EK_SILU | EK_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<64, 2>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 2, 1, EK_AFF | EK_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 2>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 2, 1, EK_SILU | EK_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 2>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<64, 0, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<64, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<64, 0, 1, EK_AFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<64, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 0, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 0, 1, EK_AFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 0, 1, EK_QSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 0>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<32, 3>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<64, 3>::SMEM_BYTES);
  cudaFuncSetAttribute(conv_tc_persist_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<128, 3>::SMEM_BYTES);
  {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) g_num_sms = n;
    const char* e = getenv("IRSDE_TC_TMA_STORE");
    g_tma_store = !(e && e[0] == '0');
    e = getenv("IRSDE_TC_ROWS");
    g_rows = !(e && e[0] == '0');
    e = getenv("IRSDE_TC_PAIR");
    g_pair = !(e && e[0] == '0');
    e = getenv("IRSDE_TC_EPI_SPEC");
    if (e && e[0] >= '0' && e[0] <= '3') g_epi_spec = e[0] - '0';
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0, 2>::SMEM_BYTES);
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 2, EK_AFF | EK_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0, 2>::SMEM_BYTES);
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 2, EK_SILU | EK_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0, 2>::SMEM_BYTES);
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 1, EK_AFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0>::SMEM_BYTES);
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0>::SMEM_BYTES);
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 2, EK_AFF>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0, 2>::SMEM_BYTES);
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfgP<256, 0, 2>::SMEM_BYTES);
    cudaFuncSetAttribute(conv_tc_persist_kernel<256, 0, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0);
    e = getenv("IRSDE_PDL");
    g_pdl = (e && e[0] == '1');  // opt-in: measured 0.7 % (UNet step) to 3 % (NAFNet step) slower than plain graph edges
  }
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    if (err) *err = std::string("cudaFuncSetAttribute failed: ") + cudaGetErrorString(le);
    g_encode = nullptr;
    return false;
  }
  return true;
}

// tile-decode multipliers and the epilogue specialisation of a finished descriptor (called at the end of both creators)
static void tc_fill_magic(TcConvDesc* d) {
  TcParams& P = d->P;
  const unsigned long long max_x = (unsigned long long)d->num_tiles + 2;
  auto mg = [&](int dv) -> unsigned {
    if (dv <= 1) return 0u;                                       // x / 1: fdiv falls back to the (free) plain division
    if (max_x * (unsigned long long)dv >= (1ull << 32)) return 0u;  // exactness not provable: plain division
    return (unsigned)(((1ull << 32) + (unsigned)dv - 1) / (unsigned)dv);
  };
  P.mg_tiles_m = mg(d->tiles_m);
  P.mg_tiles_n = mg(P.tiles_n);
  P.mg_tilesW = mg(P.tilesW);
  P.mg_tilesH = mg(P.tilesH);
  P.mg_B = mg(P.B);
  P.mg_pairs_m = mg((d->tiles_m + 1) >> 1);
  int k = 0;
  if (P.qsm) k |= EK_QSM;
  if (P.ss || P.bias || P.mult_vec) k |= EK_AFF;
  if (P.silu) k |= EK_SILU;
  if (P.res) k |= EK_RES;
  const bool known = k == 0 || k == EK_QSM || k == EK_AFF || k == EK_SILU || k == (EK_AFF | EK_SILU) || k == (EK_AFF | EK_RES) ||
                     k == (EK_SILU | EK_RES);
  P.epi_kind = (d->mode != 3 && P.tma_store && P.out && known) ? k : EK_GENERIC;
}

TcConvDesc* tc_conv_create(const bf16* in, int in_pitch, int B, int Hin, int Win, int Cin, int planes,
                           const bf16* wpacked, int Cout, int ntaps, const TcTap* taps, int nphases, const Epilogue& ep,
                           bf16* out, int out_pitch, int Hout, int Wout, std::string* err, int flags) {
  auto bad = [&](const char* m) -> TcConvDesc* {
    if (err) *err = m;
    return nullptr;
  };
  if (!g_encode) return bad("tc_init not called");
  if (ntaps < 1 || ntaps > 16) return bad("ntaps out of range");
  const bool nchw_out = (out == nullptr);  // caller sets the fp32 NCHW pointer at launch (tc_conv_set_out_nchw)
  const int Cout_w = (Cout + 7) / 8 * 8;   // rows of the packed weight tensor (zero padded)
  if (Cin % 8 || in_pitch % 8) return bad("input channel count / pitch must be multiples of 8");
  if (!nchw_out && (Cout % 8 || out_pitch % 8)) return bad("output channel count / pitch must be multiples of 8");
  if (((uintptr_t)in & 15) || ((uintptr_t)wpacked & 15) || ((uintptr_t)out & 15)) return bad("operands must be 16-byte aligned");
  if (ep.res && (((uintptr_t)ep.res & 15) || ep.res_pitch % 8)) return bad("residual must be 16-byte aligned");
  TcConvDesc* d = new TcConvDesc();
  TcParams& P = d->P;
  memset(&P, 0, sizeof P);
  P.B = B; P.H = Hin; P.W = Win; P.Cin = Cin; P.Cout = Cout;
  P.ntaps = ntaps; P.kchunks = (Cin + 63) / 64;
  P.nphases = nphases;
  P.Hout = Hout; P.Wout = Wout; P.os = nphases == 4 ? 2 : 1;
  P.out = out; P.out_pitch = out_pitch;
  P.res = (const bf16*)ep.res; P.res_pitch = ep.res_pitch;
  P.bias = ep.bias; P.mult_vec = ep.mult_vec; P.ss = ep.ss; P.t_ptr = ep.t_ptr; P.ss_S = ep.ss_S; P.ss_off = ep.ss_off;
  P.ss_img_stride = ep.ss_img_stride; P.silu = ep.silu;
  for (int i = 0; i < ntaps; ++i) P.taps[i] = taps[i];
  P.qsm = (flags & TC_FLAG_QSOFTMAX) ? 1 : 0;
  P.w_per_image = (flags & TC_FLAG_W_PER_IMAGE) ? 1 : 0;
  if (P.w_per_image && (ntaps != 1 || nphases != 1)) return bad("per-image weights need a 1x1 conv");
  if ((P.qsm || P.w_per_image) && !g_tma_store) return bad("fused attention epilogues need the TMA-store epilogue");
  int BN = Cout >= 256 ? 256 : (Cout > 64 ? 128 : (Cout > 32 ? 64 : 32));
  if (BN == 256 && Cout % 256 != 0 && Cout % 128 == 0) BN = 128;  // e.g. to_qkv (384): no half-empty N tile
  bool rows = false;
  if (g_rows && nphases == 1 && planes == 1 && ntaps == 9 && BN <= 128) {
    rows = true;
    for (int i = 0; rows && i < 9; ++i) rows = taps[i].dh == i / 3 - 1 && taps[i].dw == i % 3 - 1 && taps[i].plane == 0;
  }
  d->mode = rows ? 2 : 0;
  // tile shape: BW x BH = 128 pixels, minimise padded work
  long long best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    int bh = 128 / bw;
    long long cost = (long long)((Win + bw - 1) / bw) * bw * ((Hin + bh - 1) / bh) * bh;
    if (best < 0 || cost < best) { best = cost; P.BW = bw; P.BH = bh; }
  }
  if (rows) {  // (BH+2) x BW box must fit the 20 KB slot: BW <= 16
    if (Win >= 16) { P.BW = 16; P.BH = 8; } else { P.BW = 8; P.BH = 16; }
    P.rows_a_bytes = (P.BH + 2) * P.BW * 128;
  }
  P.logBW = 0;
  while ((1 << P.logBW) < P.BW) P.logBW++;
  P.tilesW = (Win + P.BW - 1) / P.BW;
  P.tilesH = (Hin + P.BH - 1) / P.BH;
  // Small problems (NAFNet's 16x16 levels, sharded deep UNet levels): fewer tiles than half the SMs means a handful
  // of CTAs stream all the weights; narrower N tiles spread that over more SMs.  Tap mode 0 accumulates in the same
  // K order for every BN, so results do not depend on this choice (and hence not on the batch size).
  if (!rows && !P.qsm && !nchw_out) {
    const int tm = P.tilesW * P.tilesH * B * nphases;
    while (BN > 64 && Cout % (BN / 2) == 0 && (long long)tm * ((Cout + BN - 1) / BN) < 74) BN >>= 1;
  }
  // CTA-pair variant: per-tap mode, full 256-channel N tiles, plain bf16 NHWC output through TMA stores
  // (not the stride-2 space-to-depth layers: 16 taps over 4 planes measured 10 % slower as pairs)
  d->pair = g_pair && g_tma_store && !rows && BN == 256 && Cout % 256 == 0 && !P.qsm && !P.w_per_image &&
            !nchw_out && planes == 1 && ntaps > 1;
  d->pair_tiles = ((P.tilesW * P.tilesH * B * nphases + 1) / 2) * (Cout / 256);
  d->BN = BN;
  d->grid = dim3((unsigned)(P.tilesW * P.tilesH * B * nphases), (unsigned)((Cout + BN - 1) / BN));
  d->tiles_m = P.tilesW * P.tilesH * B * nphases;
  d->num_tiles = d->tiles_m * ((Cout + BN - 1) / BN);
  // Tile order.  Default: M fastest - the CTAs of a wave share one N tile, i.e. one set of weight tiles (3x3 layers: the
  // weights of an N tile are MBs, the A tensor of the deep levels fits in L2).  1x1 convs with several N tiles (to_qkv:
  // 384 outputs = 3 tiles) invert that: their weights are a few KB but the A tensor of the 256^2 / 128^2 levels is larger
  // than L2, and M-fastest re-read it from DRAM once per N tile (ncu r02: 403 MB read for a 134 MB input).  N fastest
  // makes the N tiles of one pixel tile run back to back in the same wave, so the second and third read hit L2.
  P.tiles_n = (Cout + BN - 1) / BN;
  // (only where the input is too large to stay in L2 anyway; small inputs keep M-fastest and its once-per-N-tile epilogue table)
  const long long a_bytes = (long long)B * Hin * Win * in_pitch * 2;
  P.n_fast = (ntaps == 1 && P.tiles_n > 1 && !P.w_per_image && a_bytes > (24ll << 20)) ? 1 : 0;
  // A: [planes][B][H][W][C] (C contiguous, pixel pitch in_pitch)
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B, (cuuint64_t)planes};
    cuuint64_t strides[4] = {(cuuint64_t)in_pitch * 2, (cuuint64_t)Win * in_pitch * 2, (cuuint64_t)Hin * Win * in_pitch * 2,
                             (cuuint64_t)B * Hin * Win * in_pitch * 2};
    if (planes < 0) {
      // 7x7 stem "row trick": the input is a zero-bordered [B][H+6][W+8][8ch] buffer; one GEMM K-chunk of
      // 64 elements = 8 consecutive pixels x 8 channels of one input row (overlapping pixel stride of 16 B),
      // so kernel row r is a single tap (dh=r) and K = 7 x 64.
      dims[0] = 64; dims[1] = (cuuint64_t)Win; dims[2] = (cuuint64_t)(Hin + 6); dims[3] = (cuuint64_t)B; dims[4] = 1;
      strides[0] = 16; strides[1] = (cuuint64_t)(Win + 8) * 16; strides[2] = (cuuint64_t)(Hin + 6) * (Win + 8) * 16;
      strides[3] = strides[2] * B;
    }
    cuuint32_t box[5] = {64, (cuuint32_t)P.BW, (cuuint32_t)P.BH, 1, 1};
    if (rows) { box[2] = (cuuint32_t)(P.BH + 2); }
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(&d->map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)in, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete d;
      if (err) { char b[96]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(A) failed: %d", (int)r); *err = b; }
      return nullptr;
    }
  }
  // B: [phase*ntaps][Cout][Cin]
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout_w, (cuuint64_t)(P.w_per_image ? B : nphases * ntaps)};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout_w * Cin * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)(d->pair ? BN / 2 : BN), 1};  // pair: each CTA loads its half of the channels
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode(&d->map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)wpacked, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete d;
      if (err) { char b[96]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(B) failed: %d", (int)r); *err = b; }
      return nullptr;
    }
  }
  // output maps for the TMA-store epilogue: [B][H][W][C] view of the destination (channel offset already in
  // `out`), pixel stride os*pitch; one map per upsample phase (base shifted by (py,px))
  memset(&d->map_o, 0, sizeof d->map_o);
  P.tma_store = 0;
  if (!nchw_out && g_tma_store) {
    const int bw = P.BW < 32 ? P.BW : 32, bh = 32 / bw;
    bool ok = true;
    for (int ph = 0; ph < nphases && ok; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      bf16* base = out + ((long long)py * Wout + px) * out_pitch;
      cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)P.os * out_pitch * 2, (cuuint64_t)P.os * Wout * out_pitch * 2,
                               (cuuint64_t)Hout * Wout * out_pitch * 2};
      cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = g_encode(&d->map_o.m[ph], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      ok = (r == CUDA_SUCCESS);
    }
    P.tma_store = ok ? 1 : 0;
  }
  tc_fill_magic(d);
  return d;
}

// fp32x3 mode: `in_split` is the conv input split by launch_split_tf32: [2 (hi, lo)][planes][B][Hin][Win][Cin] fp32 (dense,
// pixel pitch Cin); `wsplit` is [phase*ntaps][2 (hi, lo)][Cout_w][Cin] fp32 (pack_tc3_* kernels); fp32 NHWC output view.
TcConvDesc* tc_conv_create_f32x3(const float* in_split, int B, int Hin, int Win, int Cin, int planes, const float* wsplit, int Cout,
                                 int ntaps, const TcTap* taps, int nphases, const Epilogue& ep, float* out, int out_pitch, int Hout,
                                 int Wout, std::string* err) {
  auto bad = [&](const char* m) -> TcConvDesc* {
    if (err) *err = m;
    return nullptr;
  };
  if (!g_encode) return bad("tc_init not called");
  if (ntaps < 1 || ntaps > 16) return bad("ntaps out of range");
  if (planes < 1) return bad("fp32x3: planes must be >= 1");
  const bool nchw_out = (out == nullptr);
  const int Cout_w = (Cout + 7) / 8 * 8;
  if (Cin % 4) return bad("fp32x3: input channels must be a multiple of 4");
  if (!nchw_out && (Cout % 4 || out_pitch % 4)) return bad("fp32x3: output channels / pitch must be multiples of 4");
  if (((uintptr_t)in_split & 15) || ((uintptr_t)wsplit & 15) || ((uintptr_t)out & 15)) return bad("operands must be 16-byte aligned");
  if (ep.res && (((uintptr_t)ep.res & 15) || ep.res_pitch % 4)) return bad("residual must be 16-byte aligned");
  TcConvDesc* d = new TcConvDesc();
  TcParams& P = d->P;
  memset(&P, 0, sizeof P);
  P.B = B; P.H = Hin; P.W = Win; P.Cin = Cin; P.Cout = Cout;
  P.ntaps = ntaps; P.kchunks = (Cin + 31) / 32;
  P.nphases = nphases;
  P.Hout = Hout; P.Wout = Wout; P.os = nphases == 4 ? 2 : 1;
  P.out = nullptr; P.out_f32 = out; P.out_pitch = out_pitch;
  P.res = nullptr; P.res_f32 = (const float*)ep.res; P.res_pitch = ep.res_pitch;
  P.bias = ep.bias; P.mult_vec = ep.mult_vec; P.ss = ep.ss; P.t_ptr = ep.t_ptr; P.ss_S = ep.ss_S; P.ss_off = ep.ss_off;
  P.ss_img_stride = ep.ss_img_stride; P.silu = ep.silu;
  P.lo_plane_off = planes;
  for (int i = 0; i < ntaps; ++i) P.taps[i] = taps[i];
  int BN = Cout > 64 ? 128 : (Cout > 32 ? 64 : 32);
  d->mode = 3;
  long long best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    int bh = 128 / bw;
    long long cost = (long long)((Win + bw - 1) / bw) * bw * ((Hin + bh - 1) / bh) * bh;
    if (best < 0 || cost < best) { best = cost; P.BW = bw; P.BH = bh; }
  }
  P.logBW = 0;
  while ((1 << P.logBW) < P.BW) P.logBW++;
  P.tilesW = (Win + P.BW - 1) / P.BW;
  P.tilesH = (Hin + P.BH - 1) / P.BH;
  {
    const int tm = P.tilesW * P.tilesH * B * nphases;
    while (BN > 32 && Cout % (BN / 2) == 0 && (long long)tm * ((Cout + BN - 1) / BN) < 74) BN >>= 1;
  }
  d->BN = BN;
  d->grid = dim3((unsigned)(P.tilesW * P.tilesH * B * nphases), (unsigned)((Cout + BN - 1) / BN));
  d->tiles_m = P.tilesW * P.tilesH * B * nphases;
  d->num_tiles = d->tiles_m * ((Cout + BN - 1) / BN);
  P.tiles_n = (Cout + BN - 1) / BN;
  P.n_fast = 0;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B, (cuuint64_t)(2 * planes)};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 4, (cuuint64_t)Win * Cin * 4, (cuuint64_t)Hin * Win * Cin * 4,
                             (cuuint64_t)B * Hin * Win * Cin * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)P.BW, (cuuint32_t)P.BH, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(&d->map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)in_split, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete d;
      if (err) { char b[96]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(A fp32) failed: %d", (int)r); *err = b; }
      return nullptr;
    }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout_w, (cuuint64_t)(nphases * ntaps * 2)};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)Cout_w * Cin * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode(&d->map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)wsplit, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete d;
      if (err) { char b[96]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled(B fp32) failed: %d", (int)r); *err = b; }
      return nullptr;
    }
  }
  memset(&d->map_o, 0, sizeof d->map_o);
  P.tma_store = 0;
  tc_fill_magic(d);
  return d;
}

void tc_conv_destroy(TcConvDesc* d) { delete d; }

bool tc_fused_attention_available() { return g_tma_store; }

void tc_conv_set_out_nchw(TcConvDesc* d, float* out, int cropH, int cropW) {
  d->P.out_nchw = out;
  d->P.cropH = cropH;
  d->P.cropW = cropW;
}

void tc_conv_set_runtime(TcConvDesc* d, const float* ss, const int* t_ptr, int ss_img_stride) {
  d->P.ss = ss;
  d->P.t_ptr = t_ptr;
  d->P.ss_img_stride = ss_img_stride;
  if (ss && d->P.epi_kind != EK_GENERIC && !(d->P.epi_kind & EK_AFF)) {  // the time-modulation table arrives at launch time
    const int k = d->P.epi_kind | EK_AFF;
    d->P.epi_kind = (k == EK_AFF || k == (EK_AFF | EK_SILU) || k == (EK_AFF | EK_RES)) ? k : EK_GENERIC;
  }
}

int tc_conv_launch(TcConvDesc* d, cudaStream_t st) {
  {
    const unsigned g = (unsigned)(d->num_tiles < g_num_sms ? d->num_tiles : g_num_sms);
    static unsigned long long* dbg_dev = nullptr;
    const char* dbg_env = getenv("IRSDE_TC_DEBUG");
    const bool dbg = dbg_env && dbg_env[0] == '1';
    if (dbg) {
      if (!dbg_dev) cudaMalloc(&dbg_dev, 256 * 8 * sizeof(unsigned long long));
      cudaMemsetAsync(dbg_dev, 0, 256 * 8 * sizeof(unsigned long long), st);
    }
    d->P.dbg = dbg ? dbg_dev : nullptr;
    if (d->pair) {
      const int ncl = d->pair_tiles < g_num_sms / 2 ? d->pair_tiles : g_num_sms / 2;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3(2 * ncl);
      cfg.blockDim = dim3(320);
      cfg.dynamicSmemBytes = TcCfgP<256, 0, 2>::SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const int ekp = g_epi_spec >= 3 ? d->P.epi_kind : EK_GENERIC;
      cudaError_t le;
      if (ekp == (EK_AFF | EK_SILU))
        le = cudaLaunchKernelEx(&cfg, conv_tc_persist_kernel<256, 0, 2, EK_AFF | EK_SILU>, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->pair_tiles);
      else if (ekp == (EK_SILU | EK_RES))
        le = cudaLaunchKernelEx(&cfg, conv_tc_persist_kernel<256, 0, 2, EK_SILU | EK_RES>, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->pair_tiles);
      else if (ekp == EK_AFF)
        le = cudaLaunchKernelEx(&cfg, conv_tc_persist_kernel<256, 0, 2, EK_AFF>, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->pair_tiles);
      else if (ekp == 0)
        le = cudaLaunchKernelEx(&cfg, conv_tc_persist_kernel<256, 0, 2, 0>, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->pair_tiles);
      else
        le = cudaLaunchKernelEx(&cfg, conv_tc_persist_kernel<256, 0, 2>, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->pair_tiles);
      return le == cudaSuccess ? 1 : -1;
    }
#define TC_LAUNCH(BNV, PV) \
  pdl_launch(conv_tc_persist_kernel<BNV, PV>, g, 320, TcCfgP<BNV, PV>::SMEM_BYTES, st, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->num_tiles)
    if (d->mode == 3) {
      switch (d->BN) {
        case 32: TC_LAUNCH(32, 3); break;
        case 64: TC_LAUNCH(64, 3); break;
        default: TC_LAUNCH(128, 3); break;
      }
    } else if (d->mode == 2) {
#define TC_LAUNCH_EK2(BNV, EKV) \
  pdl_launch(conv_tc_persist_kernel<BNV, 2, 1, EKV>, g, 320, TcCfgP<BNV, 2>::SMEM_BYTES, st, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->num_tiles)
      const int ek2 = g_epi_spec >= 2 ? d->P.epi_kind : EK_GENERIC;
      switch (d->BN) {
        case 32: TC_LAUNCH(32, 2); break;
        case 64:
          if (ek2 == (EK_AFF | EK_SILU)) TC_LAUNCH_EK2(64, EK_AFF | EK_SILU);
          else if (ek2 == (EK_SILU | EK_RES)) TC_LAUNCH_EK2(64, EK_SILU | EK_RES);
          else TC_LAUNCH(64, 2);
          break;
        default:
          if (ek2 == (EK_AFF | EK_SILU)) TC_LAUNCH_EK2(128, EK_AFF | EK_SILU);
          else if (ek2 == (EK_SILU | EK_RES)) TC_LAUNCH_EK2(128, EK_SILU | EK_RES);
          else TC_LAUNCH(128, 2);
          break;
      }
#undef TC_LAUNCH_EK2
    } else {
#define TC_LAUNCH_EK(BNV, EKV) \
  pdl_launch(conv_tc_persist_kernel<BNV, 0, 1, EKV>, g, 320, TcCfgP<BNV, 0>::SMEM_BYTES, st, d->map_a, d->map_b, d->map_o, d->P, d->tiles_m, d->num_tiles)
      const int ek = g_epi_spec >= 1 ? d->P.epi_kind : EK_GENERIC;
      switch (d->BN) {
        case 32: TC_LAUNCH(32, 0); break;
        case 64:
          if (ek == 0) TC_LAUNCH_EK(64, 0);
          else if (ek == EK_AFF) TC_LAUNCH_EK(64, EK_AFF);
          else TC_LAUNCH(64, 0);
          break;
        case 128:
          if (ek == 0) TC_LAUNCH_EK(128, 0);
          else if (ek == EK_AFF) TC_LAUNCH_EK(128, EK_AFF);
          else if (ek == EK_QSM) TC_LAUNCH_EK(128, EK_QSM);
          else TC_LAUNCH(128, 0);
          break;
        default:
          if (g_epi_spec >= 3 && ek == EK_AFF) TC_LAUNCH_EK(256, EK_AFF);   // downsample convs / deep to_out (bias)
          else if (g_epi_spec >= 3 && ek == 0) TC_LAUNCH_EK(256, 0);       // deep res_conv
          else TC_LAUNCH(256, 0);
          break;
      }
#undef TC_LAUNCH_EK
    }
#undef TC_LAUNCH
    if (dbg) {
      unsigned long long h[256 * 8];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, dbg_dev, sizeof h, cudaMemcpyDeviceToHost);
      double a[8] = {0};
      for (unsigned i = 0; i < g; ++i) for (int k = 0; k < 8; ++k) a[k] += (double)h[i * 8 + k] / g;
      fprintf(stderr, "TCDBG BN=%d mode=%d grid=%u tiles=%d ntaps=%d kch=%d | prod wait %.0f / %.0f | mma wait full %.0f tempty %.0f / %.0f | epi wait tfull %.0f / %.0f (tiles/cta %.1f)\n",
              d->BN, d->mode, g, d->num_tiles, d->P.ntaps, d->P.kchunks, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
    }
    return 1;
  }
}

}  // namespace irsde
