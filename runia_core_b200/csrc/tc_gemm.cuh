// 3xTF32 tensor-core contraction for sm_100a: tcgen05.mma (kind::tf32) with FP32 accumulators in
// TMEM, operands staged in 128B-swizzled shared memory.
//
//   acc[m, n] = sum_k f(A[m0+m, k]) * B[n0+n, k]       A: [M, K] fp32 (streamed), B: [NB, K] fp32
//
// FP32-faithful by operand splitting: x = hi + lo with hi = tf32(x), lo = tf32(x - hi); the kernel
// issues  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the dropped lo*lo term is ~2^-22 relative).
//
// The kernel runs as CTA PAIRS (thread-block cluster of 2, tcgen05 cta_group::2): one UMMA covers a
// 256 x 256 output panel, CTA r of the pair owns rows [128 r, 128 r + 128) of it (its A tile, its
// TMEM accumulator, its epilogue) and stages HALF of the B panel (128 of the 256 B rows) in its own
// shared memory.  Per MMA each SM therefore reads 4 KB of A and 4 KB of B instead of 4 + 8, the
// L2 -> SM traffic for B halves, and a stage shrinks to 64 KB so that three stages fit.
//
// Roles inside one CTA (512 threads, 1 CTA / SM):
//   warp 0      TMA producer for this CTA's half of the two B planes (pre-split in HBM at setup
//               time); completion bytes are posted on the LEADER CTA's `full` barrier
//   warp 1      MMA issuer (leader CTA only, one elected lane): 12 tcgen05.mma.cta_group::2 per
//               32-wide k-block; tcgen05.commit multicasts the `empty` / `tmem_full` arrivals to both CTAs
//   warp 2      TMEM allocator (512 columns = two 128 x 256 accumulators per CTA, double buffered)
//   warps 4-7   epilogue: tcgen05.ld 32 columns at a time; thread t owns output row t, so every
//               row-wise reduction (sum of squares, log-sum-exp, top-k filter) is thread-local
//   warps 8-15  A converters.  The TMA warp streams the RAW fp32 A tile of a k-block straight into the
//               stage's A_hi plane (128B-swizzled by the TMA unit, i.e. already in UMMA layout, rows
//               and columns out of bounds zero-filled); the converters read it back with LDS.128,
//               apply the fused prologue (subtract centre, clip), split into hi / lo, overwrite the
//               hi plane in place and write the lo plane.  The streamed operand is therefore read
//               from HBM exactly once, as raw fp32, by the copy engine: no thread of the CTA has a
//               global load in flight, so the fence.proxy.async each converter needs before handing
//               the stage to the tensor core (a MEMBAR in SASS) never waits on HBM latency.
// Pipelines: smem ring (full/empty mbarriers; the leader's `full` collects the TMA bytes of both
// CTAs and one arrival per converter warp of both CTAs), TMEM ring (tmem_full via tcgen05.commit,
// the leader's tmem_empty collects the epilogue threads of both CTAs).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace runia {
namespace tc {

constexpr int TM = 128;        // rows per CTA
constexpr int TM2 = 256;       // rows per CTA pair (UMMA M with cta_group::2)
constexpr int TN = 256;        // columns per panel (UMMA N)
constexpr int TNH = TN / 2;    // B rows staged by each CTA of the pair
constexpr int TK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int THREADS = 512;
constexpr int CONV_THREADS = 256;  // warps 8..15
constexpr int CONV_WARPS = CONV_THREADS / 32;
constexpr int A_PLANE_BYTES = TM * TK * 4;  // 16 KB
constexpr int B_PLANE_BYTES = TNH * TK * 4;  // 16 KB (this CTA's half of the panel)
constexpr int STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * B_PLANE_BYTES;  // 64 KB
constexpr int SMEM_BAR_BYTES = 256;
constexpr int EPI_STAGE_BYTES = 4 * 4096;  // one 32 x 32 fp32 sub-tile per epilogue warp (TMA-store staging)
constexpr int SMEM_ALIGN = 1024;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Blocking wait on a (local) mbarrier phase.  The suspend-time hint lets the hardware park the
// thread until the phase flips (or the hint expires) instead of spinning: spinning waiters share
// issue slots with the MMA issuer and the epilogue warps of the same SM sub-partition.
// Default (.acquire.cta) semantics on purpose, also for barriers the peer CTA signals: what crosses
// the pair is shared / tensor memory handed to the async proxy (ordered by fence.proxy.async resp.
// tcgen05.fence on the producer side), never global memory, and a cluster-scope acquire / release
// costs an L1 invalidate (CCTL.IVALL) resp. MEMBAR.GPU that would drain the converters' prefetch.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive on a barrier given by its shared::cluster address (own or peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// 2-CTA TMA load: the box lands in THIS CTA's shared memory, the bytes are posted on `leader_bar`
// (shared::cluster address of the pair leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t leader_bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar), "r"(x), "r"(y)
      : "memory");
}
// plain TMA load into this CTA's shared memory, bytes posted on this CTA's own barrier
__device__ __forceinline__ void tma_load_2d_local(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// pull a box into L2 ahead of its load (no shared memory, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// executed by the same warp of BOTH CTAs of the pair
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem of both CTAs] (+)= A[smem of both CTAs] * B[smem halves of both CTAs], kind::tf32; leader only
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread t <-> lane base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzle shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: next 8-row group
  d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=256 (pair), N=n
template <int N>
struct InstrDesc {
  static constexpr uint32_t value = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TM2 >> 4) << 24);
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct Prologue {
  const float *sub;  // [K] or nullptr
  float clip;        // +inf = none
};

// Work assigned to one CTA PAIR: row tiles (256 rows each)  tile_first, tile_first + tile_step, ...
// < tile_end, each crossed with the B panels [panel_lo, panel_hi).  Row scorers run persistently
// (one pair per TPC, tile_step = number of pairs) so that the smem ring, the TMEM double buffer and
// the converters' prefetch keep flowing across tiles; kNN / KDE give each pair one row tile and one
// bank split.  Both CTAs of a pair must be given the SAME Work.
struct Work {
  int64_t tile_first, tile_end, tile_step;
  int panel_lo, panel_hi;
};

// The epilogue policy E provides:
//   __device__ void set_stage(uint32_t smem)                     -- this warp's 4 KB staging buffer (once)
//   __device__ void begin(int row_in_tile, int64_t row)          -- once per thread per row tile
//   __device__ void consume(int64_t col0, const float (&v)[32], int warp_in_epi, int lane)
//       32 consecutive columns col0.. of this thread's row
//   __device__ void panel_done(int panel)                        -- after a whole panel
//   __device__ void finish()                                     -- after the last panel of a row tile
// TNV = panel width (UMMA N): 256 for the contraction-heavy scorers; 32 for the narrow linear heads (ReAct / DICE:
// C <= 32 classes), where the B planes are 16 rows per CTA, the MMAs are a small fraction of a stage and the
// pipeline runs at the speed of the TMA stream and the converters.  The TMEM double buffer keeps its 256-column stride.
// NCAT (narrow panels only): the B_hi and B_lo planes of a stage are adjacent, so A_hi x [B_hi | B_lo] is ONE UMMA of
// N = 2 TNV and A_lo x B_hi a second one into its own accumulator -- 2 instead of 3 instructions per K step (at N = 32
// an instruction costs its fixed minimum, not its FLOPs).  Accumulator columns: [0, TNV/2) hi x classes of CTA 0,
// [TNV/2, TNV) lo x the same classes, [TNV, 3TNV/2) hi x classes of CTA 1, [3TNV/2, 2TNV) lo x those; [2TNV, 3TNV)
// A_lo x B_hi in class order.  The epilogue receives all 3 TNV columns and adds the three products.
// NPROD = 1 (the kNN seed and candidate filter): a single TF32 product A_hi x B_hi -- a third of the MMAs, no B_lo
// traffic, no A_lo store; the caller widens whatever it derives from the result by the split's rounding bound (see
// KnnSeedEpi).
// DIRECT (with NPROD = 1, no prologue): the raw fp32 A tile IS the operand -- kind::tf32 reads the top 19 bits of each
// 32-bit word, which is the truncation the converters would have written -- so the TMA load completes on the leader's
// `full` barrier like the B planes, the converter warps have nothing to do, and a stage is one A plane + one B plane
// (32 KB: twice the stages in the same shared memory, which is what hides the L2 -> SM latency once a stage holds only
// 4 MMAs).
template <class E, int TNV = TN, int NSTAGES = STAGES, bool NCAT = false, int NPROD = 3, bool DIRECT = false>
__device__ __forceinline__ void run_tiles(const CUtensorMap *tmA, int K, const Prologue pro,
                                          const CUtensorMap *tmB_hi, const CUtensorMap *tmB_lo, const Work work,
                                          E &epi, unsigned char *smem_raw) {
  // stage = raw/hi A plane, lo A plane, this CTA's halves of the two B planes (narrow panels: smaller B planes, more stages)
  constexpr int BPLANE = (TNV / 2) * TK * 4;
  static_assert(!DIRECT || NPROD == 1, "the raw tile can only stand in for A_hi");
  constexpr int STG = DIRECT ? A_PLANE_BYTES + BPLANE : 2 * A_PLANE_BYTES + 2 * BPLANE;
  constexpr int B_OFF = DIRECT ? A_PLANE_BYTES : 2 * A_PLANE_BYTES;
  static_assert(!NCAT || 3 * TNV <= TN, "concatenated panels must fit one accumulator slot");
  static_assert(BPLANE % 1024 == 0, "B planes must keep the 1024-byte alignment of the 128B swizzle");
  static_assert(8 * (3 * NSTAGES + 4) + 8 <= SMEM_BAR_BYTES, "barrier block too small");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  // carve shared memory (1024-byte aligned for the 128B swizzle); identical offsets in both CTAs
  const uint32_t base = (smem_u32(smem_raw) + SMEM_ALIGN - 1) & ~(uint32_t)(SMEM_ALIGN - 1);
  unsigned char *gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t epi_stage0 = base + NSTAGES * STG;  // 1024-byte aligned (128B-swizzled boxes)
  const uint32_t bar0 = epi_stage0 + EPI_STAGE_BYTES;
  auto sA_hi = [&](int s) { return base + s * STG; };
  auto sA_lo = [&](int s) { return base + s * STG + A_PLANE_BYTES; };
  auto sB_hi = [&](int s) { return base + s * STG + B_OFF; };
  auto sB_lo = [&](int s) { return base + s * STG + B_OFF + BPLANE; };  // not used when NPROD == 1
  auto full_bar = [&](int s) { return bar0 + 8 * s; };                      // used in the leader only
  auto empty_bar = [&](int s) { return bar0 + 8 * (NSTAGES + s); };          // one per CTA
  auto tfull_bar = [&](int b) { return bar0 + 8 * (2 * NSTAGES + b); };      // one per CTA
  auto tempty_bar = [&](int b) { return bar0 + 8 * (2 * NSTAGES + 2 + b); }; // used in the leader only
  auto raw_bar = [&](int s) { return bar0 + 8 * (2 * NSTAGES + 4 + s); };    // one per CTA: raw A tile landed
  const uint32_t tmem_slot = bar0 + 8 * (3 * NSTAGES + 4);
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(gbase + NSTAGES * STG + EPI_STAGE_BYTES + 8 * (3 * NSTAGES + 4));
  const uint32_t sub_smem = bar0 + SMEM_BAR_BYTES;  // [nkb * TK] floats: the centre, zero-padded
  float *sub_ptr = reinterpret_cast<float *>(gbase + NSTAGES * STG + EPI_STAGE_BYTES + SMEM_BAR_BYTES);

  const int nkb = (K + TK - 1) / TK;
  const int n_panels = work.panel_hi - work.panel_lo;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmA);
    tma_prefetch_desc(tmB_hi);
    tma_prefetch_desc(tmB_lo);
  }
  if (pro.sub)
    for (int k = threadIdx.x; k < nkb * TK; k += THREADS) sub_ptr[k] = k < K ? __ldg(pro.sub + k) : 0.f;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NSTAGES; ++s) {
      mbar_init(full_bar(s), DIRECT ? 1 : 1 + 2 * CONV_WARPS);  // leader's expect_tx arrival (+ converter warps of both CTAs)
      mbar_init(empty_bar(s), 1);                  // tcgen05.commit (multicast)
      mbar_init(raw_bar(s), 1);                    // this CTA's TMA warp (expect_tx)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);         // tcgen05.commit (multicast)
      mbar_init(tempty_bar(b), 2 * 128);  // epilogue threads of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  __syncwarp();
  tc_fence_before();
  __syncthreads();  // sub_ptr visible to the converters
  cluster_sync_all();  // barriers of both CTAs initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // -------------- TMA producer (this CTA's raw A tile and its half of the B planes) --------------
    if (lane == 0) {
      int it = 0;
      for (int64_t t = work.tile_first; t < work.tile_end; t += work.tile_step) {
        const int row0 = (int)(t * TM2 + (int64_t)rank * TM);
        for (int p = 0; p < n_panels; ++p) {
          const int n0 = (work.panel_lo + p) * TNV + (int)rank * (TNV / 2);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % NSTAGES;
            const uint32_t ph = (it / NSTAGES) & 1;
            // L2 prefetch one row tile ahead: with 3 stages in flight an HBM round trip is most of the slack the
            // ring has, a load that hits L2 is not (the TF32 probe sustains 1.07 PFLOP/s, the HBM-fed scorers 0.89,
            // the L2-fed multi-panel DDU 1.04)
            if (p == 0 && t + work.tile_step < work.tile_end)
              tma_prefetch_2d(tmA, kb * TK, (int)((t + work.tile_step) * TM2 + (int64_t)rank * TM));
            mbar_wait(empty_bar(s), ph ^ 1);
            const uint32_t lbar = map_to_cta(full_bar(s), 0);
            if constexpr (DIRECT) {
              if (rank == 0) mbar_expect_tx(full_bar(s), 2 * A_PLANE_BYTES + 2 * BPLANE);  // A tile + B_hi half of both CTAs
              tma_load_2d_pair(sA_hi(s), tmA, lbar, kb * TK, row0);
            } else {
              mbar_expect_tx(raw_bar(s), A_PLANE_BYTES);
              tma_load_2d_local(sA_hi(s), tmA, raw_bar(s), kb * TK, row0);
              if (rank == 0) mbar_expect_tx(full_bar(s), (NPROD == 1 ? 2 : 4) * (TNV / 2) * TK * 4);  // plane(s) of both CTAs
            }
            tma_load_2d_pair(sB_hi(s), tmB_hi, lbar, kb * TK, n0);
            if constexpr (NPROD != 1) tma_load_2d_pair(sB_lo(s), tmB_lo, lbar, kb * TK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA) ------------------------------
    if (rank == 0) {
      int it = 0, pc = 0;
      for (int64_t t = work.tile_first; t < work.tile_end; t += work.tile_step)
      for (int p = 0; p < n_panels; ++p, ++pc) {
        const int ab = pc & 1;
        const uint32_t aph = (pc >> 1) & 1;
        mbar_wait(tempty_bar(ab), aph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ab * TN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % NSTAGES;
          const uint32_t ph = (it / NSTAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t dA_hi = make_smem_desc(sA_hi(s)), dA_lo = make_smem_desc(sA_lo(s));
            const uint64_t dB_hi = make_smem_desc(sB_hi(s)), dB_lo = make_smem_desc(sB_lo(s));
#pragma unroll
            for (int k4 = 0; k4 < TK / 8; ++k4) {
              const uint64_t adv = (uint64_t)((k4 * 8 * 4) >> 4);  // 32 bytes per UMMA_K=8 step
              if constexpr (NPROD == 1) {
                umma_tf32(tacc, dA_hi + adv, dB_hi + adv, InstrDesc<TNV>::value, (kb | k4) != 0 ? 1u : 0u);
              } else if constexpr (NCAT) {
                umma_tf32(tacc, dA_hi + adv, dB_hi + adv, InstrDesc<2 * TNV>::value, (kb | k4) != 0 ? 1u : 0u);
                umma_tf32(tacc + 2 * TNV, dA_lo + adv, dB_hi + adv, InstrDesc<TNV>::value, (kb | k4) != 0 ? 1u : 0u);
              } else {
                umma_tf32(tacc, dA_lo + adv, dB_hi + adv, InstrDesc<TNV>::value, (kb | k4) != 0 ? 1u : 0u);
                umma_tf32(tacc, dA_hi + adv, dB_lo + adv, InstrDesc<TNV>::value, 1u);
                umma_tf32(tacc, dA_hi + adv, dB_hi + adv, InstrDesc<TNV>::value, 1u);
              }
            }
            umma_commit(empty_bar(s));                       // smem slot reusable (both CTAs) when these MMAs retire
            if (kb == nkb - 1) umma_commit(tfull_bar(ab));   // accumulators ready for both epilogues
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------ epilogue ------------------------------
    const int ew = warp - 4;  // == warp % 4: TMEM lane quadrant this warp may read
    const int row_in_tile = ew * 32 + lane;
    int pc = 0;
    epi.set_stage(epi_stage0 + (uint32_t)ew * 4096u);
    for (int64_t t = work.tile_first; t < work.tile_end; t += work.tile_step) {
    epi.begin(row_in_tile, t * TM2 + (int64_t)rank * TM + row_in_tile);
    for (int p = 0; p < n_panels; ++p, ++pc) {
      const int ab = pc & 1;
      const uint32_t aph = (pc >> 1) & 1;
      mbar_wait(tfull_bar(ab), aph);
      tc_fence_after();
      const int64_t n0 = (int64_t)(work.panel_lo + p) * TNV;
#pragma unroll 1
      for (int c0 = 0; c0 < (NCAT ? 3 * TNV : TNV); c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * TN + c0), v);
        epi.consume(n0 + c0, v, ew, lane);
      }
      tc_fence_before();
      mbar_arrive_cluster(map_to_cta(tempty_bar(ab), 0));
      epi.panel_done(work.panel_lo + p);
    }
    epi.finish();
    }
  } else if (warp >= 8) {
    // ------------------------------ A converters ------------------------------
    const int ct = threadIdx.x - 256;  // 0..255
    const int chunk = ct & 7;          // 16-byte chunk (4 floats) inside the 128-byte k-block row
    const int r0 = ct >> 3;            // rows r0, r0+32, r0+64, r0+96 (all share r & 7, hence the swizzle)
    const uint32_t off0 = (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((chunk ^ (r0 & 7)) << 4));
    const int64_t n_my_tiles =
        work.tile_first < work.tile_end ? (work.tile_end - work.tile_first + work.tile_step - 1) / work.tile_step : 0;
    const int64_t n_items = DIRECT ? 0 : n_my_tiles * (int64_t)n_panels * nkb;  // flat sequence of (tile, panel, k-block)
    const uint32_t lfull0 = map_to_cta(full_bar(0), 0);  // leader's full barriers, 8 bytes apart
    const bool has_sub = pro.sub != nullptr;
    const bool has_clip = pro.clip < INFINITY;
    int s = 0, kb = 0;
    uint32_t ph = 0;
    for (int64_t item = 0; item < n_items; ++item) {
      float4 sub = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_sub)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(sub.x), "=f"(sub.y), "=f"(sub.z), "=f"(sub.w)
                     : "r"(sub_smem + (uint32_t)(kb * TK + chunk * 4) * 4));
      mbar_wait(raw_bar(s), ph);  // raw fp32 tile of this k-block is in the A_hi plane
      const uint32_t a_hi = sA_hi(s) + off0, a_lo = sA_lo(s) + off0;
      float4 x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(x[i].x), "=f"(x[i].y), "=f"(x[i].z), "=f"(x[i].w)
                     : "r"(a_hi + (uint32_t)i * 4096u));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float e[4] = {x[i].x - sub.x, x[i].y - sub.y, x[i].z - sub.z, x[i].w - sub.w};
        float h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (has_clip) e[q] = e[q] > pro.clip ? pro.clip : e[q];
          // hi = the top 19 bits (truncation: one LOP3, and e - hi is exact); lo = e - hi rounded to TF32 by adding
          // half an ulp before the mask.  e = hi + lo to 2^-22 |e|, like cvt.rna on both, at 3 ALU instructions
          // instead of 7 (cvt.rna.tf32 is emulated: add, |x| < inf test, select, mask).
          h[q] = __uint_as_float(__float_as_uint(e[q]) & 0xffffe000u);
          l[q] = __uint_as_float((__float_as_uint(e[q] - h[q]) + 0x1000u) & 0xffffe000u);
        }
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + (uint32_t)i * 4096u), "f"(h[0]), "f"(h[1]),
                     "f"(h[2]), "f"(h[3])
                     : "memory");
        if constexpr (NPROD != 1)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + (uint32_t)i * 4096u), "f"(l[0]), "f"(l[1]),
                       "f"(l[2]), "f"(l[3])
                       : "memory");
      }
      fence_proxy_async();  // make the generic-proxy stores visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(lfull0 + 8 * (uint32_t)s);  // one arrival per converter warp
      if (++kb == nkb) kb = 0;
      if (++s == NSTAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  }
  // ------------------------------ teardown ------------------------------
  // no CTA of the pair may exit (or free TMEM) while the other can still signal its barriers
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dynamic shared memory: stages + barriers + the zero-padded centre [ceil(K / TK) * TK floats] + alignment slack
static inline size_t smem_bytes(int K) {
  return (size_t)STAGES * STAGE_BYTES + EPI_STAGE_BYTES + SMEM_BAR_BYTES + (size_t)((K + TK - 1) / TK) * TK * 4 + SMEM_ALIGN;
}
// same for a narrow-panel variant (panel width tn, nstages stages)
inline size_t smem_bytes_variant(int K, int tn, int nstages) {
  return (size_t)nstages * (2 * A_PLANE_BYTES + 2 * (tn / 2) * TK * 4) + EPI_STAGE_BYTES + SMEM_BAR_BYTES +
         (size_t)((K + TK - 1) / TK) * TK * 4 + SMEM_ALIGN;
}
constexpr int kMaxK = 4096;  // keeps the centre within the shared-memory budget

}  // namespace tc
}  // namespace runia
