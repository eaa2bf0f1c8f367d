// bf16 GEMM on the 5th-generation tensor cores: persistent, warp-specialised
//   warp 0        TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring; whole warp in the loop, one elected lane issues)
//   warp 1        MMA issuer     (one elected lane, tcgen05.mma kind::f16, fp32 accumulators in TMEM; the critical path: kept
//                                 to constant-hi / incremented-lo descriptors and incremental ring indices)
//   warps 2..9/13 epilogue       (tcgen05.ld -> fused epilogue -> global / TMA store), overlapped with the next tiles' main
//                                 loops through 2 or 4 TMEM accumulator stages
// Operands may be K-major or MN-major ("transposed") so forward (X W^T), dX (dY W) and the split-K weight gradient
// (dY^T X) all run on the same kernel.  CTAS = 2 instantiations pair the two CTAs of a cluster on 256-row tiles
// (cta_group::2): see the comment above gemm_tc_kernel and choose_tiles() for where they are used.
#include <stdlib.h>
#include "epilogue.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace swin {

// Development builds (-DSWIN_GEMM_PROF, libswin_b200_prof.so): the roles of the first CTA (pair) accumulate the clocks they
// spend blocked on each barrier, read back through swin_debug_gemm_prof().  Compiled out of the product library.
#ifdef SWIN_GEMM_PROF
__device__ unsigned long long g_gemm_prof[16];
#define PROF_DECL unsigned long long prof_w0 = 0, prof_w1 = 0; const long long prof_t0 = clock64();
#define PROF_WAIT(acc, stmt) do { const long long t__ = clock64(); stmt; acc += (unsigned long long)(clock64() - t__); } while (0)
#define PROF_FLUSH(i0, i1, itot) do { if (blockIdx.x < CTAS) { atomicAdd(&g_gemm_prof[(i0)], prof_w0); atomicAdd(&g_gemm_prof[(i1)], prof_w1); \
    atomicAdd(&g_gemm_prof[(itot)], (unsigned long long)(clock64() - prof_t0)); } } while (0)
#else
#define PROF_DECL
#define PROF_WAIT(acc, stmt) stmt
#define PROF_FLUSH(i0, i1, itot)
#endif

constexpr int TBM = 128, TBK = 64;
constexpr int kEpiWarps = 8;                 // generic / class-2 epilogues: two warps per TMEM lane quarter
constexpr int kEpiWarpsMax = 12;             // class 1 (STORE / GELU, ALU-heavy): three warps per lane quarter
__host__ __device__ constexpr int epi_warps(int epi_class) { return epi_class == 1 ? kEpiWarpsMax : kEpiWarps; }
__host__ __device__ constexpr int gemm_threads(int epi_class) { return 64 + 32 * epi_warps(epi_class); }
constexpr int kMaxStages = 8;

struct GemmTcParams {
  int block_n;            // MMA N (multiple of 32, <= 256)
  int n_tiles, m_tiles, splits, kb_total, kb_per_split;
  int K;                  // reduction length: the last k-block may hold fewer than TBK valid columns (e.g. K = 96)
  int stages;
  uint32_t a_bytes, b_bytes;   // TMA bytes per stage (expect_tx)
  uint32_t epi_bytes_per_warp; // epilogue staging per warp in dynamic smem (4 KB; 8 KB for fp32 class-2 double buffers)
  float* colsum_a;             // split-K dW only: bias gradient via an all-ones N=16 MMA into TMEM columns [256,272)
  int tma_epi;                 // 1: STORE/GELU with bf16 outputs go out through TMA stores (tmD / tmD2)
  uint32_t stage_bytes;        // smem stride per stage: a_bytes + b_bytes rounded up to the 1024-byte swizzle-atom alignment
  int ctas;                    // 1, or 2: CTA-pair tiles (cta_group::2, M = 256; b_bytes is this CTA's HALF of the B tile)
  EpiParams epi;
};

// Work decomposition shared by the three warp roles (they must walk the same sequence): unit = (tile, k-split),
// round-robin over the CTAs.  Forward / dX GEMMs have splits == 1.  For the split-K weight gradient the host picks
// `splits` so that tiles * splits fills the 148 SMs in whole rounds (see pick_splits): the units of one round then sweep
// the SAME k-range of every tile at the same time, so an operand line fetched for one tile is an L2 hit for the others.
// (A stream-K split with per-CTA contiguous ranges balances perfectly but staggers the sharers by tens of microseconds;
// measured on the stage-2 dW it re-read the operands 3.4x from DRAM.)
struct WorkIter {
  int tile, kb0, kb1;
  int unit, total_units, stride, splits, kb_per_split, kb_total;
  __device__ __forceinline__ explicit WorkIter(const GemmTcParams& p) {
    kb_total = p.kb_total; splits = p.splits; kb_per_split = p.kb_per_split;
    unit = blockIdx.x / p.ctas; stride = gridDim.x / p.ctas; total_units = p.m_tiles * p.n_tiles * p.splits;   // m_tiles: tiles of 128 * ctas rows
    tile = kb0 = kb1 = 0;
  }
  __device__ __forceinline__ bool next() {
    if (unit >= total_units) return false;
    if (splits == 1) { tile = unit; kb0 = 0; kb1 = kb_total; }
    else {
      const int ks = unit % splits;
      tile = unit / splits;
      kb0 = ks * kb_per_split;
      kb1 = min(kb_total, kb0 + kb_per_split);
    }
    unit += stride;
    return true;
  }
};

// One 32x32 accumulator chunk in the coalesced layout: lane = 4 consecutive columns (piece) of rows
// rr = 4*i + rsub, i = 0..7.  All smem/global loads of the 8 rows are issued before any dependent math so
// each lane keeps 8 independent memory operations in flight (the serial version was latency-bound).
template <int EPI>
__device__ __forceinline__ void epi_chunk(const EpiParams& p, const float4* __restrict__ stage, const long long* __restrict__ rowdst,
                                          const float* __restrict__ rowscale, int col, int rsub, int piece) {
  if (col >= p.N) return;
  float4 a[8];
  long long off[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = 4 * i + rsub;
    a[i] = stage[rr * 8 + (piece ^ (rr & 7))];
    const long long dr = rowdst[rr];
    off[i] = dr < 0 ? -1 : dr * p.ldd + col;
  }
  if (p.bias != nullptr) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i].x += b4.x; a[i].y += b4.y; a[i].z += b4.z; a[i].w += b4.w; }
  }
  if (EPI == SWIN_EPI_STORE) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (off[i] >= 0) store4(p.D, p.d_dtype, off[i], a[i]);
  } else if (EPI == SWIN_EPI_GELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (off[i] < 0) continue;
      float4 g, d;
      gelu_fast(a[i].x, &g.x, &d.x); gelu_fast(a[i].y, &g.y, &d.y); gelu_fast(a[i].z, &g.z, &d.z); gelu_fast(a[i].w, &g.w, &d.w);
      store4(p.D, p.d_dtype, off[i], g);
      store4(p.D2, p.d_dtype, off[i], d);
    }
  } else if (EPI == SWIN_EPI_RESIDUAL || EPI == SWIN_EPI_SCATTER_RESIDUAL) {
    // all 8 residual loads are issued unconditionally (clamped address, read-only path) BEFORE any use, so they
    // overlap; a predicated load-then-convert sequence gets serialised by in-order issue
    float4 r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + (off[i] >= 0 ? off[i] : 0)));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (off[i] < 0) continue;
      const float sc = rowscale[4 * i + rsub];
      store4(p.D, SWIN_F32, off[i], make_float4(fmaf(sc, a[i].x, r[i].x), fmaf(sc, a[i].y, r[i].y), fmaf(sc, a[i].z, r[i].z), fmaf(sc, a[i].w, r[i].w)));
    }
  } else if (EPI == SWIN_EPI_DGELU) {
    float4 u[8];
    if (p.d_dtype == SWIN_BF16) {
      uint2 raw[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) raw[i] = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.aux) + (off[i] >= 0 ? off[i] : 0)));
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = make_float4(bf16_lo(raw[i].x), bf16_hi(raw[i].x), bf16_lo(raw[i].y), bf16_hi(raw[i].y));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.aux) + (off[i] >= 0 ? off[i] : 0)));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (off[i] >= 0) store4(p.D, p.d_dtype, off[i], make_float4(a[i].x * u[i].x, a[i].y * u[i].y, a[i].z * u[i].z, a[i].w * u[i].w));
  } else {  // SWIN_EPI_ATOMIC_ADD
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (off[i] >= 0) atomicAdd(reinterpret_cast<float4*>(reinterpret_cast<float*>(p.D) + off[i]), a[i]);
  }
}

// EPI_CLASS: 0 = generic coalesced-lane epilogue; 1 = bf16 STORE/GELU through TMA stores; 2 = DGELU (bf16) / RESIDUAL (fp32):
//            the aux tile is TMA-loaded, updated in place in the TMEM row layout and TMA-stored.
//
// CTAS == 2: the two CTAs of a cluster (one TPC) work on ONE 256 x block_n tile with cta_group::2 MMAs.  Each CTA loads its
// own 128 rows of A and HALF of the B tile (so a k-block costs 16 KB + block_n * 64 B of shared-memory fill per CTA instead
// of 16 KB + block_n * 128 B, and the ring holds more k-blocks), keeps its 128 accumulator rows in its own TMEM and runs
// the epilogue on them.  Protocol (the CUTLASS 2-SM one): every TMA of either CTA posts its bytes on the LEADER's full
// barrier; the leader's MMA thread issues for both and its commits arrive on the empty / accumulator-full barriers of
// BOTH CTAs (multicast); the epilogue warps of both CTAs arrive on the leader's accumulator-empty barrier.
template <bool A_MN, bool B_MN, int EPI_CLASS, int CTAS>
__global__ void __launch_bounds__(gemm_threads(EPI_CLASS), 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const __grid_constant__ CUtensorMap tmD,
                                                                   const __grid_constant__ CUtensorMap tmD2, GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 8];
  __shared__ uint32_t tmem_base_slot;
  constexpr int kEW = epi_warps(EPI_CLASS);          // epilogue warps of this instantiation
  __shared__ __align__(8) uint64_t aux_bars[kEpiWarps][4];
  __shared__ __align__(1024) __nv_bfloat16 ones_tile[(EPI_CLASS == 0 && A_MN && B_MN) ? 64 * 64 : 8];   // all-ones B operand (MN-major)
  __shared__ long long epi_rowdst[kEpiWarps][32];
  __shared__ float epi_rowscale[kEpiWarps][32];

  // 1024-byte aligned operand ring
  uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.stage_bytes;
  const uint32_t tx_bytes = (p.a_bytes + p.b_bytes) * CTAS;      // the leader's full barrier collects both CTAs' loads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  constexpr int TM = TBM * CTAS;                                // rows of a work tile
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[kMaxStages + s]); };
  auto tfull_bar = [&](int s) { return smem_u32(&bars[2 * kMaxStages + s]); };
  auto tempty_bar = [&](int s) { return smem_u32(&bars[2 * kMaxStages + 4 + s]); };
  const bool do_colsum = (EPI_CLASS == 0 && A_MN && B_MN) && p.colsum_a != nullptr;
  // TMEM columns per accumulator stage; with the bias-gradient MMA the unit owns all 512 columns (sum in [256,272))
  const uint32_t acc_stride = do_colsum ? 512u : (p.block_n <= 128 ? 128u : 256u);
  const uint32_t num_acc = 512u / acc_stride;                       // 4, 2 or 1 stages in flight

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEW * CTAS); }
    for (int w = 0; w < kEpiWarps; ++w)
      for (int b = 0; b < 4; ++b) mbar_init(smem_u32(&aux_bars[w][b]), 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (CTAS == 2) cluster_sync_all();               // both CTAs' barriers exist before any remote arrive / TMEM pairing
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(smem_u32(&tmem_base_slot), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(&tmem_base_slot), 512); tmem_relinquish(); }
  }
  if (EPI_CLASS == 0 && A_MN && B_MN) {
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) ones_tile[i] = __float2bfloat16(1.0f);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // everything above touched only shared memory / TMEM and overlapped the previous kernel's tail (PDL); from here on global memory
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================================================== TMA producer
    // whole warp in the loop, one elected lane issues (no per-active-lane retry loops around the TMA instructions)
    {
      uint32_t it = 0, s = 0, ph = 0;
      const uint32_t lbar0 = CTAS == 2 ? mapa_u32(full_bar(0), 0) : full_bar(0);       // (leader's) full barrier of slot 0
      const uint32_t b_chunks = p.b_bytes / 8192u;                                     // MN-major B: 64-column chunks per CTA
      PROF_DECL
      for (WorkIter w(p); w.next();) {
        const int tile = w.tile, kb0 = w.kb0, kb1 = w.kb1;
        const int n_blk = tile % p.n_tiles, m_blk = tile / p.n_tiles;
        // this CTA's rows of A and (CTA pair) its half of the B tile
        const int m0 = m_blk * TM + (int)cta_rank * TBM, n0 = n_blk * p.block_n + (int)cta_rank * (p.block_n / 2) * (CTAS - 1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          if (it != 0 && ++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
          PROF_WAIT(prof_w0, mbar_wait(empty_bar(s), ph ^ 1));
          if (elect_one()) {
            const uint32_t sa = smem0 + s * stage_bytes, sb = sa + p.a_bytes;
            const int k0 = kb * TBK;
            if (CTAS == 1) {
              const uint32_t fb = lbar0 + 8 * s;
              mbar_expect_tx(fb, tx_bytes);
              if (!A_MN) {
                tma_load_2d(sa, &tmA, fb, k0, m0);
              } else {
                tma_load_2d(sa, &tmA, fb, m0, k0);
                tma_load_2d(sa + 8192, &tmA, fb, m0 + 64, k0);
              }
              if (!B_MN) {
                tma_load_2d(sb, &tmB, fb, k0, n0);
              } else {
                for (uint32_t b = 0; b < b_chunks; ++b) tma_load_2d(sb + b * 8192, &tmB, fb, n0 + 64 * (int)b, k0);
              }
            } else {
              if (cta_rank == 0) mbar_expect_tx(full_bar(s), tx_bytes);      // one arrival per phase: the leader's, expecting both CTAs' bytes
              const uint32_t lbar = lbar0 + 8 * s;                           // the leader's full barrier
              if (!A_MN) {
                tma_load_2d_pair(sa, &tmA, lbar, k0, m0);
              } else {
                tma_load_2d_pair(sa, &tmA, lbar, m0, k0);
                tma_load_2d_pair(sa + 8192, &tmA, lbar, m0 + 64, k0);
              }
              if (!B_MN) {
                tma_load_2d_pair(sb, &tmB, lbar, k0, n0);
              } else {
                for (uint32_t b = 0; b < b_chunks; ++b) tma_load_2d_pair(sb + b * 8192, &tmB, lbar, n0 + 64 * (int)b, k0);
              }
            }
          }
          __syncwarp();
        }
      }
      if (CTAS == 2) {
        // producer tail: the leader's multicast commits arrive on THIS CTA's empty barriers; do not leave (and let the CTA
        // retire) before the last arrival of every slot has landed
        const uint32_t n_last = it < (uint32_t)p.stages ? it : (uint32_t)p.stages;
        for (uint32_t j = it - n_last; j < it; ++j) mbar_wait(empty_bar(j % p.stages), (j / p.stages) & 1);
      }
      if (lane == 0) PROF_FLUSH(0 + 8 * cta_rank, 1 + 8 * cta_rank, 2 + 8 * cta_rank);       // [0] producer blocked on empty, [2] producer total
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // The whole warp walks the loop (every lane polls the barriers); ONE elected lane issues.  The issue path is the
    // critical resource of this kernel (measured: the tensor pipe idles between MMAs whenever the issuing thread needs more
    // cycles per tcgen05.mma than the MMA runs), so it is kept to two adds per instruction: the smem descriptors are a
    // constant high word plus a low word that advances by a constant per stage and per 16-wide k-step, stage / phase /
    // accumulator indices are carried incrementally (no runtime divisions), and elect.sync lets ptxas issue the
    // single-thread instructions without its per-active-lane retry loop.
    if (cta_rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(p.block_n, A_MN, B_MN, TM);
      const uint32_t idesc_ones = CTAS == 1 ? umma_idesc_bf16(16, A_MN, B_MN) : umma_idesc_bf16(32, A_MN, B_MN, TM);
      // descriptor = hi:lo with hi = SBO 1024 | version 1 | 128B swizzle; lo = (addr >> 4) | LBO << 16 (16 B K-major, 8192 B MN-major)
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t kStepA = A_MN ? (2048u >> 4) : (32u >> 4), kStepB = B_MN ? (2048u >> 4) : (32u >> 4);
      const uint32_t a_lo0 = ((smem0 & 0x3FFFFu) >> 4) | ((A_MN ? (8192u >> 4) : 1u) << 16);
      const uint32_t b_lo0 = (((smem0 + p.a_bytes) & 0x3FFFFu) >> 4) | ((B_MN ? (8192u >> 4) : 1u) << 16);
      const uint32_t ones_lo = ((smem_u32(ones_tile) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16);
      const uint32_t stage16 = stage_bytes >> 4;
      const uint32_t last_ksteps = (uint32_t)(p.K - (p.kb_total - 1) * TBK + 15) / 16;     // 16-wide k-steps of the last (maybe partial) k-block
      auto desc = [&](uint32_t lo) { return ((uint64_t)kDescHi << 32) | (uint64_t)lo; };
      auto mma = [&](uint32_t d, uint32_t alo, uint32_t blo, uint32_t id, uint32_t accum) {
        if (CTAS == 1) umma_bf16(d, desc(alo), desc(blo), id, accum); else umma_bf16_pair(d, desc(alo), desc(blo), id, accum);
      };
      uint32_t s = 0, ph = 0, acc = 0, acc_ph = 0;
      PROF_DECL
      for (WorkIter w(p); w.next();) {
        const int kb0 = w.kb0, kb1 = w.kb1;
        const bool colsum_unit = do_colsum && (w.tile % p.n_tiles) == 0;      // the epilogue takes the row sums from the n_blk == 0 units
        PROF_WAIT(prof_w0, mbar_wait(tempty_bar(acc), acc_ph ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        for (int kb = kb0; kb < kb1; ++kb) {
          PROF_WAIT(prof_w1, mbar_wait(full_bar(s), ph));
          tc_fence_after();
          if (elect_one()) {
#ifdef SWIN_GEMM_PROF
            const long long tm0 = clock64();
#endif
            const uint32_t a_lo = a_lo0 + s * stage16, b_lo = b_lo0 + s * stage16;
            const uint32_t first = kb > kb0 ? 1u : 0u;
            if (kb + 1 < p.kb_total || last_ksteps == TBK / 16) {
#pragma unroll
              for (uint32_t k = 0; k < TBK / 16; ++k) {
                mma(d_tmem, a_lo + k * kStepA, b_lo + k * kStepB, idesc, k ? 1u : first);
                if (colsum_unit) mma(d_tmem + 256, a_lo + k * kStepA, ones_lo + k * (2048u >> 4), idesc_ones, k ? 1u : first);
              }
            } else {        // partial last k-block: only the k-steps that hold data (TMA zero-fills the rest; K = 96 -> 64 + 32)
              for (uint32_t k = 0; k < last_ksteps; ++k) {
                mma(d_tmem, a_lo + k * kStepA, b_lo + k * kStepB, idesc, k ? 1u : first);
                if (colsum_unit) mma(d_tmem + 256, a_lo + k * kStepA, ones_lo + k * (2048u >> 4), idesc_ones, k ? 1u : first);
              }
            }
#ifdef SWIN_GEMM_PROF
            const long long tm1 = clock64();
#endif
            if (CTAS == 1) umma_commit(empty_bar(s)); else umma_commit_pair(empty_bar(s));          // smem slot reusable once these MMAs retire
            if (kb + 1 == kb1) { if (CTAS == 1) umma_commit(tfull_bar(acc)); else umma_commit_pair(tfull_bar(acc)); }   // accumulator complete
#ifdef SWIN_GEMM_PROF
            if (blockIdx.x == 0) {        // [11] clocks issuing MMAs, [12] clocks issuing commits, [13] k-blocks
              atomicAdd(&g_gemm_prof[11], (unsigned long long)(tm1 - tm0)); atomicAdd(&g_gemm_prof[12], (unsigned long long)(clock64() - tm1));
              atomicAdd(&g_gemm_prof[13], 1ull);
            }
#endif
          }
          __syncwarp();
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
        }
        if (++acc == num_acc) { acc = 0; acc_ph ^= 1; }
      }
      if (lane == 0) PROF_FLUSH(3, 4, 5);         // [3] MMA blocked on accumulator-empty, [4] on operand-full, [5] MMA warp total
    }
    __syncwarp();
  } else {
    // ===================================================== epilogue (4 warps = 128 TMEM lanes)
    // TMEM gives each thread one accumulator ROW; global memory wants a warp per row segment.  Each warp
    // therefore transposes 32x32 fp32 chunks through a private, XOR-swizzled smem stage and runs the fused
    // epilogue in the coalesced layout (lane = 4 consecutive columns of one of 4 rows per instruction).
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int ew = warp - 2;                  // 0..7
    // per-epilogue-warp staging lives in dynamic smem right after the operand ring (4 KB per warp; 2 x 4 KB for class 2)
    float4* stage = reinterpret_cast<float4*>(smem_raw + (smem0 - smem_u32(smem_raw)) + (uint32_t)p.stages * stage_bytes + (uint32_t)ew * p.epi_bytes_per_warp);
    uint32_t epi_sb = 0;                                   // class-1 epilogue: which of the warp's two staging buffers is next
    uint32_t u = 0;
    uint32_t aux_n = 0;                       // class 2: aux tiles requested so far by this warp (buffer = n & 1)
    PROF_DECL
    for (WorkIter w(p); w.next(); ++u) {
      const int tile = w.tile;
      const int n_blk = tile % p.n_tiles, m_blk = tile / p.n_tiles;
      const int row_base = m_blk * TM + (int)cta_rank * TBM + q * 32;
      const int n0 = n_blk * p.block_n;
      const uint32_t acc = u % num_acc, acc_ph = (u / num_acc) & 1;
      // the two warps of a lane quarter take alternate 32-column chunks; which of them starts at chunk 0 flips every
      // tile so odd chunk counts (N = 96: 3 chunks) balance out across tiles
      const int chunk_sel = (ew >> 2) ^ (int)(u & 1);
      constexpr int kGroups = kEW / 4;          // warps per lane quarter
      if (EPI_CLASS == 2) {
        // ---- DGELU (bf16: D = (acc+bias) * aux) / RESIDUAL (fp32: D = aux + row_scale * (acc+bias)):
        //      the aux tile of each 32x32 chunk is TMA-loaded ahead of use into a per-warp ring (4 x 2 KB bf16 tiles: up to
        //      three chunks ahead, i.e. a warp's whole share of a unit is requested before it waits for the accumulator;
        //      2 x 4 KB fp32 tiles: one chunk ahead), updated in place by the thread that owns the row (TMEM row layout)
        //      and TMA-stored from the same buffer.
        const bool resid = p.epi.epilogue == SWIN_EPI_RESIDUAL;
        const uint32_t tile_bytes = resid ? 4096u : 2048u;
        const uint32_t nbuf = resid ? 2u : 4u;
        uint8_t* sbuf = reinterpret_cast<uint8_t*>(stage);
        const uint32_t sbuf_a = smem_u32(sbuf);
        const uint32_t abar0 = smem_u32(&aux_bars[ew][0]);
        float rscale = 1.0f;
        if (resid && p.epi.row_scale != nullptr) {
          int rr = row_base + lane; if (rr >= p.epi.M) rr = p.epi.M - 1;
          rscale = p.epi.row_scale[rr / p.epi.rows_per_image];
        }
        const int c0 = chunk_sel * 32;
        const int nch = c0 < p.block_n ? (p.block_n - c0 + 63) / 64 : 0;       // this warp's chunks: c0, c0 + 64, ...
        auto issue_aux = [&](int k) {
          if (lane == 0) {
            tma_store_wait_read<0>();                 // the store that last read the target buffer has drained it
            const uint32_t b = aux_n % nbuf;
            mbar_expect_tx(abar0 + 8 * b, tile_bytes);
            tma_load_2d(sbuf_a + b * tile_bytes, &tmD2, abar0 + 8 * b, n0 + c0 + 64 * k, row_base);
          }
          ++aux_n;
        };
        uint32_t cur = aux_n;
        int issued = 0;
        for (; issued < nch && issued < (int)nbuf - 1; ++issued) issue_aux(issued);
        PROF_WAIT(prof_w0, mbar_wait(tfull_bar(acc), acc_ph));
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
        uint32_t v[32];
        if (nch > 0) tmem_ld32(taddr + c0, v);
        for (int k = 0; k < nch; ++k, ++cur) {
          const int c = c0 + 64 * k;
          float4 b4[8];
          if (p.epi.bias != nullptr) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) b4[kk] = __ldg(reinterpret_cast<const float4*>(p.epi.bias + n0 + c) + kk);
          } else {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) b4[kk] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const uint32_t b = cur % nbuf;
          uint8_t* tile = sbuf + b * tile_bytes;
          tmem_ld_wait();
          mbar_wait(abar0 + 8 * b, (cur / nbuf) & 1);
          if (resid) {
            float4 ax[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) ax[cc] = *reinterpret_cast<const float4*>(tile + lane * 128 + ((cc ^ (lane & 7)) << 4));
            if (issued < nch) { issue_aux(issued); ++issued; }     // request the next aux tile of this unit
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
              float4 o;
              o.x = fmaf(rscale, __uint_as_float(v[4 * cc + 0]) + b4[cc].x, ax[cc].x);
              o.y = fmaf(rscale, __uint_as_float(v[4 * cc + 1]) + b4[cc].y, ax[cc].y);
              o.z = fmaf(rscale, __uint_as_float(v[4 * cc + 2]) + b4[cc].z, ax[cc].z);
              o.w = fmaf(rscale, __uint_as_float(v[4 * cc + 3]) + b4[cc].w, ax[cc].w);
              *reinterpret_cast<float4*>(tile + lane * 128 + ((cc ^ (lane & 7)) << 4)) = o;
            }
          } else {
            const uint32_t swz = (uint32_t)((lane >> 1) & 3);
            int4 ax[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) ax[cc] = *reinterpret_cast<const int4*>(tile + lane * 64 + ((cc ^ swz) << 4));
            if (issued < nch) { issue_aux(issued); ++issued; }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const uint32_t g[4] = {(uint32_t)ax[cc].x, (uint32_t)ax[cc].y, (uint32_t)ax[cc].z, (uint32_t)ax[cc].w};
              uint32_t o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int kx = 8 * cc + 2 * e;          // columns kx, kx+1
                const float bx = (kx & 3) == 0 ? b4[kx >> 2].x : b4[kx >> 2].z, by = (kx & 3) == 0 ? b4[kx >> 2].y : b4[kx >> 2].w;
                o[e] = pack_bf16((__uint_as_float(v[kx]) + bx) * bf16_lo(g[e]), (__uint_as_float(v[kx + 1]) + by) * bf16_hi(g[e]));
              }
              *reinterpret_cast<int4*>(tile + lane * 64 + ((cc ^ swz) << 4)) = make_int4(o[0], o[1], o[2], o[3]);
            }
          }
          if (k + 1 < nch) tmem_ld32(taddr + c + 64, v);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&tmD, sbuf_a + b * tile_bytes, n0 + c, row_base); tma_store_commit(); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTAS == 1) mbar_arrive(tempty_bar(acc)); else mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0)); }
        continue;
      }
      if (EPI_CLASS == 1) {
        // ---- STORE / GELU with bf16 outputs: math in the TMEM row layout, bf16 tile staged in the TMA 64B-swizzle
        //      layout, one TMA store per 32x32 chunk (no per-lane global stores, no address arithmetic)
        PROF_WAIT(prof_w0, mbar_wait(tfull_bar(acc), acc_ph));
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
        const bool gelu = p.epi.epilogue == SWIN_EPI_GELU;
        // two staging buffers per warp, used alternately: a chunk's TMA store needs ~1k cycles before it has read its buffer
        // (cp.async.bulk.wait_group.read), as long as the math of a chunk -- with one buffer every chunk waited for it
        const uint32_t sb_bytes = gelu ? 4096u : 2048u;
        const uint32_t swz = (uint32_t)((lane >> 1) & 3);
        uint32_t v[32];
        int c = (int)(((uint32_t)(ew >> 2) + u) % kGroups) * 32;        // rotate the starting warp every tile
        if (c < p.block_n) tmem_ld32(taddr + c, v);
        for (; c < p.block_n; c += 32 * kGroups) {
          float4 b4[8];
          if (p.epi.bias != nullptr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) b4[k] = __ldg(reinterpret_cast<const float4*>(p.epi.bias + n0 + c) + k);
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) b4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          tmem_ld_wait();
          uint32_t pu[16], ph[16];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float x0 = __uint_as_float(v[4 * k + 0]) + b4[k].x, x1 = __uint_as_float(v[4 * k + 1]) + b4[k].y;
            const float x2 = __uint_as_float(v[4 * k + 2]) + b4[k].z, x3 = __uint_as_float(v[4 * k + 3]) + b4[k].w;
            if (gelu) {                                // pu <- gelu(u) (D), ph <- gelu'(u) (D2)
              f32x2 ga, da, gb, db;
              gelu_fast2(x0, x1, &ga, &da); gelu_fast2(x2, x3, &gb, &db);
              float g0, g1, g2, g3, d0, d1, d2, d3;
              unpk2(ga, g0, g1); unpk2(gb, g2, g3); unpk2(da, d0, d1); unpk2(db, d2, d3);
              pu[2 * k] = pack_bf16(g0, g1); pu[2 * k + 1] = pack_bf16(g2, g3);
              ph[2 * k] = pack_bf16(d0, d1); ph[2 * k + 1] = pack_bf16(d2, d3);
            } else {
              pu[2 * k] = pack_bf16(x0, x1); pu[2 * k + 1] = pack_bf16(x2, x3);
            }
          }
          if (c + 32 * kGroups < p.block_n) tmem_ld32(taddr + c + 32 * kGroups, v);      // next chunk's TMEM read overlaps the stores below
          uint8_t* sbuf = reinterpret_cast<uint8_t*>(stage) + epi_sb * sb_bytes;
          const uint32_t sbuf_a = smem_u32(sbuf);
          epi_sb ^= 1u;
          if (lane == 0) tma_store_wait_read<1>();                   // the stores of the chunk before the previous one have drained this buffer
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint32_t o = (uint32_t)lane * 64u + ((cc ^ swz) << 4);
            *reinterpret_cast<int4*>(sbuf + o) = make_int4(pu[4 * cc], pu[4 * cc + 1], pu[4 * cc + 2], pu[4 * cc + 3]);
            if (gelu) *reinterpret_cast<int4*>(sbuf + 2048 + o) = make_int4(ph[4 * cc], ph[4 * cc + 1], ph[4 * cc + 2], ph[4 * cc + 3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (gelu) { tma_store_2d(&tmD, sbuf_a, n0 + c, row_base); tma_store_2d(&tmD2, sbuf_a + 2048, n0 + c, row_base); }
            else tma_store_2d(&tmD, sbuf_a, n0 + c, row_base);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTAS == 1) mbar_arrive(tempty_bar(acc)); else mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0)); }
        continue;
      }
      {
        long long drow = 0; float scale = 1.f;
        const bool live = epi_row_setup(p.epi, row_base + lane, &drow, &scale);
        epi_rowdst[ew][lane] = live ? drow : -1;
        epi_rowscale[ew][lane] = scale;
      }
      __syncwarp();
      // Residual epilogues read their aux rows (fp32 x, scattered by the window-reverse map) with plain loads right before use:
      // with 8 float4 per lane in flight the read stream is latency-bound (bytes in flight / latency).  The NEXT unit's rows
      // are therefore pulled into L2 now, a whole unit ahead: lane = one row, one prefetch per 128-byte line of this warp's chunks.
      if (p.epi.epilogue == SWIN_EPI_SCATTER_RESIDUAL || p.epi.epilogue == SWIN_EPI_RESIDUAL) {
        WorkIter wn = w;
        if (wn.next()) {
          const int nrow = (wn.tile / p.n_tiles) * TM + (int)cta_rank * TBM + q * 32 + lane;
          const int nn0 = (wn.tile % p.n_tiles) * p.block_n;
          long long drow = 0; float scale = 1.f;
          if (epi_row_setup(p.epi, nrow, &drow, &scale)) {
            const float* src = reinterpret_cast<const float*>(p.epi.aux) + drow * p.epi.ldd + nn0;
            for (int cc = ((ew >> 2) ^ (int)((u + 1) & 1)) * 32; cc < p.block_n && nn0 + cc < p.epi.N; cc += 64)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(src + cc));
          }
        }
      }
      PROF_WAIT(prof_w0, mbar_wait(tfull_bar(acc), acc_ph));
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
      const int piece = lane & 7, rsub = lane >> 3;
      uint32_t v[32];
      int c = chunk_sel * 32;
      if (c < p.block_n) tmem_ld32(taddr + c, v);
      for (; c < p.block_n; c += 64) {
        tmem_ld_wait();
#pragma unroll
        for (int pc = 0; pc < 8; ++pc)
          stage[lane * 8 + (pc ^ (lane & 7))] = make_float4(__uint_as_float(v[4 * pc]), __uint_as_float(v[4 * pc + 1]),
                                                            __uint_as_float(v[4 * pc + 2]), __uint_as_float(v[4 * pc + 3]));
        __syncwarp();
        if (c + 64 < p.block_n) tmem_ld32(taddr + c + 64, v);      // next chunk's TMEM read overlaps this chunk's stores
        const int col = n0 + c + piece * 4;
        switch (p.epi.epilogue) {
          case SWIN_EPI_STORE: epi_chunk<SWIN_EPI_STORE>(p.epi, stage, epi_rowdst[ew], epi_rowscale[ew], col, rsub, piece); break;
          case SWIN_EPI_GELU: epi_chunk<SWIN_EPI_GELU>(p.epi, stage, epi_rowdst[ew], epi_rowscale[ew], col, rsub, piece); break;
          case SWIN_EPI_RESIDUAL:
          case SWIN_EPI_SCATTER_RESIDUAL: epi_chunk<SWIN_EPI_RESIDUAL>(p.epi, stage, epi_rowdst[ew], epi_rowscale[ew], col, rsub, piece); break;
          case SWIN_EPI_DGELU: epi_chunk<SWIN_EPI_DGELU>(p.epi, stage, epi_rowdst[ew], epi_rowscale[ew], col, rsub, piece); break;
          default: epi_chunk<SWIN_EPI_ATOMIC_ADD>(p.epi, stage, epi_rowdst[ew], epi_rowscale[ew], col, rsub, piece); break;
        }
        __syncwarp();
      }
      if (do_colsum && n_blk == 0 && ew < 4) {        // one warp per lane quarter: row sums of A^T = bias gradient
        uint32_t cs[16];
        tmem_ld16(taddr + 256, cs);
        tmem_ld_wait();
        const int m = row_base + lane;
        if (m < p.epi.M) atomicAdd(p.colsum_a + m, __uint_as_float(cs[0]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CTAS == 1) mbar_arrive(tempty_bar(acc)); else mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0)); }
    }
#ifdef SWIN_GEMM_PROF
    if (ew == 0 && lane == 0) PROF_FLUSH(6 + 8 * cta_rank, 7 + 8 * cta_rank, 15);      // [6] epilogue warp 0 blocked on accumulator-full
#endif
  }
  if (EPI_CLASS != 0 && warp >= 2 && lane == 0) tma_store_wait_all<0>();
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();      // pair: neither CTA retires while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

static int pick_block_n(int N) {
  if (N <= 256 && N % 16 == 0) return N;             // any legal UMMA N in one tile (e.g. 48 for the patch-embed dW)
  const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  for (int c : cand) if (N % c == 0) return c;
  return 0;
}

// CTA-pair mode (cta_group::2, 256-row tiles).  swin_gemm_pair_mode() / SWIN_GEMM_PAIR: 0 disables it, 1 (default) applies the shape policy
// below, 2 uses it wherever it is legal (tests, A/B tools).  The B half-tile of a CTA must be whole swizzle atoms: any multiple of 8 rows when
// B is K-major, a multiple of 64 columns when it is MN-major.
static int g_pair_mode = [] { const char* e = getenv("SWIN_GEMM_PAIR"); return e ? atoi(e) : 1; }();
static int pair_mode() { return __atomic_load_n(&g_pair_mode, __ATOMIC_RELAXED); }
int set_pair_mode(int mode) {
  if (mode < 0) return pair_mode();
  return __atomic_exchange_n(&g_pair_mode, mode > 2 ? 2 : mode, __ATOMIC_RELAXED);
}
static int pick_pair_block_n(int N, bool b_mn) {
  if (!b_mn) { const int bn = pick_block_n(N); return (bn > 0 && bn % 16 == 0) ? bn : 0; }
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  return 0;
}

// Split-K factor of the weight-gradient GEMM: tiles * splits should fill the SMs (or SM pairs) in whole rounds (one round if that wastes
// < 10 % of the machine, else the better of one / two / three rounds), with at least 4 k-blocks per unit.
static int pick_splits(int tiles, int kb_total, int slots) {
  int maxs = kb_total / 4;
  if (maxs < 1) maxs = 1;
  int best = 1; double best_eff = 0.0;
  for (int rounds = 1; rounds <= 3; ++rounds) {
    int s = (rounds * slots) / tiles;
    if (s < 1) s = 1;
    if (s > maxs) s = maxs;
    const int units = tiles * s;
    const double eff = (double)units / (double)(ceil_div(units, slots) * slots);
    if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    if (best_eff >= 0.9) break;
  }
  return best;
}

// Tile family, tile width, tile grid and split-K factor of one problem: pure host logic (no CUDA call), also reachable
// through swin_gemm_plan() so that the policy is testable without a GPU.
static int choose_tiles(const swin_gemm_args* a, GemmTcParams* pp) {
  GemmTcParams& p = *pp;
  const bool b_mn = a->b_trans != 0;
  p.block_n = pick_block_n(a->N);
  SWIN_REQUIRE(p.block_n > 0, "gemm(bf16): unsupported N %d", a->N);
  p.ctas = 1;
  {
    const int mt = ceil_div(a->M, TBM), pbn = pick_pair_block_n(a->N, b_mn), mode = pair_mode();
    // policy (measured per shape, tools/pair_ab.sh -> profiles/r01/gemm_pair_ab.txt): pair tiles win or tie under every epilogue
    // except DGELU (its aux-tile ring couples the two CTAs' epilogues).  An MN-major B needs half-tiles of whole 64-column
    // atoms, so N = 384 would drop from 192- to 128-wide tiles: a wash for dX, kept on 1-CTA tiles.  The split-K weight
    // gradient pairs up only for an even tile count (a half-empty pair is pure loss when every unit runs the whole K range).
    // Elsewhere an odd tile count wastes half a pair tile once, acceptable from 8 tiles up.
    const bool dw = a->epilogue == SWIN_EPI_ATOMIC_ADD;
    const bool shape_ok = dw ? (mt % 2 == 0) : (mt % 2 == 0 || mt >= 8);
    const bool epi_ok = a->epilogue != SWIN_EPI_DGELU && (dw || !b_mn || pbn >= p.block_n);
    if (mode > 0 && pbn > 0 && mt >= 2 && (mode >= 2 || (shape_ok && epi_ok))) { p.ctas = 2; p.block_n = pbn; }
  }
  p.m_tiles = ceil_div(a->M, TBM * p.ctas);
  p.n_tiles = a->N / p.block_n;
  p.kb_total = ceil_div(a->K, TBK);
  p.K = a->K;
  p.splits = a->epilogue == SWIN_EPI_ATOMIC_ADD ? pick_splits(p.m_tiles * p.n_tiles, p.kb_total, persistent_sms() / p.ctas) : 1;
  p.kb_per_split = ceil_div(p.kb_total, p.splits);
  p.splits = ceil_div(p.kb_total, p.kb_per_split);
  return 0;
}

int gemm_tc_plan(const swin_gemm_args* a, int* out6) {
  SWIN_REQUIRE(a && out6, "gemm_plan: null pointer");
  SWIN_REQUIRE(a->dtype == SWIN_BF16, "gemm_plan: only the bf16 (tcgen05) GEMM has a tile plan");
  SWIN_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->N % 16 == 0, "gemm_plan: bad shape (N must be a multiple of 16)");
  GemmTcParams p;
  const int rc = choose_tiles(a, &p);
  if (rc) return rc;
  out6[0] = p.ctas; out6[1] = p.block_n; out6[2] = p.m_tiles; out6[3] = p.n_tiles; out6[4] = p.splits; out6[5] = p.kb_per_split;
  return 0;
}

int gemm_tc(const swin_gemm_args* a, cudaStream_t st) {
  EpiParams ep;
  int rc = make_epi_params(a, &ep);
  if (rc) return rc;
  SWIN_REQUIRE(a->A && a->B, "gemm: null operand");
  SWIN_REQUIRE(a->N % 16 == 0, "gemm(bf16): N must be a multiple of 16 (got %d)", a->N);
  SWIN_REQUIRE(a->lda % 8 == 0 && a->ldb % 8 == 0, "gemm(bf16): lda/ldb must be multiples of 8");
  if (a->M == 0) return 0;
  const bool a_mn = a->a_trans != 0, b_mn = a->b_trans != 0;
  GemmTcParams p;
  p.epi = ep;
  rc = choose_tiles(a, &p);
  if (rc) return rc;
  p.colsum_a = nullptr;
  if (a->colsum_a != nullptr) {
    SWIN_REQUIRE(a->epilogue == SWIN_EPI_ATOMIC_ADD && a_mn && b_mn, "gemm: colsum_a needs ATOMIC_ADD with a_trans = b_trans = 1");
    p.colsum_a = a->colsum_a;
  }
  p.a_bytes = TBM * TBK * 2;
  const int bn_cta = p.block_n / p.ctas;                       // B rows (columns of D) staged by one CTA
  const int bn_rows = b_mn ? ceil_div(bn_cta, 64) * 64 : bn_cta;
  p.b_bytes = (uint32_t)bn_rows * TBK * 2;
  p.stage_bytes = (p.a_bytes + p.b_bytes + 1023u) & ~1023u;
  const uint32_t stage_bytes = p.stage_bytes;
  const uint32_t budget = 187 * 1024;
  p.stages = (int)(budget / stage_bytes);
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  SWIN_REQUIRE(p.stages >= 2, "gemm(bf16): tile does not fit in shared memory");

  CUtensorMap tmA, tmB;
  if (!a_mn) rc = make_tmap_bf16_2d(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda * 2, TBK, TBM, CU_TENSOR_MAP_SWIZZLE_128B);
  else       rc = make_tmap_bf16_2d(&tmA, a->A, (uint64_t)a->M, (uint64_t)a->K, (uint64_t)a->lda * 2, 64, TBK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_bf16_2d(&tmB, a->B, (uint64_t)a->K, (uint64_t)a->N, (uint64_t)a->ldb * 2, TBK, (uint32_t)bn_cta, CU_TENSOR_MAP_SWIZZLE_128B);
  else       rc = make_tmap_bf16_2d(&tmB, a->B, (uint64_t)a->N, (uint64_t)a->K, (uint64_t)a->ldb * 2, 64, TBK, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;

  // TMA epilogues: class 1 = bf16 STORE / GELU outputs; class 2 = DGELU (bf16 aux/out) and RESIDUAL (fp32 aux/out);
  // all need the tile to be a whole number of 32-column chunks
  CUtensorMap tmD = tmA, tmD2 = tmA;
  p.tma_epi = 0;
  if (p.block_n % 32 == 0) {
    if ((a->epilogue == SWIN_EPI_STORE || a->epilogue == SWIN_EPI_GELU) && a->d_dtype == SWIN_BF16) {
      rc = make_tmap_bf16_2d(&tmD, a->D, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldd * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      if (a->epilogue == SWIN_EPI_GELU) {
        rc = make_tmap_bf16_2d(&tmD2, a->D2, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldd * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
      }
      p.tma_epi = 1;
    } else if (a->epilogue == SWIN_EPI_DGELU && a->d_dtype == SWIN_BF16) {
      rc = make_tmap_bf16_2d(&tmD, a->D, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldd * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      rc = make_tmap_bf16_2d(&tmD2, a->aux, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldd * 2, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      p.tma_epi = 2;
    } else if (a->epilogue == SWIN_EPI_RESIDUAL && (212u * 1024 - 8192u * kEpiWarps - 1024) / stage_bytes >= 4) {
      rc = make_tmap_f32_2d(&tmD, a->D, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldd * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      rc = make_tmap_f32_2d(&tmD2, a->aux, (uint64_t)a->N, (uint64_t)a->M, (uint64_t)a->ldd * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
      p.tma_epi = 2;
    }
  }
  // class 2: 2 x 4 KB (fp32) or 4 x 2 KB (bf16) aux ring per warp; class 1: two staging buffers of 2 KB (STORE) or 2 x 2 KB (GELU: D and D2)
  p.epi_bytes_per_warp = (p.tma_epi == 2 || (p.tma_epi == 1 && a->epilogue == SWIN_EPI_GELU)) ? 8192u : 4096u;
  const uint32_t epi_bytes = p.epi_bytes_per_warp * (uint32_t)epi_warps(p.tma_epi);
  {
    const uint32_t ring_budget = 212 * 1024 - epi_bytes - 1024;     // dynamic smem: ring + epilogue staging + alignment slack
    int st2 = (int)(ring_budget / stage_bytes);
    if (st2 < p.stages) p.stages = st2;
  }
  const size_t smem = (size_t)p.stages * stage_bytes + epi_bytes + 1024;
  const int total_units = p.m_tiles * p.n_tiles * p.splits;
  const int slots = persistent_sms() / p.ctas;                          // CTAs, or CTA pairs (one per TPC)
  const int grid = (total_units < slots ? total_units : slots) * p.ctas;
#define LAUNCH_TC(AM, BM, TE, CT)                                                                                 \
  do {                                                                                                            \
    { const int ar = ensure_dyn_smem((const void*)gemm_tc_kernel<AM, BM, TE, CT>, 212 * 1024); if (ar) return ar; }  \
    cudaLaunchConfig_t cfg = {};                                                                                  \
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)gemm_threads(TE));                          \
    cfg.dynamicSmemBytes = smem; cfg.stream = st;                                                                 \
    cudaLaunchAttribute lattr[2];                                                                                 \
    int nattr = 0;                                                                                                \
    if (CT > 1) {                                                                                                 \
      lattr[0].id = cudaLaunchAttributeClusterDimension;                                                          \
      lattr[0].val.clusterDim.x = CT; lattr[0].val.clusterDim.y = 1; lattr[0].val.clusterDim.z = 1;               \
      nattr = 1;                                                                                                  \
    }                                                                                                             \
    nattr += pdl_attr(&lattr[nattr]);                                                                             \
    cfg.attrs = lattr; cfg.numAttrs = nattr;                                                                      \
    cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<AM, BM, TE, CT>, tmA, tmB, tmD, tmD2, p);            \
    if (le != cudaSuccess) { set_error("gemm_tc launch: %s", cudaGetErrorString(le)); return (int)le; }           \
  } while (0)
#define LAUNCH_TC1(AM, BM, TE)                                                                                    \
  do { if (p.ctas == 2) LAUNCH_TC(AM, BM, TE, 2); else LAUNCH_TC(AM, BM, TE, 1); } while (0)
#define LAUNCH_TC2(AM, BM)                                                                                        \
  do {                                                                                                            \
    if (p.tma_epi == 2) LAUNCH_TC1(AM, BM, 2); else if (p.tma_epi == 1) LAUNCH_TC1(AM, BM, 1); else LAUNCH_TC1(AM, BM, 0); \
  } while (0)
  if (!a_mn && !b_mn) LAUNCH_TC2(false, false);
  else if (!a_mn && b_mn) LAUNCH_TC2(false, true);
  else if (a_mn && b_mn) LAUNCH_TC2(true, true);
  else LAUNCH_TC2(true, false);
#undef LAUNCH_TC2
#undef LAUNCH_TC1
#undef LAUNCH_TC
  SWIN_LAUNCH_CHECK();
  return 0;
}

#ifdef SWIN_GEMM_PROF
}  // namespace swin
extern "C" int swin_debug_gemm_prof(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, swin::g_gemm_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(swin::g_gemm_prof, z, sizeof(z)); }
  return 0;
}
namespace swin {
#endif
}  // namespace swin
