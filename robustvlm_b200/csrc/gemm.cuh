// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] * B[N,K]^T     (16-bit operands, fp32 accumulation in TMEM)
//
// This is the contraction engine under every linear layer of the CLIP ViT
// forward and input-gradient backward (open_clip VisionTransformer, SURVEY.md
// Appendix A; reference call sites train/adversarial_training_clip.py:253-257,
// train/pgd_train.py:32,38).  The reference runs these as cuBLAS SGEMMs through
// torch.nn.Linear; here they are TMA -> shared memory (128B swizzle) ->
// tcgen05.mma -> TMEM -> fused epilogue.
//
// Roles (GEMM_THREADS = 128 + 32 x epilogue warps, 1 CTA / SM, persistent over output tiles):
//   warp 0 lane 0 : TMA producer   (kStages-deep smem ring, full/empty mbarriers)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma, double-buffered TMEM accumulator)
//   warp 2        : TMEM allocator
//   warps 4..11   : epilogue; GEMM_EPI_WARPS / 4 warps per TMEM lane quarter, each takes one
//                   slice of the accumulator columns.  Two paths:
//                   TMA (16-bit outputs): tcgen05.ld -> registers (lane = row) -> Epi math ->
//                     64B-swizzled per-warp staging tile -> cp.async.bulk.tensor store per
//                     output and 32-column chunk (an input tile arrives by TMA load);
//                   transposing (fp32 outputs, patch epilogues): tcgen05.ld -> registers ->
//                     swizzled smem transpose -> Epi functor on row-contiguous float4 pieces
//
// Tile: BLOCK_M = 128 (UMMA M, one TMEM lane per row), BLOCK_N = BN (UMMA N),
// BLOCK_K = 64 elements (= one 128-byte swizzle row).
//
// Operand majors: the default is K-major for both (activations [M,K] row-major,
// weights [N,K] row-major).  B_MN / A_MN select MN-major operands (B given as
// [K,N] row-major, A given as [K,M] row-major) so transposed products (wgrad,
// attention backward) need no data transposition.
#pragma once
#include "ptx.cuh"

namespace rvlm {

// Epilogue geometry (build-time; A/B builds: make OUTDIR=../_lib/var_x EXTRA="-DRVLM_EPI_WARPS=16
// -DRVLM_EPI_GC=16", tools/ab_variants.sh).  Measured in-step on B200 (config 2, ms per step of
// the K = 1024 GEMMs gelu / gelu' / out-proj): 8 warps x 32-column granules 127.9 / 112.9 / 38.6;
// 8 x 16: 135.9 / 135.1 / 44.0; 16 x 16: 128.0 / 136.4 / 46.6 -- the width of the row pieces a
// warp-wide access covers matters more than the number of warps, so the default stays 8 x 32.
#ifndef RVLM_EPI_WARPS
#define RVLM_EPI_WARPS 8
#endif
#ifndef RVLM_EPI_GC
#define RVLM_EPI_GC 32
#endif
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_EPI_WARPS = RVLM_EPI_WARPS;        // 8 or 16
constexpr int GEMM_EPI_GC = RVLM_EPI_GC;              // granule: 16 or 32 accumulator columns
constexpr int GEMM_THREADS = 128 + GEMM_EPI_WARPS * 32;
static_assert(GEMM_EPI_WARPS == 8 || GEMM_EPI_WARPS == 16, "epilogue warps: 8 or 16");
static_assert(GEMM_EPI_GC == 16 || GEMM_EPI_GC == 32, "granule: 16 or 32 columns");

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierBytes = 512;   // pipeline barriers + one mbarrier per epilogue warp
  // per-warp staging: transpose tile of the classic epilogue = two 2 KB TMA tiles of the TMA one
  static constexpr int kStagingBytes = GEMM_EPI_WARPS * 32 * GEMM_EPI_GC * 4;
  static constexpr int kTotal =
      kStages * kStageBytes + kStagingBytes + kBarrierBytes + 1024 /*align slack*/;
};

// The epilogue functor interface:
//   struct Aux;                                     // what the functor reads from HBM
//   struct Col;                                     // per-column operand (bias), once per chunk
//   __device__ Aux  load(int row, int col) const;   // issued early (latency hidden)
//   __device__ Col  col(int col) const;
//   __device__ void operator()(int row, int col, const float4& acc, const Aux&, const Col&) const
// called once per (row, 4 consecutive columns) with row < M and col < N already
// guaranteed (N is required to be a multiple of 32 by the host wrappers).


// pieces per batch of an epilogue functor: its kPiecesPerBatch member if it has one, else "all"
template <class E>
constexpr auto epi_batch(int) -> decltype(E::kPiecesPerBatch) { return E::kPiecesPerBatch; }
template <class E>
constexpr int epi_batch(long) { return 1 << 20; }

__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16(taddr, r); }

// Epilogue of one accumulator tile for one epilogue warp: drains this warp's 32 TMEM lanes x
// (BN / slices) columns in granules of GC columns (slices = epilogue warps per lane quarter).
// Per granule: tcgen05.ld -> registers -> XOR-swizzled warp-private smem transpose -> each lane
// owns float4 pieces (4 consecutive columns of one row; GC / 4 lanes per row, so a warp-wide
// access covers whole sectors of 32 / (GC / 4) rows) -> Epi functor.  Latency is hidden inside the
// warp as far as it goes: the HBM operands of the epilogue and the per-column operand (bias) are
// requested first, the NEXT granule's tcgen05.ld is issued as soon as the registers are staged,
// and the transposed pieces are read back in one batch before any math / store is issued.
template <int BN, bool GUARD, class Epi>
__device__ __forceinline__ void epilogue_drain_impl(const Epi& epi, float* stg, uint32_t taddr,
                                                    int row0, int n0, int slice, int M, int N,
                                                    int lane) {
  constexpr int GC = GEMM_EPI_GC;
  constexpr int SLICES = GEMM_EPI_WARPS / 4;
  constexpr int NG = BN / SLICES / GC;       // granules per warp
  constexpr int LPR = GC / 4;                // lanes per row after the transpose (4 or 8)
  constexpr int RPP = 32 / LPR;              // rows per pass (8 or 4)
  constexpr int NPASS = LPR;                 // passes over the 32 rows
  static_assert(NG >= 1, "tile too narrow for this epilogue geometry");
  const int rl = lane / LPR, c4 = lane % LPR;
  // 16-byte groups XOR-swizzled by row so that both the row-wise writes and the piece-wise reads
  // are bank-conflict free (row pitch 128 B: row & 7; row pitch 64 B: (row >> 1) & 3)
  auto swz = [](int row) { return GC == 32 ? (row & 7) : ((row >> 1) & 3); };
  const int cbase = slice * NG * GC;
  const int rem = N - (n0 + cbase);          // columns left for this slice (N % 32 == 0)
  const int ng = rem <= 0 ? 0 : (rem >= NG * GC ? NG : rem / GC);   // warp-uniform
  if (ng == 0) return;
  uint32_t r[GC];
  auto ld = [&](int g) { tmem_ld_cols(taddr + cbase + g * GC, r); };
  ld(0);
  // pieces handled per batch: all of a granule's, unless the functor asks for fewer (a functor
  // with many live registers per piece, Epi::kPiecesPerBatch)
  constexpr int NB = epi_batch<Epi>(0) < NPASS ? epi_batch<Epi>(0) : NPASS;
  static_assert(NPASS % NB == 0, "pieces per batch must divide the passes");
#pragma unroll 1
  for (int g = 0; g < ng; ++g) {
    const int col = n0 + cbase + g * GC + 4 * c4;
    typename Epi::Aux aux[NB];
    auto load_aux = [&](int b0) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int row = row0 + rl + RPP * (b0 + i);
        if (!GUARD || row < M) aux[i] = epi.load(row, col);
      }
    };
    load_aux(0);
    const typename Epi::Col ca = epi.col(col);
    tmem_ld_wait();
    // transpose: lane == row writes its GC columns
#pragma unroll
    for (int j = 0; j < LPR; ++j)
      *reinterpret_cast<uint4*>(stg + lane * GC + ((j ^ swz(lane)) << 2)) =
          make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    __syncwarp();
    if (g + 1 < ng) ld(g + 1);
#pragma unroll
    for (int b0 = 0; b0 < NPASS; b0 += NB) {
      if (b0 > 0) load_aux(b0);
      float4 v[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int rr = rl + RPP * (b0 + i);
        v[i] = *reinterpret_cast<const float4*>(stg + rr * GC + ((c4 ^ swz(rr)) << 2));
      }
      if (b0 + NB >= NPASS) __syncwarp();   // staging tile free for the next granule
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int row = row0 + rl + RPP * (b0 + i);
        if (!GUARD || row < M) epi(row, col, v[i], aux[i], ca);
      }
    }
  }
}

// does the functor offer the TMA-store path (tma_on / tma_chunk / prepare)?
template <class E>
constexpr auto epi_has_tma(int) -> decltype(&E::tma_on, true) { return true; }
template <class E>
constexpr bool epi_has_tma(long) { return false; }
template <class E>
auto epi_prepare(E& e, int M, int N, int) -> decltype(e.prepare(M, N), void()) { e.prepare(M, N); }
template <class E>
void epi_prepare(E&, int, int, long) {}

template <class E>
__device__ __forceinline__ auto epi_tma_prefetch(const E& e, int row0, int lane, int col,
                                                 uint8_t* stg, TmaCtx& c, int)
    -> decltype(e.tma_prefetch(row0, lane, col, stg, c), void()) {
  e.tma_prefetch(row0, lane, col, stg, c);
}
template <class E>
__device__ __forceinline__ void epi_tma_prefetch(const E&, int, int, int, uint8_t*, TmaCtx&, long) {}

// TMA epilogue of one accumulator tile for one epilogue warp: lane = accumulator row, the
// warp's columns in 32-column chunks; the functor converts a chunk in registers, writes its
// row into the warp's staging tile and lane 0 hands it to the TMA (Epi::tma_chunk).  The next
// chunk's tcgen05.ld is in flight while the current one is processed; a functor with an input
// tile has chunk 0's requested before the wait on the accumulator (tma_prefetch).
template <int BN>
struct TmaSlice {
  static constexpr int SLICES = GEMM_EPI_WARPS / 4;
  static constexpr int NCH = BN / SLICES / 32;
  int cbase, nch;
  __device__ __forceinline__ TmaSlice(int n0, int slice, int N) {
    cbase = slice * NCH * 32;
    const int rem = N - (n0 + cbase);          // columns left for this slice (N % 32 == 0)
    nch = rem <= 0 ? 0 : (rem >= NCH * 32 ? NCH : rem / 32);   // warp-uniform
  }
};
template <int BN, class Epi>
__device__ __forceinline__ void epilogue_drain_tma(const Epi& epi, uint8_t* stg, uint32_t taddr,
                                                   int row0, int n0, int slice, int M, int N,
                                                   int lane, TmaCtx& ctx) {
  const TmaSlice<BN> sl(n0, slice, N);
  constexpr int NCH = TmaSlice<BN>::NCH;
  if (sl.nch == 0 || row0 >= M) return;
  uint32_t r[2][32];
  tmem_ld32(taddr + sl.cbase, r[0]);
#pragma unroll
  for (int cc = 0; cc < NCH; ++cc) {
    if (cc < sl.nch) {
      tmem_ld_wait();
      const bool more = cc + 1 < NCH && cc + 1 < sl.nch;
      if (more) tmem_ld32(taddr + sl.cbase + (cc + 1) * 32, r[(cc + 1) & 1]);
      ctx.next_col = more ? n0 + sl.cbase + (cc + 1) * 32 : -1;
      epi.tma_chunk(row0, lane, n0 + sl.cbase + cc * 32, r[cc & 1], stg, ctx);
    }
  }
}

// The row guard is hoisted to the warp's 32-row slab: full slabs (all but the last M tile)
// run branch-free, so the pieces' dependent chains (MUFU, conversions, stores) interleave
// instead of executing one piece at a time.
template <int BN, class Epi>
__device__ __forceinline__ void epilogue_drain(const Epi& epi, float* stg, uint32_t taddr,
                                               int row0, int n0, int slice, int M, int N,
                                               int lane, uint64_t* tfull, uint32_t tphase,
                                               TmaCtx& ctx) {
  if constexpr (epi_has_tma<Epi>(0) && GEMM_EPI_GC == 32) {
    if (epi.tma_on()) {   // (grid-uniform)
      uint8_t* st8 = reinterpret_cast<uint8_t*>(stg);
      const TmaSlice<BN> sl(n0, slice, N);
      if (sl.nch > 0 && row0 < M) epi_tma_prefetch(epi, row0, lane, n0 + sl.cbase, st8, ctx, 0);
      mbar_wait(tfull, tphase, 4);
      tc_fence_after();
      epilogue_drain_tma<BN>(epi, st8, taddr, row0, n0, slice, M, N, lane, ctx);
      return;
    }
  }
  mbar_wait(tfull, tphase, 4);
  tc_fence_after();
  if (row0 + 32 <= M) epilogue_drain_impl<BN, false>(epi, stg, taddr, row0, n0, slice, M, N, lane);
  else if (row0 < M) epilogue_drain_impl<BN, true>(epi, stg, taddr, row0, n0, slice, M, N, lane);
}
// before an epilogue warp leaves the kernel: its outstanding TMA stores are complete
template <class Epi>
__device__ __forceinline__ void epilogue_finish(const Epi& epi, int lane) {
  if constexpr (epi_has_tma<Epi>(0)) {
    if (epi.tma_on() && lane == 0) tma_store_wait0();
    __syncwarp();
  }
}

template <int BN, bool A_MN, bool B_MN, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              int M, int N, int K, int fmt, const __grid_constant__ Epi epi) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array, not an integer round trip: keeps the
  // address space visible to the compiler -> LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  float* staging = reinterpret_cast<float*>(smem + kStages * S::kStageBytes);
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes + S::kStagingBytes);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + GEMM_BM - 1) / GEMM_BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + GEMM_BK - 1) / GEMM_BK;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], GEMM_EPI_WARPS * 32);
    }
    for (int i = 0; i < GEMM_EPI_WARPS; ++i) mbar_init(&bars[32 + i], 1);   // epilogue input tiles
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
        uint8_t* sa = smem_a + stage * S::kABytes;
        uint8_t* sb = smem_b + stage * S::kBBytes;
        if constexpr (!A_MN) {
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
        } else {
#pragma unroll
          for (int j = 0; j < GEMM_BM / 64; ++j)
            tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m_blk * GEMM_BM + j * 64,
                        kb * GEMM_BK);
        }
        if constexpr (!B_MN) {
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n_blk * BN + j * 64, kb * GEMM_BK);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_f16((uint32_t)fmt, GEMM_BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase, 3);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * S::kABytes);
        const uint32_t b_addr = smem_u32(smem_b + stage * S::kBBytes);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          // K-major: advance 16 elements = 32 B inside the swizzle row.
          // MN-major: advance 16 K-rows = 2 x 1024 B groups.
          const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                      : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                      : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_f16(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tfull_bar[acc]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int slice = (warp - 4) >> 2;  // which slice of the accumulator columns
    float* stg = staging + (warp - 4) * 32 * GEMM_EPI_GC;
    TmaCtx tctx{&bars[32 + (warp - 4)], 0u, -1};
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row0 = m_blk * GEMM_BM + q * 32;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      epilogue_drain<BN>(epi, stg, taddr, row0, n_blk * BN, slice, M, N, lane, &tfull_bar[acc],
                         acc_phase, tctx);
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
    epilogue_finish(epi, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// 2-CTA variant: a cluster of two CTAs (one SM pair) computes a 256 x 256 output tile with
// tcgen05.mma.cta_group::2.  Each CTA stages its own 128 rows of A and HALF of the B tile
// (128 of the 256 N-rows), so per 128x256 of output an SM pulls 32 KB per k-block from L2
// instead of 48 KB -- the L2 -> SM fill rate is what bounds the 1-CTA kernel -- and the
// ring is 6 stages deep.  The leader CTA (cluster rank 0) issues every MMA; both CTAs run
// TMA producers (completing on the LEADER's full barrier) and epilogues (each drains its own
// 128 TMEM lanes).  K-major operands only.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (count 1) on the mbarrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// TMA load whose completion is signalled on the LEADER CTA's barrier (bit 24 of a
// shared::cluster address selects the CTA of the pair)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m,
                                                 uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit: arrive on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

// SPLIT: both operands are given as fp16 hi + lo pairs (x = hi + lo to ~22 bits) and every
// product is formed as hi*hi + lo*hi + hi*lo, three MMAs into the same TMEM accumulator -- the
// fp32-class forward FARE's cancellation site needs (the dropped lo*lo
// term is ~2^-22 relative).  A stage then holds four 16 KB tiles and the ring is 3 deep; per
// stage the tensor pipe has three times the work, so the fill latency is covered as before.
template <bool SPLIT>
struct Gemm2Smem {
  static constexpr int kStages = SPLIT ? 3 : 6;
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;   // 128 x 64
  static constexpr int kBBytes = 128 * GEMM_BK * 2;       // half of the 256-row B tile
  static constexpr int kStageBytes = (SPLIT ? 2 : 1) * (kABytes + kBBytes);
  static constexpr int kStagingBytes = GEMM_EPI_WARPS * 32 * GEMM_EPI_GC * 4;
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + 512 + 1024;
};

template <class Epi, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmAlo,
                   const __grid_constant__ CUtensorMap tmBlo, int M, int N, int K, int fmt,
                   const __grid_constant__ Epi epi) {
  using S = Gemm2Smem<SPLIT>;
  constexpr int kStages = S::kStages;
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array, not an integer round trip: keeps the
  // address space visible to the compiler -> LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage layout: [A hi | B hi] (+ [A lo | B lo] when SPLIT), 16 KB each
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + S::kABytes;
  float* staging = reinterpret_cast<float*>(smem + kStages * S::kStageBytes);
  uint64_t* bars =
      reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes + S::kStagingBytes);
  uint64_t* full_bar = bars;                      // [kStages]  (used in the leader only)
  uint64_t* empty_bar = bars + kStages;           // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;       // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]        (used in the leader only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_m = (M + 255) / 256;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + GEMM_BK - 1) / GEMM_BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (SPLIT) {
      tma_prefetch_desc(&tmAlo);
      tma_prefetch_desc(&tmBlo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);    // the leader's arrive.expect_tx (+ both CTAs' TMA bytes)
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS * 32);
    }
    for (int i = 0; i < GEMM_EPI_WARPS; ++i) mbar_init(&bars[32 + i], 1);   // epilogue input tiles
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1, 1);
        // Both producers are released by the same multicast commit (their empty barriers), so
        // the peer needs no arrival of its own on the leader's full barrier: its TMA bytes
        // simply count against the leader's expect_tx (a transiently negative tx-count is
        // legal while the leader's arrival is still pending).  No cross-CTA hop in the loop.
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
        uint8_t* sa = smem_a + stage * S::kStageBytes;
        uint8_t* sb = smem_b + stage * S::kStageBytes;
        tma_load_2d_pair(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * 256 + (int)rank * 128);
        tma_load_2d_pair(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN + (int)rank * 128);
        if (SPLIT) {
          tma_load_2d_pair(sa + 32768, &tmAlo, &full_bar[stage], kb * GEMM_BK,
                           m_blk * 256 + (int)rank * 128);
          tma_load_2d_pair(sb + 32768, &tmBlo, &full_bar[stage], kb * GEMM_BK,
                           n_blk * BN + (int)rank * 128);
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader only) =====================
    const uint32_t idesc = make_idesc_f16((uint32_t)fmt, 256, BN, 0u, 0u);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase, 3);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + stage * S::kStageBytes);
        const uint32_t b_addr = smem_u32(smem_b + stage * S::kStageBytes);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) {
          const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_f16_2cta(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          if (SPLIT) {
            umma_f16_2cta(tmem_d, make_smem_desc_sw128(a_addr + 32768 + k * 32, 16, 1024), db,
                          idesc, 1u);
            umma_f16_2cta(tmem_d, da, make_smem_desc_sw128(b_addr + 32768 + k * 32, 16, 1024),
                          idesc, 1u);
          }
        }
        umma_commit_2cta(&empty_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit_2cta(&tfull_bar[acc]);
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 TMEM lanes) =====================
    const int q = warp & 3;
    const int slice = (warp - 4) >> 2;
    float* stg = staging + (warp - 4) * 32 * GEMM_EPI_GC;
    TmaCtx tctx{&bars[32 + (warp - 4)], 0u, -1};
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row0 = m_blk * 256 + (int)rank * 128 + q * 32;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      epilogue_drain<BN>(epi, stg, taddr, row0, n_blk * BN, slice, M, N, lane, &tfull_bar[acc],
                         acc_phase, tctx);
      tc_fence_before();
      if (leader) mbar_arrive(&tempty_bar[acc]);
      else mbar_arrive_cluster(&tempty_bar[acc], 0);
    }
    epilogue_finish(epi, lane);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(512u)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------
// Host side: tensor-map construction + launch
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled();

// 2-D tensor map over a row-major [rows, cols] 16-bit matrix with row pitch
// `ld` elements; box = [box_rows, 64 cols], 128B swizzle, zero OOB fill.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols = 64);
// 3-D variant: [d2, d1, d0] with strides (s2, s1, 1) elements, box [1, box1, box0].
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t s1, uint64_t s2, uint32_t box0, uint32_t box1);

int sm_count();

// launch accounting (bench.py's gpu_launches) and optional per-launch CUDA-event timing
// of the GEMM kernel (bench.py's roofline.achieved)
void count_launch(int n = 1);
struct GemmProfScope {
  cudaStream_t st;
  int slot;
  // flops: algorithmic 2MNK of the product; mma_factor: tensor-pipe passes per product (3 for
  // split operands) -- reported separately so the roofline can state both
  // tag: the kernel's name in the per-kernel breakdown (static string); is_gemm: counts towards
  // the tcgen05 GEMM totals of rvlm_profile_gemm_collect; n, k: the GEMM shape (breakdown key)
  GemmProfScope(cudaStream_t s, double flops, double mma_factor = 1.0, const char* tag = "gemm",
                bool is_gemm = true, int n = 0, int k = 0);
  ~GemmProfScope();
};
// a non-GEMM kernel launch in the same per-kernel breakdown (attention, LayerNorm ...)
struct KernelProfScope : GemmProfScope {
  KernelProfScope(cudaStream_t s, const char* tag, double flops = 0.0)
      : GemmProfScope(s, flops, 1.0, tag, false, 0, 0) {}
};
// tag of an epilogue functor: its kTag member if it has one
template <class E>
constexpr auto epi_tag(int) -> decltype(E::kTag) { return E::kTag; }
template <class E>
constexpr const char* epi_tag(long) { return "gemm"; }

// "Has this been done on the current device?" for the per-function cudaFuncSetAttribute
// opt-ins: the attribute is per device, so the cache is too (a process may drive several GPUs).
struct DeviceOnce {
  unsigned long long mask[2] = {0ull, 0ull};
  static int dev() {
    int d = 0;
    cudaGetDevice(&d);
    return d & 127;
  }
  bool done() const {
    const int d = dev();
    return (__atomic_load_n(&mask[d >> 6], __ATOMIC_ACQUIRE) >> (d & 63)) & 1ull;
  }
  void mark() {
    const int d = dev();
    __atomic_fetch_or(&mask[d >> 6], 1ull << (d & 63), __ATOMIC_RELEASE);
  }
};

// the same for opt-ins whose size varies with the call: remembers the largest value set per device
struct DeviceMax {
  size_t v[128] = {};
  bool need(size_t bytes) {
    const int d = DeviceOnce::dev();
    if (bytes <= v[d]) return false;
    v[d] = bytes;
    return true;
  }
};

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int fmt,
                  const Epi& epi, cudaStream_t stream, int max_ctas = 0) {
  if (N % 32 != 0 || M <= 0 || K <= 0) return -22;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap_2d(&tmA, A, M, K, lda, GEMM_BM);
  else       rc = make_tmap_2d(&tmA, A, K, M, lda, GEMM_BK);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_2d(&tmB, B, N, K, ldb, BN);
  else       rc = make_tmap_2d(&tmB, B, K, N, ldb, GEMM_BK);
  if (rc) return rc;
  auto kern = gemm16_kernel<BN, A_MN, B_MN, Epi>;
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GemmSmem<BN>::kTotal);
    if (e != cudaSuccess) return -(int)e - 1000;
    attr_set.mark();
  }
  const int num_tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + BN - 1) / BN);
  int grid = sm_count();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (num_tiles < grid) grid = num_tiles;
  cudaError_t e;
  {
    Epi epi_k = epi;
    epi_prepare(epi_k, M, N, 0);   // tensor maps of a TMA-store epilogue
    GemmProfScope prof(stream, 2.0 * (double)M * (double)N * (double)K, 1.0, epi_tag<Epi>(0), true,
                       N, K);
    kern<<<grid, GEMM_THREADS, GemmSmem<BN>::kTotal, stream>>>(tmA, tmB, M, N, K, fmt, epi_k);
    e = cudaGetLastError();
  }
  count_launch();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

bool gemm_use_2cta();

// Alo / Blo non-null: split operands (fp16 hi + lo), three MMAs per product.
template <class Epi, bool SPLIT = false>
int launch_gemm16_2cta(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int fmt,
                       const Epi& epi, cudaStream_t stream, const void* Alo = nullptr,
                       const void* Blo = nullptr) {
  if (N % 32 != 0 || M <= 0 || K <= 0) return -22;
  if (SPLIT && (!Alo || !Blo || fmt != 0)) return -22;
  CUtensorMap tmA, tmB, tmAlo, tmBlo;
  int rc = make_tmap_2d(&tmA, A, M, K, lda, GEMM_BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, B, N, K, ldb, 128);
  if (rc) return rc;
  if (SPLIT) {
    rc = make_tmap_2d(&tmAlo, Alo, M, K, lda, GEMM_BM);
    if (rc) return rc;
    rc = make_tmap_2d(&tmBlo, Blo, N, K, ldb, 128);
    if (rc) return rc;
  } else {
    tmAlo = tmA;
    tmBlo = tmB;
  }
  auto kern = gemm16_2cta_kernel<Epi, SPLIT>;
  using S = Gemm2Smem<SPLIT>;
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         S::kTotal);
    if (e != cudaSuccess) return -(int)e - 1000;
    attr_set.mark();
  }
  const int num_tiles = ((M + 255) / 256) * ((N + 255) / 256);
  int grid = sm_count() & ~1;
  if (2 * num_tiles < grid) grid = 2 * num_tiles;
  cudaError_t e;
  {
    Epi epi_k = epi;
    epi_prepare(epi_k, M, N, 0);   // tensor maps of a TMA-store epilogue
    GemmProfScope prof(stream, 2.0 * (double)M * (double)N * (double)K, SPLIT ? 3.0 : 1.0,
                       epi_tag<Epi>(0), true, N, K);
    kern<<<grid, GEMM_THREADS, S::kTotal, stream>>>(tmA, tmB, tmAlo, tmBlo, M, N, K, fmt, epi_k);
    e = cudaGetLastError();
  }
  count_launch();
  return e == cudaSuccess ? 0 : -(int)e - 1000;
}

// K-major GEMM used by the engine: the 2-CTA kernel unless disabled (RVLM_GEMM_2CTA=0)
template <class Epi>
int launch_gemm_kmajor(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                       int fmt, const Epi& epi, cudaStream_t stream) {
  if (gemm_use_2cta()) return launch_gemm16_2cta(A, lda, B, ldb, M, N, K, fmt, epi, stream);
  return launch_gemm16<256, false, false>(A, lda, B, ldb, M, N, K, fmt, epi, stream);
}
// Forward GEMM of the engine: split operands when the lo halves are given (2-CTA kernel only)
template <class Epi>
int launch_gemm_fwd(const void* A, const void* Alo, int lda, const void* B, const void* Blo,
                    int ldb, int M, int N, int K, int fmt, const Epi& epi, cudaStream_t stream) {
  if (Alo && Blo)
    return launch_gemm16_2cta<Epi, true>(A, lda, B, ldb, M, N, K, fmt, epi, stream, Alo, Blo);
  return launch_gemm_kmajor(A, lda, B, ldb, M, N, K, fmt, epi, stream);
}

}  // namespace rvlm
