// conv_rs2.cu -- CTA-pair (tcgen05 cta_group::2) variant of the row-shared 3x3 kernel (conv_rs.cu) for the layers
// with 64 / 128 output channels.
//
// Why: with N = 64 or 128 accumulator columns the single-CTA MMA is bound by SHARED-MEMORY bandwidth, not by the
// tensor pipe (ncu, conv_rs<64>: l1tex data pipe 76 % busy, 1728 tensor-core wavefronts per tile = 36 MMAs x
// (32 for the 128x16 A slice + 16 for the 64x16 B slice) against 32 math cycles per MMA).  A CTA pair runs ONE
// M = 256 MMA over two output tiles: each CTA supplies its own 128 A rows (its slab) and only HALF of the B tile
// (N/2 output channels), so per CTA and MMA the tensor core reads 32 + 8 (N = 64) or 32 + 16 (N = 128) wavefronts,
// and each CTA loads / stores only half of every weight tile.
//
// Pair protocol (rank 0 = leader):
//   * both CTAs run the same tile sequence; rank r owns output tile row 2*typ + r of pair-tile typ;
//   * TMA loads land in the executing CTA's shared memory but count their bytes on the LEADER's full barrier
//     (cp.async.bulk.tensor ... cta_group::2, barrier address mapped with mapa); the leader expects both halves;
//   * the leader's elected lane issues tcgen05.mma.cta_group::2; tcgen05.commit ... multicast::cluster arrives on the
//     empty / accumulator-full barriers of BOTH CTAs;
//   * each CTA's epilogue drains its own TMEM half and hands the accumulator stage back with a cluster-scope
//     arrive on the leader's barrier (count = both CTAs' epilogue warps).
// Warps: 0 = TMA producer (slabs, weights), 1 = MMA issuer + TMEM owner, 2..9 = epilogue (two per TMEM lane quarter),
// 10 = ReLU-gate TMA ring, present only in gated data-gradient launches (kR2ThreadsGate threads, see below).
#include "common.cuh"
#include "conv_common.cuh"
#include "dreamb200.h"

#include <stdlib.h>

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int device_sm_count();

constexpr int kR2Tw = 8, kR2Th = 16;
constexpr int kR2EpiSplit = 2;
constexpr int kR2Threads = 64 + 128 * kR2EpiSplit;
constexpr int kR2ThreadsGate = kR2Threads + 32;                          // + the gate TMA warp (gated launches only)
constexpr int kR2GateSlots = 4;                                          // 64-column gate chunks in flight per CTA (max)
constexpr int kR2Rows = kR2Th + 2;                                       // image rows per slab
constexpr int kR2Pitch = 1280;                                           // 10 pixels x 128 B
constexpr int kR2SlabTx = kR2Rows * kR2Pitch;
constexpr int kR2SlabBytes = ((kR2SlabTx + 1023) / 1024) * 1024;

struct Rs2Extra {
  int sa, sb;        // ring depths: activation slabs, weight half-tiles (streamed mode)
  int gate_slots;    // > 0: the ReLU gate is staged through shared memory by warp 10 (kR2ThreadsGate threads)
};

__device__ __forceinline__ uint64_t umma_desc_k_sw128_sbo2(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int BLOCK_N, bool RESIDENT, bool PLAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kR2ThreadsGate, 1)
conv_rs2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
                const __grid_constant__ CUtensorMap tmG, const __grid_constant__ ConvParams p,
                const __grid_constant__ Rs2Extra x) {
  constexpr int kBHalf = (BLOCK_N / 2) * 128;                            // this CTA's half of a weight tile
  constexpr int kTmemCols = 2 * BLOCK_N;                                 // 128 or 256
  constexpr uint32_t kIdesc = umma_idesc_f16_m256(BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int sa = x.sa, sb = x.sb;
  const int n_wtiles = 9 * p.kchunks;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_a + sa * kR2SlabBytes;
  const uint32_t smem_out = smem_b + (RESIDENT ? n_wtiles : 3 * sb) * kBHalf;      // sb counts GROUPS of 3 tiles
  const uint32_t smem_pool = smem_out + 2 * kStageOutBytes;
  const int gate_slots = PLAIN ? 0 : x.gate_slots;
  const uint32_t smem_gate = smem_pool + (p.pool ? 2 * kPoolBytes : 0);
  const uint32_t bar_base = smem_gate + (uint32_t)gate_slots * kGateSlotBytes;
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (sa + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * sa + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * sa + sb + s); };
  const uint32_t misc = bar_base + 8u * (2 * sa + 2 * sb);
  auto tfull_bar = [&](int a) { return misc + 8u * a; };
  auto tempty_bar = [&](int a) { return misc + 16u + 8u * a; };
  const uint32_t wbar = misc + 32u;
  const uint32_t tmem_ptr_smem = misc + 40u;
  auto gfull = [&](int s) { return misc + 64u + 8u * s; };               // kR2GateSlots <= 4
  auto gempty = [&](int s) { return misc + 96u + 8u * s; };
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  const uint32_t smem_bias = misc + 128u;
  float* smem_bias_gen = reinterpret_cast<float*>(smem_gen + (smem_bias - smem_base));
  stage_bias(p, smem_bias_gen, BLOCK_N);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (p.pool) tma_prefetch_desc(&tmP);
    for (int s = 0; s < sa; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < sb; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * 4 * kR2EpiSplit); }
    mbar_init(wbar, 1);
    for (int s = 0; s < gate_slots; ++s) { mbar_init(gfull(s), 1); mbar_init(gempty(s), 4 * kR2EpiSplit); }
    if (gate_slots > 0) tma_prefetch_desc(&tmG);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are counted on the leader's barriers) =====================
    if (RESIDENT) {
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(wbar, (uint32_t)(2 * n_wtiles * kBHalf));
        const uint32_t wbar_l = mapa_cluster(wbar, 0);
        for (int tap = 0; tap < 9; ++tap)
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_3d_2sm(smem_b + (tap * p.kchunks + kc) * kBHalf, &tmB, wbar_l, kc * 64,
                            (int)rank * (BLOCK_N / 2), tap);
      }
      __syncwarp();
    }
    int as_ = 0, bs_ = 0;
    uint32_t aph = 0, bph = 0;
    for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
      int n, tx, typ, b;
      decode_tile(p, tile, n, tx, typ, b);
      const int x0 = tx * kR2Tw, y0 = (typ * 2 + (int)rank) * kR2Th;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(aempty(as_), aph ^ 1u);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(afull(as_), (uint32_t)(2 * kR2SlabTx));
          tma_load_4d_2sm(smem_a + as_ * kR2SlabBytes, &tmA, mapa_cluster(afull(as_), 0), kc * 64, x0 - 1, y0 - 1, b);
        }
        if (++as_ == sa) { as_ = 0; aph ^= 1u; }
        if (!RESIDENT) {
          // streamed weights travel in GROUPS of three half tiles (the vertical taps r = 0..2 of one horizontal tap
          // s) per barrier: with one tile per barrier the wait / elect / commit overhead of the issuing warp (~300
          // cycles) was paid per 4 MMAs and paced the 128-channel kernel
          for (int s = 0; s < 3; ++s) {
            mbar_wait(bempty(bs_), bph ^ 1u);
            if (elect_one()) {
              if (rank == 0) mbar_expect_tx(bfull(bs_), (uint32_t)(2 * 3 * kBHalf));
              const uint32_t bar = mapa_cluster(bfull(bs_), 0);
#pragma unroll
              for (int r = 0; r < 3; ++r)
                tma_load_3d_2sm(smem_b + (bs_ * 3 + r) * kBHalf, &tmB, bar, kc * 64, (int)rank * (BLOCK_N / 2), r * 3 + s);
            }
            if (++bs_ == sb) { bs_ = 0; bph ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp, one elected lane) =====================
    if (rank == 0) {
      if (RESIDENT) mbar_wait(wbar, 0);
      int as_ = 0, bs_ = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t accph = 0;
      for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
        mbar_wait(tempty_bar(acc), accph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const bool last_kc = kc == p.kchunks - 1;
          mbar_wait(afull(as_), aph);
          tc_fence_after();
          const uint32_t slab = smem_a + as_ * kR2SlabBytes;
          if (RESIDENT) {
            if (elect_one()) {
#pragma unroll
              for (int s = 0; s < 3; ++s) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                  const uint64_t bdesc = umma_desc_k_sw128(smem_b + ((r * 3 + s) * p.kchunks + kc) * kBHalf);
                  const uint64_t adesc =
                      umma_desc_k_sw128_sbo2(slab + (uint32_t)(r * kR2Pitch) + (uint32_t)s * 128u, kR2Pitch);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, kIdesc, (kc | s | r | k) != 0 ? 1u : 0u);
                }
              }
              umma_commit_2sm(aempty(as_));
              if (last_kc) umma_commit_2sm(tfull_bar(acc));
            }
          } else {
#pragma unroll 1
            for (int s = 0; s < 3; ++s) {
              mbar_wait(bfull(bs_), bph);
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                  const uint64_t bdesc = umma_desc_k_sw128(smem_b + (bs_ * 3 + r) * kBHalf);
                  const uint64_t adesc =
                      umma_desc_k_sw128_sbo2(slab + (uint32_t)(r * kR2Pitch) + (uint32_t)s * 128u, kR2Pitch);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_f16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, kIdesc, (kc | s | r | k) != 0 ? 1u : 0u);
                }
                umma_commit_2sm(bempty(bs_));
                if (s == 2) umma_commit_2sm(aempty(as_));
                if (s == 2 && last_kc) umma_commit_2sm(tfull_bar(acc));
              }
              if (++bs_ == sb) { bs_ = 0; bph ^= 1u; }
            }
          }
          if (++as_ == sa) { as_ = 0; aph ^= 1u; }
        }
        acc ^= 1;
        if (acc == 0) accph ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ===================== ReLU-gate TMA (data-gradient launches) =====================
    // The gate is the forward output of the layer below: 128 B per pixel, long evicted from L2.  Read per thread (one
    // row each, rows 128 B apart) every warp-level load touches 32 cache lines: 1024 L1 wavefronts per tile -- on the
    // pipe the tensor core's operand reads already fill to 80 %; the gated 64 -> 64 data gradient ran at 3.1 ms against
    // 1.6 ms forward.  Staged by TMA (128 wavefronts written + 128 read per tile) the tile costs a quarter of that,
    // and the loads run `gate_slots` chunks ahead of the epilogue.
    if (gate_slots > 0) {
      int gs = 0;
      uint32_t gph = 0;
      for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
        int n, tx, typ, b;
        decode_tile(p, tile, n, tx, typ, b);
        const int x0 = tx * kR2Tw, y0 = (typ * 2 + (int)rank) * kR2Th;
        for (int c = 0; c < BLOCK_N / 64; ++c) {
          mbar_wait(gempty(gs), gph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(gfull(gs), kGateSlotBytes);
            tma_load_4d(smem_gate + (uint32_t)gs * kGateSlotBytes, &tmG, gfull(gs), c * 64, x0, y0, b);
          }
          if (++gs == gate_slots) { gs = 0; gph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs drain their own TMEM half) =====================
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const int ly = row / kR2Tw, lx = row - ly * kR2Tw;
    int acc = 0;
    uint32_t accph = 0;
    uint32_t chunk_ctr = 0;
    float csum[kCsumSize<BLOCK_N, kR2EpiSplit>];
#pragma unroll
    for (int i = 0; i < kCsumSize<BLOCK_N, kR2EpiSplit>; ++i) csum[i] = 0.0f;
    float breg[32];
    const bool bias_regs = PLAIN && BLOCK_N == 64 && p.bias != nullptr;   // (!PLAIN: registers go to csum / gate)
#pragma unroll
    for (int i = 0; i < 32; ++i) breg[i] = bias_regs ? __ldg(p.bias + hsel * 32 + i) : 0.0f;
    const uint32_t tempty_l0 = mapa_cluster(tempty_bar(0), 0), tempty_l1 = mapa_cluster(tempty_bar(1), 0);
    GateRing gring;
    gring.smem = smem_gate; gring.full_bar = gfull(0); gring.empty_bar = gempty(0); gring.slots = gate_slots;
    gring.idx = 0; gring.phase = 0;
    for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
      int n, tx, typ, b;
      decode_tile(p, tile, n, tx, typ, b);
      const int ty = typ * 2 + (int)rank;
      const int ox = tx * kR2Tw + lx, oy = ty * kR2Th + ly;
      const bool valid = (ox < p.Wo) && (oy < p.Ho);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
      epilogue_nhwc_tile<BLOCK_N, kR2EpiSplit, true, PLAIN, true>(p, &tmC, &tmP, t_row, smem_out, smem_pool, smem_bias, smem_bias_gen,
                                                     acc ? tempty_l1 : tempty_l0, 0, tx, ty, b, ox, oy, valid, row, lane,
                                                     epi_tid, chunk_ctr, hsel, csum, bias_regs ? breg : nullptr, nullptr,
                                                     tfull_bar(acc), accph,     // (waits for the tile's MMAs itself)
                                                     &gring);                   // (gated launches always stage the gate)
      acc ^= 1;
      if (acc == 0) accph ^= 1u;
    }
    flush_colsum<BLOCK_N, kR2EpiSplit>(p, csum, lane, hsel);
    if (epi_tid < 32) {
      if (elect_one()) tma_store_wait_read<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // nobody leaves (or frees TMEM) while the peer may still signal / be signalled
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, bool RESIDENT>
static int launch_rs2(const dreamb200_conv_desc* d, int gate_slots, cudaStream_t stream) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.tw = kR2Tw; p.th = kR2Th;
  p.tiles_x = (d->Wo + kR2Tw - 1) / kR2Tw;
  p.tiles_y = (d->Ho + 2 * kR2Th - 1) / (2 * kR2Th);                  // PAIRS of 16-row tiles
  p.n_tiles = 1;
  p.B = d->B; p.Ho = d->Ho; p.Wo = d->Wo;
  p.total_tiles = p.tiles_x * p.tiles_y * d->B;
  DB_REQUIRE((long long)p.tiles_x * p.tiles_y * d->B < (1ll << 24) && p.tiles_x < 65536 && p.tiles_y < 65536,
             "conv: too many tiles for one launch (%d x %d x %d)", p.tiles_x, p.tiles_y, d->B);
  p.absmax = d->absmax;
  p.gate = reinterpret_cast<const __half*>(d->gate);
  p.out_scale = d->out_scale;
  p.colsum = d->colsum;
  p.mg_n = div_magic(1);
  p.mg_x = div_magic(p.tiles_x);
  p.mg_y = div_magic(p.tiles_y);
  p.in_stride = 1;
  p.taps = 9;
  p.kchunks = d->Cin / 64;
  p.bias = d->bias;
  p.residual = reinterpret_cast<const __half*>(d->residual);
  p.residual_f32 = d->residual_f32;
  p.y_f32 = d->y_f32;
  p.Cout_pad = d->Cout_pad;
  p.relu = d->relu;
  p.pool = d->y_pool != nullptr ? 1 : 0;
  p.store_full = d->y != nullptr ? 1 : 0;

  constexpr int kBHalf = (BLOCK_N / 2) * 128;
  const int out_bytes = 2 * kStageOutBytes + (p.pool ? 2 * kPoolBytes : 0);
  Rs2Extra x;
  x.gate_slots = gate_slots;
  const int gate_bytes = gate_slots * (int)kGateSlotBytes;
  int budget = 232448 - 1024 - out_bytes - gate_bytes - 1024 - BLOCK_N * 4;
  if (RESIDENT) {
    budget -= 9 * p.kchunks * kBHalf;
    x.sa = budget / kR2SlabBytes;
    if (x.sa > 6) x.sa = 6;
    x.sb = 1;
  } else {
    x.sa = 3;
    x.sb = (budget - x.sa * kR2SlabBytes) / (3 * kBHalf);              // groups of three half tiles
    if (x.sb > 6) x.sb = 6;
  }
  DB_REQUIRE(x.sa >= 2 && x.sb >= 1 && (RESIDENT || x.sb >= 2), "conv_rs2: shared memory budget too small");
  const int smem_bytes =
      1024 + x.sa * kR2SlabBytes + (RESIDENT ? 9 * p.kchunks : 3 * x.sb) * kBHalf + out_bytes + gate_bytes + 1024 +
      BLOCK_N * 4;

  CUtensorMap tmA, tmB, tmC, tmP, tmG;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmP, 0, sizeof(tmP));
  memset(&tmG, 0, sizeof(tmG));
  const uint32_t es4[4] = {1, 1, 1, 1};
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    uint32_t box[4] = {64, 10u, (uint32_t)kR2Rows, 1};
    if (make_tensor_map_f16(&tmA, d->x, 4, dims, str, box, es4, "rs2 activation")) return -1;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout_pad, 9};
    uint64_t str[2] = {(uint64_t)d->Cin * 2, (uint64_t)d->Cout_pad * d->Cin * 2};
    uint32_t box[3] = {64, (uint32_t)(BLOCK_N / 2), 1};
    uint32_t es[3] = {1, 1, 1};
    if (make_tensor_map_f16(&tmB, d->w, 3, dims, str, box, es, "rs2 weights")) return -1;
  }
  if (d->y_pool != nullptr) {
    const uint64_t Wp = (uint64_t)(d->Wo / 2), Hp = (uint64_t)(d->Ho / 2), C = (uint64_t)d->Cout_pad;
    uint64_t dims[4] = {C, Wp, Hp, (uint64_t)d->B};
    uint64_t str[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
    uint32_t box[4] = {64, kR2Tw / 2, kR2Th / 2, 1};
    if (make_tensor_map_f16(&tmP, d->y_pool, 4, dims, str, box, es4, "rs2 pooled output")) return -1;
  }
  if (d->y != nullptr) {
    uint64_t dims[4] = {(uint64_t)d->Cout_pad, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->y_stride_w * 2, (uint64_t)d->y_stride_h * 2, (uint64_t)d->y_stride_b * 2};
    uint32_t box[4] = {64, kR2Tw, kR2Th, 1};
    if (make_tensor_map_f16(&tmC, d->y, 4, dims, str, box, es4, "rs2 output")) return -1;
  }
  if (gate_slots > 0) {
    // the gate has the output's geometry, densely packed NHWC
    const uint64_t C = (uint64_t)d->Cout_pad;
    uint64_t dims[4] = {C, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
    uint64_t str[3] = {C * 2, (uint64_t)d->Wo * C * 2, (uint64_t)d->Ho * d->Wo * C * 2};
    uint32_t box[4] = {64, kR2Tw, kR2Th, 1};
    if (make_tensor_map_f16(&tmG, d->gate, 4, dims, str, box, es4, "rs2 gate")) return -1;
  }
  const bool plain = d->residual == nullptr && d->residual_f32 == nullptr && d->y_f32 == nullptr &&
                     d->gate == nullptr && d->out_scale == nullptr && d->colsum == nullptr && d->absmax == nullptr;
  auto kern = plain ? conv_rs2_kernel<BLOCK_N, RESIDENT, true> : conv_rs2_kernel<BLOCK_N, RESIDENT, false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[plain]) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set[plain] = true;
  }
  const int sms = device_sm_count() & ~1;
  int grid = 2 * p.total_tiles < sms ? 2 * p.total_tiles : sms;
  kern<<<grid, gate_slots > 0 ? kR2ThreadsGate : kR2Threads, smem_bytes, stream>>>(tmA, tmB, tmC, tmP, tmG, p, x);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// Returns 1 and launches when the layer qualifies for the CTA-pair kernel, 0 otherwise, <0 on error.
int try_conv_rs2(const dreamb200_conv_desc* d, cudaStream_t stream) {
  static int mode = -1;
  if (mode < 0) {
    // bit 0: 64 -> 64 channels, bit 1: 128 output channels with streamed weights (Cin >= 128), bit 2: other
    // 64-output-channel layers, bit 3: 64 -> 128 (resident).  Measured at B = 128 (tools/rs2_check.py, same box):
    // 64 -> 64 @400x400 + pool 1.47 -> 1.31 ms; 128 -> 128 @200x200 + pool 1.34 -> 1.07 ms (weights in groups of three
    // tiles per barrier; one tile per barrier was SLOWER than conv_rs, 1.60 ms); 64 -> 128 unchanged (0.58 ms) -- so
    // bits 0 and 1 are on by default.
    const char* e = getenv("DREAMB200_RS2");
    mode = e ? atoi(e) : 7;      // round 2: bit 2 on too (128 -> 64 @100x100: 0.284 -> 0.233 ms, profiles/r02_ab_rs2_tail.txt)
  }
  if (mode == 0) return 0;
  if (d->out_mode != DREAMB200_OUT_NHWC_F16 || d->taps != 9 || d->in_stride != 1) return 0;
  if (d->Ho != d->H || d->Wo != d->W) return 0;
  if (d->Cout_pad != 64 && d->Cout_pad != 128) return 0;
  if (d->Ho < 2 * kR2Th) return 0;
  for (int t = 0; t < 9; ++t)
    if (d->tap_dy[t] != t / 3 - 1 || d->tap_dx[t] != t % 3 - 1) return 0;
  const double util = (double)d->Wo * d->Ho /
                      ((double)((d->Wo + kR2Tw - 1) / kR2Tw) * ((d->Ho + 2 * kR2Th - 1) / (2 * kR2Th)) * 256.0);
  static double min_util = -1.0;
  if (min_util < 0.0) {
    const char* e = getenv("DREAMB200_RS2_MIN_UTIL");       // A/B knob: how empty the last pair of tile rows may be
    min_util = e ? atof(e) : 0.7;   // 100x100 maps fill 75 % of their tile pairs and still gain (64 -> 64: 0.156 -> 0.128 ms)
  }
  if (util < min_util) return 0;
  const int kchunks = d->Cin / 64;
  // ReLU gate of a data-gradient launch: staged through shared memory by TMA (DREAMB200_GATE_TMA=0: per-thread loads)
  static int gate_tma = -1;
  if (gate_tma < 0) {
    const char* e = getenv("DREAMB200_GATE_TMA");
    gate_tma = (e && e[0] == '0') ? 0 : 1;
  }
  if (d->gate != nullptr && (!gate_tma || ((uintptr_t)d->gate & 15) != 0)) return 0;     // -> conv_rs / conv_tc
  // 64 output channels: short tiles (36 MMAs), four chunks = four tiles ahead; 128: a tile is >= 72 64-cycle MMAs and
  // its two chunks free their slots early in the epilogue -- two slots are enough and leave the weight ring its depth
  // (with four, 256 -> 128 @100x100 went 1.27 -> 1.45 ms)
  const int gate_slots = d->gate != nullptr ? (d->Cout_pad == 64 ? kR2GateSlots : 2) : 0;
  // resident half tiles need room for at least two activation slabs next to them
  const int out_bytes = 2 * kStageOutBytes + (d->y_pool != nullptr ? 2 * kPoolBytes : 0) + gate_slots * (int)kGateSlotBytes;
  const int resident_bytes = 9 * kchunks * (d->Cout_pad / 2) * 128;
  const bool resident = 232448 - 1024 - out_bytes - 1024 - d->Cout_pad * 4 - resident_bytes >= 2 * kR2SlabBytes;
  int rc;
  if (d->Cout_pad == 64) {
    if (kchunks == 1 && !(mode & 1)) return 0;
    if (kchunks > 1 && !(mode & 4)) return 0;
    rc = resident ? launch_rs2<64, true>(d, gate_slots, stream) : launch_rs2<64, false>(d, gate_slots, stream);
  } else {
    if (kchunks == 1 ? !(mode & 8) : !(mode & 2)) return 0;
    rc = resident ? launch_rs2<128, true>(d, gate_slots, stream) : launch_rs2<128, false>(d, gate_slots, stream);
  }
  return rc == 0 ? 1 : rc;
}

}  // namespace db200
