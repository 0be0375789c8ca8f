// conv_tc2.cu -- CTA-pair (tcgen05 cta_group::2) variant of the generic tap-sum convolution (conv_tc.cu) for the wide
// layers: N = 256 accumulator columns when Cout % 256 == 0, N = 128 for the sub-pixel phase convolutions with 128
// output channels.  Default since round 2 (bit-identical to conv_tc on tools/tc2_check.py and
// tests/test_gpu_kernel_variants.py; 8-10 % faster per wide layer, profiles/r02_ab_pair_kernels.txt).
//
// Phase groups: the four sub-pixel phases of a stride-2 ConvTranspose / of a folded nearest-upsample + 3x3 conv are
// four small-tap convolutions over the SAME input with their own weights, tap offsets and interleaved output view.
// Launched one by one (round 1) each paid its own prologue, pipeline fill / drain and wave quantisation on few tiles
// (25x25 maps: 4.2 waves of CTA pairs per launch); dreamb200_conv2d_fwd_phases runs them as ONE launch in which the
// phase is simply the slowest tile coordinate -- per-phase weight / output tensor maps travel as kernel parameters.
//
// Why: the wide layers run with the tensor pipe 90-95 % active at the 1000 W power cap (SM clock 1.35 GHz), i.e. their
// rate is set by energy per FLOP.  In a pair each CTA fetches only HALF of every weight tile: per k-block 16 KB (A) +
// 16 KB (B/2) instead of 16 + 32 KB cross the L2 -> SM fabric and are written to / read from shared memory (tensor-core
// operand reads per MMA: 32 + 32 wavefronts instead of 32 + 64) -- a third less operand traffic for the same math.
//
// Pair protocol = conv_rs2.cu's: both CTAs run the same pair-tile sequence (rank r owns M-tile 2*mp + r of the same
// output-channel tile n); TMA loads count their bytes on the LEADER's full barrier; the leader's elected lane issues
// tcgen05.mma.cta_group::2 (M = 256, N = 256); tcgen05.commit multicast arrives on both CTAs' empty / accumulator-full
// barriers; each CTA drains its own TMEM half and returns the stage with an arrive on the leader's barrier.
// Warps: 0 = TMA producer (operands), 1 = MMA issuer + TMEM owner, 2..9 = epilogue, 10 = fp32 residual ring (only in
// launches with res_slots > 0: the ResNet expansion layers).
#include "common.cuh"
#include "conv_common.cuh"
#include "dreamb200.h"

#include <stdlib.h>

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int make_tensor_map_f32_sw128(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes, const uint32_t* box, const char* what);
int device_sm_count();
double conv_choose_tile(int Wo, int Ho, int in_stride, bool even, int* tw_out, int* th_out);   // conv_tc.cu

constexpr int kT2Split = 2;
constexpr int kT2Threads = 64 + 128 * kT2Split;
constexpr int kT2ThreadsRes = kT2Threads + 32;            // + the residual-ring TMA warp (launches with res_slots > 0)
constexpr int kT2MaxPhases = 4;
constexpr int kNotEligible = -100;

// per-phase tensor maps of one launch: weights [taps][Cout_pad][Cin] and the (possibly interleaved) output view
struct PhaseMaps {
  CUtensorMap b[kT2MaxPhases];
  CUtensorMap c[kT2MaxPhases];
};

template <int kT2N, bool PLAIN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT2ThreadsRes, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ PhaseMaps pm,
                const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmY, const __grid_constant__ ConvParams p) {
  constexpr int kT2BHalf = (kT2N / 2) * 128;               // this CTA's half of a weight tile: N/2 rows x 128 B
  constexpr int kT2StageBytes = kABytes + kT2BHalf;        // 32 KB (N = 256) / 24 KB (N = 128)
  constexpr int kTmemCols = 2 * kT2N;                      // two accumulator stages
  constexpr uint32_t kIdesc = umma_idesc_f16_m256(kT2N);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t smem_ab = smem_base;
  const uint32_t smem_out = smem_ab + stages * kT2StageBytes;
  const uint32_t out_bytes = 2 * kStageOutBytes + (p.pool ? 2 * kPoolBytes : 0);
  const uint32_t smem_pool = smem_out + 2 * kStageOutBytes;
  // fp32 residual ring (ResNet identity stream, non-PLAIN launches only): see the producer loop below
  const int res_slots = PLAIN ? 0 : p.res_slots;
  const uint32_t smem_res = smem_out + out_bytes;
  const uint32_t bar_base = smem_res + (uint32_t)res_slots * kResSlotBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * stages + 4);
  auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * stages + 6 + s); };
  auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * stages + 6 + 8 + s); };      // res_slots <= 8
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  const uint32_t smem_bias = bar_base + 512u;
  float* smem_bias_gen = reinterpret_cast<float*>(smem_gen + (smem_bias - smem_base));
  stage_bias(p, smem_bias_gen, kT2N);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    for (int ph = 0; ph < (p.phases > 1 ? p.phases : 1); ++ph) {
      tma_prefetch_desc(&pm.b[ph]);
      tma_prefetch_desc(&pm.c[ph]);
    }
    if (p.pool) tma_prefetch_desc(&tmP);
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * 4 * kT2Split); }
    if (res_slots > 0) {
      tma_prefetch_desc(&tmR);
      // (in-place fp32 output: a slot is released by the one thread that issued its TMA stores, see ResRing)
      const uint32_t releasers = p.res_inplace ? 1u : (uint32_t)(4 * kT2Split);
      if (p.res_inplace) tma_prefetch_desc(&tmY);
      for (int s = 0; s < res_slots; ++s) { mbar_init(rfull_bar(s), 1); mbar_init(rempty_bar(s), releasers); }
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const int num_kb = p.taps * p.kchunks;

  // pair-tile q -> output-channel tile n (fastest) and M-tile m = 2 * (q / n_tiles) + rank -> (tx, ty, b); an odd
  // M-tile count leaves the last pair's second tile at b == B: its loads are TMA zero fill, its stores clip away
  // (phase groups: the phase is the slowest coordinate, q = phase * tiles_per_phase + q')
  auto decode = [&](int q, int& sp, int& n, int& tx, int& ty, int& b) {
    sp = p.phases > 1 ? fast_div(q, p.mg_phase) : 0;
    q -= sp * p.tiles_per_phase;
    const int t = fast_div(q, p.mg_n);
    n = q - t * p.n_tiles;
    const int m = 2 * t + (int)rank;
    const int t2 = fast_div(m, p.mg_x);
    tx = m - t2 * p.tiles_x;
    b = fast_div(t2, p.mg_y);
    ty = t2 - b * p.tiles_y;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are counted on the leader's barriers) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int q = pair; q < p.total_tiles; q += n_pairs) {
      int sp, n, tx, ty, b;                                   // sp = sub-pixel phase of a phase group (0 otherwise)
      decode(q, sp, n, tx, ty, b);
      const int x0 = tx * p.tw * p.in_stride, y0 = ty * p.th * p.in_stride;
      const CUtensorMap* tmB = &pm.b[sp];
      for (int tap = 0; tap < p.taps; ++tap) {
        const int xi = x0 + p.dx[sp * p.taps + tap], yi = y0 + p.dy[sp * p.taps + tap];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_ab + stage * kT2StageBytes;
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), (uint32_t)(2 * (p.tw * p.th * 128 + kT2BHalf)));
            const uint32_t bar = mapa_cluster(full_bar(stage), 0);
            tma_load_4d_2sm(sa, &tmA, bar, kc * 64, xi, yi, b);
            tma_load_3d_2sm(sa + kABytes, tmB, bar, kc * 64, n * kT2N + (int)rank * (kT2N / 2), tap);
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp, one elected lane) =====================
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int q = pair; q < p.total_tiles; q += n_pairs) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kT2N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_ab + stage * kT2StageBytes;
          if (elect_one()) {
            const uint64_t adesc = umma_desc_k_sw128(sa);
            const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(empty_bar(stage));
            if (kb == num_kb - 1) umma_commit_2sm(tfull_bar(as));
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ===================== fp32 residual ring (launches with res_slots > 0 carry this 11th warp) =====================
    // The identity stream of a ResNet expansion layer: 4 bytes read per output element against 2 * Cin MACs.
    // Read by the epilogue threads themselves it was latency-bound -- 32 KB in flight per SM, 2.8 TB/s over the
    // chip.  Here each tile's residual (own CTA's M-tile, own barriers: no pair traffic) is fetched by TMA one
    // 64-column chunk per slot, as far ahead of the epilogue as the ring has free slots.
    // A warp of its own: when the operand producer issued these loads between the operand loads of two tiles, the ring
    // ran dry while that warp sat in the (single-stage, hence serial) operand chain of the next tile, and the operand
    // chain stalled while it waited for ring slots -- epilogue and MMA phases ran back to back instead of overlapped
    // (11 us per tile for 320 KB of traffic, 4.0 TB/s of DRAM traffic under ncu).
    if (res_slots > 0) {
      int rs = 0;
      uint32_t rph = 0;
      for (int q = pair; q < p.total_tiles; q += n_pairs) {
        int sp, n, tx, ty, b;
        decode(q, sp, n, tx, ty, b);
        for (int c = 0; c < kT2N / 64; ++c) {
          mbar_wait(rempty_bar(rs), rph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(rfull_bar(rs), (uint32_t)(2 * p.tw * p.th * 128));
            const uint32_t dst = smem_res + (uint32_t)rs * kResSlotBytes;
            tma_load_4d(dst, &tmR, rfull_bar(rs), n * kT2N + c * 64, tx * p.tw, ty * p.th, b);
            tma_load_4d(dst + kResSubBytes, &tmR, rfull_bar(rs), n * kT2N + c * 64 + 32, tx * p.tw, ty * p.th, b);
          }
          if (++rs == res_slots) { rs = 0; rph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs drain their own TMEM half) =====================
    const int q4 = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const int ly = row / p.tw, lx = row - ly * p.tw;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t chunk_ctr = 0;
    float csum[kCsumSize<kT2N, kT2Split>];
#pragma unroll
    for (int i = 0; i < kCsumSize<kT2N, kT2Split>; ++i) csum[i] = 0.0f;
    ResRing ring;
    ring.smem = smem_res; ring.full_bar = rfull_bar(0); ring.empty_bar = rempty_bar(0); ring.slots = res_slots;
    ring.idx = 0; ring.phase = 0;
    ring.y_tm = (res_slots > 0 && p.res_inplace) ? &tmY : nullptr; ring.rel_idx = 0; ring.pending = 0;
    const uint32_t tempty_l0 = mapa_cluster(tempty_bar(0), 0), tempty_l1 = mapa_cluster(tempty_bar(1), 0);
    for (int q = pair; q < p.total_tiles; q += n_pairs) {
      int sp, n, tx, ty, b;
      decode(q, sp, n, tx, ty, b);
      const int ox = tx * p.tw + lx, oy = ty * p.th + ly;
      const bool valid = (ly < p.th) && (ox < p.Wo) && (oy < p.Ho) && (b < p.B);
      const uint32_t t_row = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(as * kT2N);
      epilogue_nhwc_tile<kT2N, kT2Split, true, PLAIN>(p, &pm.c[sp], &tmP, t_row, smem_out, smem_pool, smem_bias, smem_bias_gen,
                                                      as ? tempty_l1 : tempty_l0, n, tx, ty, b, ox, oy, valid, row, lane,
                                                      epi_tid, chunk_ctr, hsel, csum, nullptr,
                                                      res_slots > 0 ? &ring : nullptr, tfull_bar(as), aphase);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    flush_colsum<kT2N, kT2Split>(p, csum, lane, hsel);
    if (epi_tid < 32) {
      if (elect_one()) tma_store_wait_read<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

template <int kT2N>
static int launch_tc2(const dreamb200_conv_desc* descs, int n_phases, cudaStream_t stream) {
  constexpr int kT2BHalf = (kT2N / 2) * 128;
  constexpr int kT2StageBytes = kABytes + kT2BHalf;
  const dreamb200_conv_desc* d = descs;                     // phase 0 carries everything the phases share
  ConvParams p;
  memset(&p, 0, sizeof(p));
  conv_choose_tile(d->Wo, d->Ho, d->in_stride, d->y_pool != nullptr, &p.tw, &p.th);
  p.pool = d->y_pool != nullptr ? 1 : 0;
  p.store_full = (d->y != nullptr) ? 1 : 0;
  p.tiles_x = (d->Wo + p.tw - 1) / p.tw;
  p.tiles_y = (d->Ho + p.th - 1) / p.th;
  p.n_tiles = d->Cout_pad / kT2N;
  p.B = d->B; p.Ho = d->Ho; p.Wo = d->Wo;
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * d->B;
  const long long pairs = (m_tiles + 1) / 2;
  DB_REQUIRE((m_tiles + 1) * p.n_tiles * n_phases < (1ll << 24) && p.tiles_x < 65536 && p.tiles_y < 65536 &&
                 p.n_tiles < 65536,
             "conv: too many tiles for one launch (%d x %d x %d x %d)", p.tiles_x, p.tiles_y, p.n_tiles, d->B);
  // (fast_div is exact for divisors below 2^16: a phase group with more pair-tiles per phase goes phase by phase)
  if (n_phases > 1 && pairs * p.n_tiles >= 65536) return kNotEligible;
  p.phases = n_phases;
  p.tiles_per_phase = (int)(pairs * p.n_tiles);
  p.mg_phase = div_magic(p.tiles_per_phase);
  p.total_tiles = p.tiles_per_phase * n_phases;
  p.absmax = d->absmax;
  p.gate = reinterpret_cast<const __half*>(d->gate);
  p.out_scale = d->out_scale;
  p.colsum = d->colsum;
  p.mg_n = div_magic(p.n_tiles);
  p.mg_x = div_magic(p.tiles_x);
  p.mg_y = div_magic(p.tiles_y);
  p.in_stride = d->in_stride;
  p.taps = d->taps;
  p.kchunks = d->Cin / 64;
  for (int ph = 0; ph < n_phases; ++ph)
    for (int t = 0; t < d->taps; ++t) {
      p.dy[ph * d->taps + t] = descs[ph].tap_dy[t];
      p.dx[ph * d->taps + t] = descs[ph].tap_dx[t];
    }
  p.bias = d->bias;
  p.residual = reinterpret_cast<const __half*>(d->residual);
  p.residual_f32 = d->residual_f32;
  p.y_f32 = d->y_f32;
  p.Cout_pad = d->Cout_pad;
  p.relu = d->relu;
  p.cout_real = d->cout_real;

  const int out_bytes = 2 * kStageOutBytes + (p.pool ? 2 * kPoolBytes : 0);
  // fp32 residual staged through shared memory (one tile = N / 64 chunk slots ahead) when the identity tensor is a
  // plain contiguous [B, Ho, Wo, Cout_pad] fp32 tensor and two operand stages still fit next to the ring
  int res_slots = 0;
  {
    const char* e = getenv("DREAMB200_RES_TMA");
    const bool on = !(e && e[0] == '0');
    if (on && d->residual_f32 != nullptr && n_phases == 1 && !p.pool && ((uintptr_t)d->residual_f32 & 15) == 0 &&
        232448 - 1024 - out_bytes - 512 - kT2N * 4 - (kT2N / 64) * (int)kResSlotBytes >= 2 * kT2StageBytes)
      res_slots = kT2N / 64;
    // with an fp32 output too (every expansion layer but a stage's last) the result leaves through the SAME slots: two
    // of them are then busy with pending stores, so the ring grows by one slot and the operand pipeline -- idle 90 % of
    // the time in these HBM-bound layers -- shrinks to a single stage
    // ... unless the layer's K loop is long (512 -> 2048: 8 k-blocks per tile): a single stage makes the operand loads
    // of a tile strictly serial (~1 us each), which then outlasts the epilogue (DREAMB200_RES_INPLACE_MAXKB, default 4)
    const char* e2 = getenv("DREAMB200_RES_INPLACE");
    const char* e3 = getenv("DREAMB200_RES_INPLACE_MAXKB");
    const int max_kb = e3 ? atoi(e3) : 4;
    if (res_slots > 0 && d->y_f32 != nullptr && ((uintptr_t)d->y_f32 & 15) == 0 && !(e2 && e2[0] == '0') &&
        d->taps * (d->Cin / 64) <= max_kb &&
        232448 - 1024 - out_bytes - 512 - kT2N * 4 - (res_slots + 1) * (int)kResSlotBytes >= kT2StageBytes) {
      res_slots += 1;
      p.res_inplace = 1;
      // Ring depth.  Five slots (two draining, one in work, two ahead) left room for ONE operand stage, which makes a
      // tile's k-block loads strictly serial; four slots (one less ahead) and two stages measured better: resnet-H
      // expansions 2.10 -> 1.94 ms per step (DREAMB200_RES_INPLACE_SLOTS=5 restores the deeper ring)
      const char* e4 = getenv("DREAMB200_RES_INPLACE_SLOTS");
      res_slots = (e4 && atoi(e4) >= 3 && atoi(e4) <= res_slots) ? atoi(e4) : res_slots - 1;
    }
  }
  p.res_slots = res_slots;
  const int budget = 232448 - 1024 - out_bytes - 512 - kT2N * 4 - res_slots * (int)kResSlotBytes;
  int stages = budget / kT2StageBytes;
  if (stages > 8) stages = 8;
  DB_REQUIRE(stages >= (p.res_inplace ? 1 : 2), "conv_tc2: not enough shared memory for the operand stages");
  p.stages = stages;
  const int smem_bytes = 1024 + stages * kT2StageBytes + out_bytes + res_slots * (int)kResSlotBytes + 512 + kT2N * 4;

  CUtensorMap tmA, tmP, tmR, tmY;
  PhaseMaps pm;
  memset(&pm, 0, sizeof(pm));
  memset(&tmP, 0, sizeof(tmP));
  memset(&tmR, 0, sizeof(tmR));
  memset(&tmY, 0, sizeof(tmY));
  if (res_slots > 0) {
    const uint64_t C = (uint64_t)d->Cout_pad;
    uint64_t dims[4] = {C, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
    uint64_t str[3] = {C * 4, (uint64_t)d->Wo * C * 4, (uint64_t)d->Ho * d->Wo * C * 4};
    uint32_t box[4] = {32, (uint32_t)p.tw, (uint32_t)p.th, 1};
    if (make_tensor_map_f32_sw128(&tmR, d->residual_f32, 4, dims, str, box, "tc2 fp32 residual")) return -1;
    if (p.res_inplace && make_tensor_map_f32_sw128(&tmY, d->y_f32, 4, dims, str, box, "tc2 fp32 output")) return -1;
  }
  const uint32_t es4[4] = {1, 1, 1, 1};
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    const uint32_t s = (uint32_t)d->in_stride;
    uint32_t box[4] = {64, (uint32_t)p.tw * s, (uint32_t)p.th * s, 1};
    uint32_t es[4] = {1, s, s, 1};
    DB_REQUIRE(box[1] <= 256 && box[2] <= 256, "conv_tc2: TMA box too large (%u x %u)", box[1], box[2]);
    if (make_tensor_map_f16(&tmA, d->x, 4, dims, str, box, es, "tc2 activation")) return -1;
  }
  for (int ph = 0; ph < n_phases; ++ph) {
    const dreamb200_conv_desc* dp = descs + ph;
    {
      uint64_t dims[3] = {(uint64_t)dp->Cin, (uint64_t)dp->Cout_pad, (uint64_t)dp->taps};
      uint64_t str[2] = {(uint64_t)dp->Cin * 2, (uint64_t)dp->Cout_pad * dp->Cin * 2};
      uint32_t box[3] = {64, (uint32_t)(kT2N / 2), 1};
      uint32_t es[3] = {1, 1, 1};
      if (make_tensor_map_f16(&pm.b[ph], dp->w, 3, dims, str, box, es, "tc2 weights")) return -1;
    }
    if (dp->y != nullptr) {
      uint64_t dims[4] = {(uint64_t)dp->Cout_pad, (uint64_t)dp->Wo, (uint64_t)dp->Ho, (uint64_t)dp->B};
      uint64_t str[3] = {(uint64_t)dp->y_stride_w * 2, (uint64_t)dp->y_stride_h * 2, (uint64_t)dp->y_stride_b * 2};
      uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
      if (make_tensor_map_f16(&pm.c[ph], dp->y, 4, dims, str, box, es4, "tc2 output")) return -1;
    }
  }
  if (d->y_pool != nullptr) {
    const uint64_t Wp = (uint64_t)(d->Wo / 2), Hp = (uint64_t)(d->Ho / 2), C = (uint64_t)d->Cout_pad;
    uint64_t dims[4] = {C, Wp, Hp, (uint64_t)d->B};
    uint64_t str[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw / 2, (uint32_t)p.th / 2, 1};
    if (make_tensor_map_f16(&tmP, d->y_pool, 4, dims, str, box, es4, "tc2 pooled output")) return -1;
  }
  const bool plain = d->residual == nullptr && d->residual_f32 == nullptr && d->y_f32 == nullptr &&
                     d->gate == nullptr && d->out_scale == nullptr && d->colsum == nullptr && d->absmax == nullptr;
  auto kern = plain ? conv_tc2_kernel<kT2N, true> : conv_tc2_kernel<kT2N, false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[plain]) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set[plain] = true;
  }
  const int sms = device_sm_count() & ~1;
  const int grid = 2 * p.total_tiles < sms ? 2 * p.total_tiles : sms;
  kern<<<grid, res_slots > 0 ? kT2ThreadsRes : kT2Threads, smem_bytes, stream>>>(tmA, pm, tmP, tmR, tmY, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static bool tc2_enabled() {
  // on by default since round 2; DREAMB200_TC2=0 switches back for A/B runs.  Read per call so a test can flip it.
  const char* e = getenv("DREAMB200_TC2");
  return !(e && e[0] == '0');
}

// Returns 1 and launches when the layer qualifies (and the kernel is switched on), 0 otherwise, <0 on error.
int try_conv_tc2(const dreamb200_conv_desc* d, cudaStream_t stream) {
  if (!tc2_enabled()) return 0;
  if (d->out_mode != DREAMB200_OUT_NHWC_F16 || d->Cout_pad % 256 != 0) return 0;
  if ((long long)d->B * d->Ho * d->Wo < 2 * 128) return 0;            // fewer than two M-tiles: nothing to pair
  const int rc = launch_tc2<256>(d, 1, stream);
  return rc == 0 ? 1 : rc;
}

// A group of sub-pixel phase convolutions (same input, shapes, tap count and epilogue; own weights / tap offsets /
// output view) as ONE launch.  Returns 1 when launched, 0 when the group does not qualify (the caller then launches
// the phases one by one), <0 on error.
int try_conv_tc2_phases(const dreamb200_conv_desc* descs, int n_phases, cudaStream_t stream) {
  if (!tc2_enabled()) return 0;
  { const char* e = getenv("DREAMB200_PHASE_GROUPS"); if (e && e[0] == '0') return 0; }
  if (n_phases < 2 || n_phases > kT2MaxPhases) return 0;
  const dreamb200_conv_desc* d = descs;
  if (d->out_mode != DREAMB200_OUT_NHWC_F16 || d->Cout_pad % 128 != 0) return 0;
  if ((long long)d->B * d->Ho * d->Wo < 2 * 128 || d->taps * n_phases > DREAMB200_MAX_TAPS) return 0;
  if (d->y_pool || d->residual || d->residual_f32 || d->y_f32 || d->gate || d->out_scale || d->colsum || d->absmax)
    return 0;
  for (int ph = 1; ph < n_phases; ++ph) {
    const dreamb200_conv_desc* q = descs + ph;
    if (q->x != d->x || q->B != d->B || q->H != d->H || q->W != d->W || q->Cin != d->Cin ||
        q->in_stride != d->in_stride || q->taps != d->taps || q->Cout_pad != d->Cout_pad || q->Ho != d->Ho ||
        q->Wo != d->Wo || q->out_mode != d->out_mode || q->bias != d->bias || q->relu != d->relu ||
        q->y_stride_w != d->y_stride_w || q->y_stride_h != d->y_stride_h || q->y_stride_b != d->y_stride_b ||
        q->y == nullptr || q->y_pool || q->residual || q->residual_f32 || q->y_f32 || q->gate || q->out_scale ||
        q->colsum || q->absmax)
      return 0;
  }
  if (d->y == nullptr) return 0;
  const int rc = d->Cout_pad % 256 == 0 ? launch_tc2<256>(descs, n_phases, stream)
                                         : launch_tc2<128>(descs, n_phases, stream);
  if (rc == kNotEligible) return 0;
  return rc == 0 ? 1 : rc;
}

}  // namespace db200
