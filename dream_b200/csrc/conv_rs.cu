// conv_rs.cu -- "row-shared" variant of the 3x3 / stride 1 / pad 1 tensor-core convolution for the layers that are
// L2->SMEM bandwidth bound in conv_tc.cu (few output channels: every MMA k-block there re-fetches a 16 KB
// activation tile for only 64 or 128 output channels).
//
// Output tile = 8 (x) x 16 (y) pixels.  For one 64-channel chunk the producer loads three activation slabs, one
// per horizontal tap s: a TMA box (64 ch, 8 x, 18 y) at x0-1+s, y0-1.  8 pixels x 128 B is exactly one 1024-byte
// SWIZZLE_128B row group, so inside a slab image row yy is row group yy, and the A operand of the vertical tap r
// is the SAME slab read from byte offset r*1024 -- a 1024-aligned start address, i.e. a perfectly ordinary UMMA
// descriptor.  Three taps share one load: activation traffic per tile and chunk drops from 9 x 16 KB to 3 x 18 KB.
// When all 9 x Cin/64 weight tiles fit (64 -> 64 channels: 72 KB) they are loaded once per CTA and stay resident,
// otherwise they stream through their own ring.  Epilogue, pooling fusion and parameter block are conv_tc's.
#include "common.cuh"
#include "conv_common.cuh"
#include "dreamb200.h"

#include <stdlib.h>

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int device_sm_count();

constexpr int kRsTw = 8, kRsTh = 16;

struct RsExtra {
  int sa, sb;        // ring depths: activation slabs, weight tiles
};

// UMMA descriptor, K-major SWIZZLE_128B, with an explicit stride between 8-row groups (SBO).  The hardware
// applies the 128B swizzle to the computed shared-memory address, so a start address that is a multiple of 128 B
// (not of 1024 B) still reads what TMA wrote -- that is what lets the HALO variant shift by one pixel.
__device__ __forceinline__ uint64_t umma_desc_k_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// SUBTILES: 1 or 2 vertically adjacent 8x16 output tiles per CTA step; they share every weight tile (halves the
//           weight traffic per MMA) and one activation slab of 16*SUBTILES+2 image rows.
// HALO    : one slab of 10 pixels x (rows) per channel chunk serves all 9 taps (row pitch 1280 B, horizontal tap s
//           = +128 B on the start address) instead of one 8-pixel slab per horizontal tap.
constexpr int kRsEpiSplit = 2;                          // 8 epilogue warps (see conv_common.cuh)
constexpr int kRsThreads = 64 + 128 * kRsEpiSplit;

// HEAD    : the network head -- BLOCK_N = 16 accumulator columns of which the first cout_real are channels, written as
//           fp32 NCHW straight from the accumulators (no staging, no TMA store); resident weights only.
template <int BLOCK_N, bool RESIDENT, int SUBTILES, bool HALO, bool HEAD = false, bool PLAIN = false>
__global__ void __launch_bounds__(kRsThreads, 1)
conv_rs_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
               const __grid_constant__ ConvParams p, const __grid_constant__ RsExtra x) {
  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kRows = kRsTh * SUBTILES + 2;                            // image rows per slab
  constexpr int kPitch = HALO ? 1280 : 1024;                             // bytes per image row inside a slab
  constexpr int kSlabBytes = ((kRows * kPitch + 1023) / 1024) * 1024;
  constexpr int kSlabTx = kRows * kPitch;                                // bytes TMA delivers per slab
  constexpr int kSlabsPerChunk = HALO ? 1 : 3;
  constexpr int kAccCols = 2 * SUBTILES * BLOCK_N;
  constexpr int kTmemCols = kAccCols <= 128 ? 128 : kAccCols <= 256 ? 256 : 512;
  constexpr uint32_t kIdesc = umma_idesc_f16_m128(BLOCK_N);
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int sa = x.sa, sb = x.sb;
  const int n_wtiles = 9 * p.kchunks;                                   // resident weight tiles
  const uint32_t smem_a = smem_base;                                    // sa slabs
  const uint32_t smem_b = smem_a + sa * kSlabBytes;                     // sb tiles (ring) or n_wtiles tiles (resident)
  const uint32_t smem_out = smem_b + (RESIDENT ? n_wtiles : sb) * kBBytes;
  const uint32_t smem_pool = smem_out + (HEAD ? 0 : 2 * kStageOutBytes);
  const uint32_t bar_base = smem_pool + (p.pool ? 2 * kPoolBytes : 0);
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (sa + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * sa + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * sa + sb + s); };
  const uint32_t misc = bar_base + 8u * (2 * sa + 2 * sb);
  auto tfull_bar = [&](int a) { return misc + 8u * a; };
  auto tempty_bar = [&](int a) { return misc + 16u + 8u * a; };
  const uint32_t wbar = misc + 32u;
  const uint32_t tmem_ptr_smem = misc + 40u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  const uint32_t smem_bias = misc + 64u;                                // BLOCK_N fp32
  float* smem_bias_gen = reinterpret_cast<float*>(smem_gen + (smem_bias - smem_base));
  stage_bias(p, smem_bias_gen, BLOCK_N);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!HEAD) tma_prefetch_desc(&tmC);
    if (p.pool) tma_prefetch_desc(&tmP);
    for (int s = 0; s < sa; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int s = 0; s < sb; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * kRsEpiSplit * SUBTILES); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===================== TMA producer (whole warp in uniform control flow, one elected lane issues) ==========
    {
      if (RESIDENT) {   // this CTA's n is fixed (n_tiles == 1 for resident layers): all weight tiles once
        if (elect_one()) {
          mbar_expect_tx(wbar, (uint32_t)(n_wtiles * kBBytes));
          for (int tap = 0; tap < 9; ++tap)
            for (int kc = 0; kc < p.kchunks; ++kc)
              tma_load_3d(smem_b + (tap * p.kchunks + kc) * kBBytes, &tmB, wbar, kc * 64, 0, tap);
        }
        __syncwarp();
      }
      int as_ = 0, bs_ = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int n, tx, ty, b;
        decode_tile(p, tile, n, tx, ty, b);
        const int x0 = tx * kRsTw, y0 = ty * kRsTh * SUBTILES;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          for (int s = 0; s < 3; ++s) {
            if (!HALO || s == 0) {
              mbar_wait(aempty(as_), aph ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(afull(as_), (uint32_t)kSlabTx);
                tma_load_4d(smem_a + as_ * kSlabBytes, &tmA, afull(as_), kc * 64, x0 - 1 + (HALO ? 0 : s), y0 - 1, b);
              }
              if (++as_ == sa) { as_ = 0; aph ^= 1u; }
            }
            if (!RESIDENT) {
              for (int r = 0; r < 3; ++r) {
                mbar_wait(bempty(bs_), bph ^ 1u);
                if (elect_one()) {
                  mbar_expect_tx(bfull(bs_), (uint32_t)kBBytes);
                  tma_load_3d(smem_b + bs_ * kBBytes, &tmB, bfull(bs_), kc * 64, n * BLOCK_N, r * 3 + s);
                }
                if (++bs_ == sb) { bs_ = 0; bph ^= 1u; }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in uniform control flow, one elected lane issues) ==========
    {
      if (RESIDENT) mbar_wait(wbar, 0);
      int as_ = 0, bs_ = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t accph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), accph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SUBTILES * BLOCK_N);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const bool last_kc = kc == p.kchunks - 1;
          uint32_t slab = 0;
#pragma unroll(RESIDENT ? 3 : 1)
          for (int s = 0; s < 3; ++s) {
            if (!HALO || s == 0) {
              mbar_wait(afull(as_), aph);
              tc_fence_after();
              slab = smem_a + as_ * kSlabBytes;
            }
            const bool slab_done = !HALO || s == 2;
            // vertical tap r = image-row offset into the slab; horizontal tap s (HALO only) = one pixel = 128 B.
            // ONE election per barrier wait: the body is straight-line UTCHMMA + commits.
            if (RESIDENT) {
              if (elect_one()) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                  const uint64_t bdesc = umma_desc_k_sw128(smem_b + ((r * 3 + s) * p.kchunks + kc) * kBBytes);
#pragma unroll
                  for (int sub = 0; sub < SUBTILES; ++sub) {
                    const uint32_t a_addr =
                        slab + (uint32_t)((r + sub * kRsTh) * kPitch) + (HALO ? (uint32_t)s * 128u : 0u);
                    const uint64_t adesc = umma_desc_k_sw128_sbo(a_addr, kPitch);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_f16(d_tmem + (uint32_t)(sub * BLOCK_N), adesc + 2u * k, bdesc + 2u * k, kIdesc,
                               (kc | s | r | k) != 0 ? 1u : 0u);
                  }
                }
                if (slab_done) umma_commit(aempty(as_));
                if (last_kc && s == 2) umma_commit(tfull_bar(acc));
              }
            } else {
#pragma unroll 1
              for (int r = 0; r < 3; ++r) {
                mbar_wait(bfull(bs_), bph);
                tc_fence_after();
                if (elect_one()) {
                  const uint64_t bdesc = umma_desc_k_sw128(smem_b + bs_ * kBBytes);
#pragma unroll
                  for (int sub = 0; sub < SUBTILES; ++sub) {
                    const uint32_t a_addr =
                        slab + (uint32_t)((r + sub * kRsTh) * kPitch) + (HALO ? (uint32_t)s * 128u : 0u);
                    const uint64_t adesc = umma_desc_k_sw128_sbo(a_addr, kPitch);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_f16(d_tmem + (uint32_t)(sub * BLOCK_N), adesc + 2u * k, bdesc + 2u * k, kIdesc,
                               (kc | s | r | k) != 0 ? 1u : 0u);
                  }
                  umma_commit(bempty(bs_));
                  if (r == 2 && slab_done) umma_commit(aempty(as_));
                  if (r == 2 && s == 2 && last_kc) umma_commit(tfull_bar(acc));
                }
                if (++bs_ == sb) { bs_ = 0; bph ^= 1u; }
              }
            }
            if (slab_done) {
              if (++as_ == sa) { as_ = 0; aph ^= 1u; }
            }
          }
        }
        acc ^= 1;
        if (acc == 0) accph ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9: two per TMEM lane quarter, 32 columns each) =====================
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const int ly = row / kRsTw, lx = row - ly * kRsTw;
    int acc = 0;
    uint32_t accph = 0;
    uint32_t chunk_ctr = 0;
    float csum[kCsumSize<BLOCK_N, kRsEpiSplit>];
#pragma unroll
    for (int i = 0; i < kCsumSize<BLOCK_N, kRsEpiSplit>; ++i) csum[i] = 0.0f;
    float breg[32];
    const bool bias_regs = PLAIN && BLOCK_N == 64 && p.n_tiles == 1 && p.bias != nullptr;
#pragma unroll
    for (int i = 0; i < 32; ++i) breg[i] = bias_regs ? __ldg(p.bias + hsel * 32 + i) : 0.0f;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int n, tx, ty, b;
      decode_tile(p, tile, n, tx, ty, b);
      if constexpr (HEAD) {
        mbar_wait(tfull_bar(acc), accph);
        tc_fence_after();
        // fp32 NCHW head: this warp's 8 of the 16 accumulator columns (hsel), channels < cout_real are real
        const int ox = tx * kRsTw + lx, oy = ty * kRsTh + ly;
        uint32_t v[8];
        tmem_ld_32x32b_x8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + hsel * 8), v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        if (ox < p.Wo && oy < p.Ho) {
          const size_t plane = (size_t)p.Ho * p.Wo;
          float* o = p.out_f32 + (size_t)b * p.cout_real * plane + (size_t)oy * p.Wo + ox;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int ch = hsel * 8 + k;
            if (ch < p.cout_real) {
              float f = __uint_as_float(v[k]) + smem_bias_gen[ch];
              if (p.relu) f = fmaxf(f, 0.0f);
              o[(size_t)ch * plane] = f;
            }
          }
        }
        acc ^= 1;
        if (acc == 0) accph ^= 1u;
        continue;
      }
#pragma unroll 1
      for (int sub = 0; sub < SUBTILES; ++sub) {
        const int tys = ty * SUBTILES + sub;                  // 16-row tile index of this sub-tile
        const int ox = tx * kRsTw + lx, oy = tys * kRsTh + ly;
        const bool valid = (ox < p.Wo) && (oy < p.Ho);
        const uint32_t t_row =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * SUBTILES + sub) * BLOCK_N);
        epilogue_nhwc_tile<BLOCK_N, kRsEpiSplit, false, PLAIN>(p, &tmC, &tmP, t_row, smem_out, smem_pool, smem_bias, smem_bias_gen,
                                                 tempty_bar(acc), n, tx, tys, b, ox, oy, valid, row, lane, epi_tid,
                                                 chunk_ctr, hsel, csum, bias_regs ? breg : nullptr, nullptr,
                                                 sub == 0 ? tfull_bar(acc) : 0u, accph);   // (sub-tile 0 waits for the MMAs)
      }
      acc ^= 1;
      if (acc == 0) accph ^= 1u;
    }
    if constexpr (!HEAD) {
      flush_colsum<BLOCK_N, kRsEpiSplit>(p, csum, lane, hsel);
      if (epi_tid < 32) {
        if (elect_one()) tma_store_wait_read<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N, bool RESIDENT, int SUBTILES, bool HALO, bool HEAD = false>
static int launch_rs(const dreamb200_conv_desc* d, cudaStream_t stream) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.tw = kRsTw; p.th = kRsTh;
  p.tiles_x = (d->Wo + kRsTw - 1) / kRsTw;
  p.tiles_y = (d->Ho + kRsTh * SUBTILES - 1) / (kRsTh * SUBTILES);
  p.n_tiles = d->Cout_pad / BLOCK_N;
  p.B = d->B; p.Ho = d->Ho; p.Wo = d->Wo;
  p.total_tiles = p.tiles_x * p.tiles_y * p.n_tiles * d->B;
  DB_REQUIRE((long long)p.tiles_x * p.tiles_y * p.n_tiles * d->B < (1ll << 24) && p.tiles_x < 65536 &&
                 p.tiles_y < 65536 && p.n_tiles < 65536,
             "conv: too many tiles for one launch (%d x %d x %d x %d)", p.tiles_x, p.tiles_y, p.n_tiles, d->B);
  p.absmax = d->absmax;
  p.gate = reinterpret_cast<const __half*>(d->gate);
  p.out_scale = d->out_scale;
  p.colsum = d->colsum;
  p.mg_n = div_magic(p.n_tiles);
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
  p.pool = (!HEAD && d->y_pool != nullptr) ? 1 : 0;
  p.store_full = d->y != nullptr ? 1 : 0;
  p.out_f32 = HEAD ? reinterpret_cast<float*>(d->y) : nullptr;
  p.cout_real = d->cout_real;

  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kRows = kRsTh * SUBTILES + 2;
  constexpr int kPitch = HALO ? 1280 : 1024;
  constexpr int kSlabBytes = ((kRows * kPitch + 1023) / 1024) * 1024;
  const int out_bytes = HEAD ? 0 : 2 * kStageOutBytes + (p.pool ? 2 * kPoolBytes : 0);
  RsExtra x;
  int budget = 232448 - 1024 - out_bytes - 1024 - BLOCK_N * 4;
  if (RESIDENT) {
    budget -= 9 * p.kchunks * kBBytes;
    x.sa = budget / kSlabBytes;
    if (x.sa > 8) x.sa = 8;
    x.sb = 1;
  } else {
    // split the budget between the two rings in proportion to what one channel chunk consumes
    const int a_need = (HALO ? 1 : 3) * kSlabBytes, b_need = 9 * kBBytes;
    x.sa = (int)((long long)budget * a_need / (a_need + b_need)) / kSlabBytes;
    if (x.sa < (HALO ? 2 : 3)) x.sa = HALO ? 2 : 3;
    x.sb = (budget - x.sa * kSlabBytes) / kBBytes;
    if (x.sb > 12) x.sb = 12;
  }
  DB_REQUIRE(x.sa >= 2 && x.sb >= 1 && (RESIDENT || x.sb >= 3), "conv_rs: shared memory budget too small");
  const int smem_bytes =
      1024 + x.sa * kSlabBytes + (RESIDENT ? 9 * p.kchunks : x.sb) * kBBytes + out_bytes + 1024 + BLOCK_N * 4;

  CUtensorMap tmA, tmB, tmC, tmP;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmP, 0, sizeof(tmP));
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    uint32_t box[4] = {64, HALO ? 10u : 8u, (uint32_t)kRows, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (make_tensor_map_f16(&tmA, d->x, 4, dims, str, box, es, "rs activation")) return -1;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout_pad, 9};
    uint64_t str[2] = {(uint64_t)d->Cin * 2, (uint64_t)d->Cout_pad * d->Cin * 2};
    uint32_t box[3] = {64, (uint32_t)BLOCK_N, 1};
    uint32_t es[3] = {1, 1, 1};
    if (make_tensor_map_f16(&tmB, d->w, 3, dims, str, box, es, "rs weights")) return -1;
  }
  if (!HEAD && d->y_pool != nullptr) {
    const uint64_t Wp = (uint64_t)(d->Wo / 2), Hp = (uint64_t)(d->Ho / 2), C = (uint64_t)d->Cout_pad;
    uint64_t dims[4] = {C, Wp, Hp, (uint64_t)d->B};
    uint64_t str[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
    uint32_t box[4] = {64, kRsTw / 2, kRsTh / 2, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (make_tensor_map_f16(&tmP, d->y_pool, 4, dims, str, box, es, "rs pooled output")) return -1;
  }
  if (!HEAD && d->y != nullptr) {
    uint64_t dims[4] = {(uint64_t)d->Cout_pad, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->y_stride_w * 2, (uint64_t)d->y_stride_h * 2, (uint64_t)d->y_stride_b * 2};
    uint32_t box[4] = {64, kRsTw, kRsTh, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (make_tensor_map_f16(&tmC, d->y, 4, dims, str, box, es, "rs output")) return -1;
  }
  // PLAIN: none of the optional epilogue inputs / outputs (see epilogue_nhwc_tile)
  // (measured, B=128: PLAIN is 10 % faster for the resident 64 -> 128 kernel, 1-5 % for the 64-channel ones, and 8 %
  //  SLOWER for the streamed 128-channel pair kernel -- a code-generation effect; that variant keeps the generic epilogue)
  const bool plain = !HEAD && !(BLOCK_N == 128 && !RESIDENT) && d->residual == nullptr && d->residual_f32 == nullptr && d->y_f32 == nullptr &&
                     d->gate == nullptr && d->out_scale == nullptr && d->colsum == nullptr && d->absmax == nullptr;
  auto kern = plain ? conv_rs_kernel<BLOCK_N, RESIDENT, SUBTILES, HALO, HEAD, !HEAD>
                    : conv_rs_kernel<BLOCK_N, RESIDENT, SUBTILES, HALO, HEAD, false>;
  static bool attr_set[2] = {false, false};
  if (!attr_set[plain]) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set[plain] = true;
  }
  const int sms = device_sm_count();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  kern<<<grid, kRsThreads, smem_bytes, stream>>>(tmA, tmB, tmC, tmP, p, x);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Returns 1 and launches when the layer qualifies for the row-shared kernel, 0 when conv_tc should handle it,
// <0 on error.
int try_conv_rs(const dreamb200_conv_desc* d, cudaStream_t stream) {
  static double min_util = -1.0;
  if (min_util < 0.0) {
    const char* e = getenv("DREAMB200_RS_MIN_UTIL");   // 2.0 disables the kernel (A/B measurements)
    min_util = e ? atof(e) : 0.85;
  }
  if (d->taps != 9 || d->in_stride != 1) return 0;
  if (d->Ho != d->H || d->Wo != d->W) return 0;
  if (d->Cout_pad % 256 == 0) return 0;                 // wide layers are tensor-bound already in conv_tc
  for (int t = 0; t < 9; ++t)
    if (d->tap_dy[t] != t / 3 - 1 || d->tap_dx[t] != t % 3 - 1) return 0;
  const double util = (double)d->Wo * d->Ho /
                      ((double)((d->Wo + kRsTw - 1) / kRsTw) * ((d->Ho + kRsTh - 1) / kRsTh) * 128.0);
  if (d->out_mode == DREAMB200_OUT_NCHW_F32) {
    // the fp32 NCHW head (Cout_pad == 16): one slab per tile instead of nine 16 KB tap tiles from L2
    static const int head = env_flag("DREAMB200_RS_HEAD", 1);
    if (!head || d->Cout_pad != 16 || d->Cin != 64 || util < 0.5) return 0;
    const int rc = launch_rs<16, true, 1, true, true>(d, stream);
    return rc == 0 ? 1 : rc;
  }
  if (d->out_mode != DREAMB200_OUT_NHWC_F16) return 0;
  if (util < min_util) return 0;
  // variants (env overrides are for A/B measurements): DREAMB200_RS_PAIR: two stacked tiles share each weight
  // tile (streamed-weight layers); DREAMB200_RS_HALO: one 10-pixel slab serves all horizontal taps.
  static const int pair = env_flag("DREAMB200_RS_PAIR", 1);
  static const int halo = env_flag("DREAMB200_RS_HALO", 1);
  // DREAMB200_RS_RESIDENT_WIDE: keep all 9 x Cin/64 weight tiles resident also for 64 -> 128 and 128 -> 64 channels
  // (144 KB of weights + two activation slabs): no weight re-streaming per tile, 16-row tiles (no pair padding).
  static const int res_wide = env_flag("DREAMB200_RS_RESIDENT_WIDE", 1);
  int rc;
  if (res_wide && halo && d->y_pool == nullptr &&
      ((d->Cout_pad == 128 && d->Cin == 64) || (d->Cout_pad == 64 && d->Cin == 128))) {
    rc = d->Cout_pad == 128 ? launch_rs<128, true, 1, true>(d, stream) : launch_rs<64, true, 1, true>(d, stream);
  } else if (d->Cout_pad % 128 == 0) {
    if (pair && halo) rc = launch_rs<128, false, 2, true>(d, stream);
    else if (pair) rc = launch_rs<128, false, 2, false>(d, stream);
    else if (halo) rc = launch_rs<128, false, 1, true>(d, stream);
    else rc = launch_rs<128, false, 1, false>(d, stream);
  } else if (d->Cout_pad == 64 && d->Cin == 64) {
    if (halo) rc = launch_rs<64, true, 1, true>(d, stream);
    else rc = launch_rs<64, true, 1, false>(d, stream);
  } else {
    if (pair && halo) rc = launch_rs<64, false, 2, true>(d, stream);
    else if (pair) rc = launch_rs<64, false, 2, false>(d, stream);
    else if (halo) rc = launch_rs<64, false, 1, true>(d, stream);
    else rc = launch_rs<64, false, 1, false>(d, stream);
  }
  return rc == 0 ? 1 : rc;
}

}  // namespace db200
