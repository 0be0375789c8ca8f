// conv_rs3.cu -- "two output rows per accumulator row" variant of the CTA-pair slab kernel (conv_rs2.cu) for the
// 3x3 layers with 64 input and 64 output channels (vgg `layer_0_1_down.2` at full resolution, the 64 -> 64 layers of
// the decoder / head).
//
// Why.  With N = 64 the pair kernel's MMA (M = 256, N = 64, K = 16: 32 tensor cycles) makes each SM read 32 (A: 128
// pixels x 32 B) + 8 (B: 32 channels x 32 B) shared-memory wavefronts: 40 wavefronts per 32 cycles, i.e. the operand
// reads -- not the tensor pipe -- bound the layer at 80 % (measured: 1.13 PFLOP/s at 400x400, the largest gap of the
// vgg-Q step).  The A operand is the expensive one, and a 3x3 convolution reads every input row for THREE output rows.
// So let one accumulator row serve TWO vertically adjacent output pixels: accumulator row m = (x, j) stands for input
// pixel column x and output rows 2j and 2j+1; columns 0..63 of the accumulator are the 64 channels of output row 2j,
// columns 64..127 those of row 2j+1.  For the input-row offset t = 0..3 (input row 2j - 1 + t) and column tap s the A
// operand is the slab window at (t, s) with 2 image rows between 8-row groups, and the B operand is
//     t = 0: W[r=0,s] -> columns 0..63                     (N = 64)
//     t = 1: [ W[r=1,s] | W[r=0,s] ] -> columns 0..127     (N = 128)
//     t = 2: [ W[r=2,s] | W[r=1,s] ] -> columns 0..127     (N = 128)
//     t = 3: W[r=2,s] -> columns 64..127                   (N = 64)
// Same tensor cycles per output as before (24 x 64 + 24 x 32 per 256 pixels), but 24 x 48 + 24 x 40 = 2112 operand
// wavefronts instead of 2880: the layer becomes tensor-bound.  In a cta_group::2 MMA each CTA supplies half of B's N
// rows, so for the N = 128 MMAs rank 0 keeps W[r=t,s] and rank 1 keeps W[r=t-1,s] at the same shared-memory offset, for
// the N = 64 MMAs each rank keeps its 32-channel half.  Every output element still receives its 36 partial products
// in the order of conv_rs2's resident path (s, then r, then k) -- the very first MMA of the columns 64..127 is issued
// as an N = 64 pair at (s = 0, t = 1) so that it can overwrite instead of accumulate -- hence results are BIT-IDENTICAL
// to conv_rs2 / conv_rs (tests/test_gpu_kernel_variants.py).
//
// The pair is two horizontally adjacent 8-pixel columns of 32 output rows; slab = 34 image rows x 10 pixels x 128 B.
// 2x2 max pooling pairs exactly the two halves of an accumulator row (vertical, in the thread) and lanes l, l^1
// (horizontal): the pooled epilogue is 32 fmax + 16 shuffles per thread.  PLAIN launches only (inference and the
// training forward of the un-pooled layers); everything else stays on conv_rs2.  (A gated data-gradient variant was
// built and measured in round 2: with the 68 KB of weights only two 43 KB slabs and two gate chunks fit, and it ran
// 2.14 ms against conv_rs2's 2.03 ms on the 400x400 layer -- removed again.)
#include "common.cuh"
#include "conv_common.cuh"
#include "dreamb200.h"

#include <stdlib.h>

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int device_sm_count();

constexpr int kR3Tw = 8, kR3J = 16;                                       // accumulator rows: 8 pixels x 16 row pairs
constexpr int kR3Rows = 2 * kR3J + 2;                                     // image rows per slab
constexpr int kR3Pitch = 1280;                                            // 10 pixels x 128 B
constexpr int kR3SlabTx = kR3Rows * kR3Pitch;
constexpr int kR3SlabBytes = ((kR3SlabTx + 1023) / 1024) * 1024;
constexpr int kR3EpiSplit = 2;
constexpr int kR3Threads = 64 + 128 * kR3EpiSplit;
constexpr int kR3PoolBytes = 64 * 128;                                    // pooled tile: 4 x 16 pixels x 64 channels
// resident weights of one CTA (byte offsets from smem_w)
constexpr uint32_t kR3Full = 64 * 128, kR3Half = 32 * 128;
__host__ __device__ constexpr uint32_t r3_full(int t, int s) { return (uint32_t)((t == 1 ? s - 1 : 2 + s)) * kR3Full; }  // (1,1),(1,2),(2,0..2)
__host__ __device__ constexpr uint32_t r3_h0(int s) { return 5u * kR3Full + (uint32_t)s * kR3Half; }
__host__ __device__ constexpr uint32_t r3_h3(int s) { return 5u * kR3Full + 3u * kR3Half + (uint32_t)s * kR3Half; }
constexpr uint32_t kR3H1 = 5u * kR3Full + 6u * kR3Half;
constexpr uint32_t kR3WBytes = kR3H1 + kR3Half;                           // 69632

struct Rs3Extra {
  int sa;            // activation slab ring depth
};

// K-major SWIZZLE_128B operand whose 8-row groups are `sbo_bytes` apart (a window into the halo slab)
__device__ __forceinline__ uint64_t r3_desc_k_sw128(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool POOL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kR3Threads, 1)
conv_rs3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC0, const __grid_constant__ CUtensorMap tmC1,
                const __grid_constant__ CUtensorMap tmP, const __grid_constant__ ConvParams p,
                const __grid_constant__ Rs3Extra x) {
  constexpr uint32_t kI64 = umma_idesc_f16_m256(64), kI128 = umma_idesc_f16_m256(128);
  constexpr int kTmemCols = 256;                                          // 2 stages x 128 columns

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int sa = x.sa;
  const uint32_t smem_w = smem_base;
  const uint32_t smem_a = smem_w + kR3WBytes;
  const uint32_t smem_out = smem_a + (uint32_t)sa * kR3SlabBytes;          // POOL: 2 x 8 KB pooled, else 2 x 16 KB full
  const uint32_t bar_base = smem_out + (POOL ? 2u * kR3PoolBytes : 2u * kStageOutBytes);
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (sa + s); };
  const uint32_t misc = bar_base + 8u * (2 * sa);
  auto tfull_bar = [&](int a) { return misc + 8u * a; };
  auto tempty_bar = [&](int a) { return misc + 16u + 8u * a; };
  const uint32_t wbar = misc + 32u;
  const uint32_t tmem_ptr_smem = misc + 40u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  const uint32_t smem_bias = misc + 64u;
  float* smem_bias_gen = reinterpret_cast<float*>(smem_gen + (smem_bias - smem_base));
  stage_bias(p, smem_bias_gen, 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (POOL) {
      tma_prefetch_desc(&tmP);
    } else {
      tma_prefetch_desc(&tmC0);
      tma_prefetch_desc(&tmC1);
    }
    for (int s = 0; s < sa; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * 4 * kR3EpiSplit); }
    mbar_init(wbar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are counted on the leader's barriers) =====================
    if (elect_one()) {
      if (rank == 0) mbar_expect_tx(wbar, 2u * kR3WBytes);
      const uint32_t wbar_l = mapa_cluster(wbar, 0);
      const int rk = (int)rank;
      // (the weight tensor map's box is 32 output channels x 64 input channels of one tap)
      auto half = [&](uint32_t off, int tap, int co0) { tma_load_3d_2sm(smem_w + off, &tmB, wbar_l, 0, co0, tap); };
      auto full = [&](uint32_t off, int tap) { half(off, tap, 0); half(off + kR3Half, tap, 32); };
      // N = 128 MMAs: rank 0 holds W[r = t, s] (output row 2j), rank 1 W[r = t - 1, s] (output row 2j + 1)
      full(r3_full(1, 1), (1 - rk) * 3 + 1);
      full(r3_full(1, 2), (1 - rk) * 3 + 2);
#pragma unroll
      for (int s = 0; s < 3; ++s) full(r3_full(2, s), (2 - rk) * 3 + s);
      // N = 64 MMAs: each rank its 32 output channels
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        half(r3_h0(s), 0 * 3 + s, 32 * rk);
        half(r3_h3(s), 2 * 3 + s, 32 * rk);
      }
      half(kR3H1, 1 * 3 + 0, 32 * rk);
    }
    __syncwarp();
    int as_ = 0;
    uint32_t aph = 0;
    for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
      int n, txp, ty, b;
      decode_tile(p, tile, n, txp, ty, b);
      const int x0 = (txp * 2 + (int)rank) * kR3Tw, y0 = ty * (2 * kR3J);
      mbar_wait(aempty(as_), aph ^ 1u);
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(afull(as_), (uint32_t)(2 * kR3SlabTx));
        tma_load_4d_2sm(smem_a + as_ * kR3SlabBytes, &tmA, mapa_cluster(afull(as_), 0), 0, x0 - 1, y0 - 1, b);
      }
      if (++as_ == sa) { as_ = 0; aph ^= 1u; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp, one elected lane) =====================
    if (rank == 0) {
      mbar_wait(wbar, 0);
      int as_ = 0;
      uint32_t aph = 0;
      int acc = 0;
      uint32_t accph = 0;
      for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
        mbar_wait(tempty_bar(acc), accph ^ 1u);
        tc_fence_after();
        mbar_wait(afull(as_), aph);
        tc_fence_after();
        const uint32_t slab = smem_a + as_ * kR3SlabBytes;
        const uint32_t d0 = tmem_base + (uint32_t)(acc * 128), d1 = d0 + 64u;
        if (elect_one()) {
          // accumulator row (x, j) <- slab pixel (row 2j + t, column x + s): 8-row groups two image rows apart
          auto adesc = [&](int t, int s) {
            return r3_desc_k_sw128(slab + (uint32_t)(t * kR3Pitch) + (uint32_t)s * 128u, 2u * kR3Pitch);
          };
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            {   // t = 0: W[r=0,s] -> output row 2j
              const uint64_t a = adesc(0, s), b = umma_desc_k_sw128(smem_w + r3_h0(s));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_2sm(d0, a + 2u * k, b + 2u * k, kI64, (s | k) != 0 ? 1u : 0u);
            }
            if (s == 0) {
              // t = 1, first column tap: the first contribution to output row 2j + 1 must OVERWRITE its columns, so the
              // two halves go as separate N = 64 MMAs (W[r=1,0] -> row 2j accumulating, W[r=0,0] -> row 2j + 1 fresh)
              const uint64_t a = adesc(1, 0), b1 = umma_desc_k_sw128(smem_w + kR3H1), b0 = umma_desc_k_sw128(smem_w + r3_h0(0));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_2sm(d0, a + 2u * k, b1 + 2u * k, kI64, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_2sm(d1, a + 2u * k, b0 + 2u * k, kI64, k != 0 ? 1u : 0u);
            } else {
              const uint64_t a = adesc(1, s), b = umma_desc_k_sw128(smem_w + r3_full(1, s));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_2sm(d0, a + 2u * k, b + 2u * k, kI128, 1u);
            }
            {   // t = 2: [W[r=2,s] | W[r=1,s]]
              const uint64_t a = adesc(2, s), b = umma_desc_k_sw128(smem_w + r3_full(2, s));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_2sm(d0, a + 2u * k, b + 2u * k, kI128, 1u);
            }
            {   // t = 3: W[r=2,s] -> output row 2j + 1
              const uint64_t a = adesc(3, s), b = umma_desc_k_sw128(smem_w + r3_h3(s));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_2sm(d1, a + 2u * k, b + 2u * k, kI64, 1u);
            }
          }
          umma_commit_2sm(aempty(as_));
          umma_commit_2sm(tfull_bar(acc));
        }
        if (++as_ == sa) { as_ = 0; aph ^= 1u; }
        acc ^= 1;
        if (acc == 0) accph ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs drain their own TMEM half) =====================
    const int q = warp & 3;
    const int hsel = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const int j = row >> 3, lx = row & 7;
    int acc = 0;
    uint32_t accph = 0;
    uint32_t chunk_ctr = 0;
    float csum[kCsumSize<64, kR3EpiSplit>];
#pragma unroll
    for (int i = 0; i < kCsumSize<64, kR3EpiSplit>; ++i) csum[i] = 0.0f;
    float breg[32];
    const bool bias_regs = !POOL && p.bias != nullptr;
#pragma unroll
    for (int i = 0; i < 32; ++i) breg[i] = bias_regs ? __ldg(p.bias + hsel * 32 + i) : 0.0f;
    const uint32_t tempty_l0 = mapa_cluster(tempty_bar(0), 0), tempty_l1 = mapa_cluster(tempty_bar(1), 0);
    for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
      int n, txp, ty, b;
      decode_tile(p, tile, n, txp, ty, b);
      const int tx = txp * 2 + (int)rank;                            // 8-pixel column of this CTA
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128);
      const uint32_t tempty_l = acc ? tempty_l1 : tempty_l0;
      if constexpr (POOL) {
        mbar_wait(tfull_bar(acc), accph);
        tc_fence_after();
        // pool first (see conv_common.cuh): vertical max = the two halves of the accumulator row, horizontal max with
        // lane ^ 1; the two lanes of a window then own 16 pooled channels each
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(t_row + (uint32_t)(hsel * 32), v0);
        tmem_ld_32x32b_x32(t_row + 64u + (uint32_t)(hsel * 32), v1);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_l);
        const uint32_t odd = (uint32_t)lane & 1u;
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float lo = fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i]));
          const float hi = fmaxf(__uint_as_float(v0[16 + i]), __uint_as_float(v1[16 + i]));
          const float send = odd ? lo : hi, keep = odd ? hi : lo;
          a[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
        }
        if (p.bias != nullptr) {
          const uint32_t bias_addr = smem_bias + (uint32_t)(hsel * 32 + (int)(odd * 16u)) * 4u;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(bias_addr + (uint32_t)i * 4u));
            a[i] += b0; a[i + 1] += b1; a[i + 2] += b2; a[i + 3] += b3;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = fmaxf(a[i], 0.0f);
        }
        uint32_t packed[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) packed[i] = pack_h2(a[2 * i], a[2 * i + 1]);
        const uint32_t pbuf = smem_out + (chunk_ctr & 1u) * kR3PoolBytes;
        ++chunk_ctr;
        if (epi_tid < 32) {                              // the store that used pbuf two tiles ago has read it
          if (elect_one()) tma_store_wait_read<1>();
        }
        named_bar_sync(1, 128 * kR3EpiSplit);
        {
          const uint32_t pr = (uint32_t)(j * 4 + (lx >> 1));             // pooled pixel of the 4 x 16 tile
          const uint32_t c16 = (uint32_t)hsel * 4u + odd * 2u;           // first of this lane's two 16-byte pieces
#pragma unroll
          for (uint32_t h = 0; h < 2; ++h) {
            const uint32_t dst = pbuf + pr * 128u + (((c16 + h) ^ (pr & 7u)) * 16u);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(packed[4 * h]), "r"(packed[4 * h + 1]),
                         "r"(packed[4 * h + 2]), "r"(packed[4 * h + 3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128 * kR3EpiSplit);
        if (epi_tid < 32) {
          if (elect_one()) {
            tma_store_4d(&tmP, pbuf, 0, tx * (kR3Tw / 2), ty * kR3J, b);
            tma_store_commit();
          }
        }
      } else {
        // un-pooled: the two column halves are two ordinary 8 x 16 pixel tiles of the even / odd output rows
        const int ox = tx * kR3Tw + lx;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int oy = ty * (2 * kR3J) + 2 * j + c;
          const bool valid = (ox < p.Wo) && (oy < p.Ho);
          epilogue_nhwc_tile<64, kR3EpiSplit, true, true>(p, c == 0 ? &tmC0 : &tmC1, &tmP, t_row + (uint32_t)(c * 64), smem_out,
                                                          0u, smem_bias, smem_bias_gen, tempty_l, 0, tx, ty, b, ox, oy,
                                                          valid, row, lane, epi_tid, chunk_ctr, hsel, csum,
                                                          bias_regs ? breg : nullptr, nullptr, c == 0 ? tfull_bar(acc) : 0u,
                                                          accph, nullptr, c == 1);
        }
      }
      acc ^= 1;
      if (acc == 0) accph ^= 1u;
    }
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

template <bool POOL>
static int launch_rs3(const dreamb200_conv_desc* d, cudaStream_t stream) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.tw = kR3Tw; p.th = kR3J;                                             // (geometry of one column half, for the epilogue)
  p.tiles_x = (d->Wo + 2 * kR3Tw - 1) / (2 * kR3Tw);                     // PAIRS of 8-pixel columns
  p.tiles_y = (d->Ho + 2 * kR3J - 1) / (2 * kR3J);
  p.n_tiles = 1;
  p.B = d->B; p.Ho = d->Ho; p.Wo = d->Wo;
  p.total_tiles = p.tiles_x * p.tiles_y * d->B;
  DB_REQUIRE((long long)p.tiles_x * p.tiles_y * d->B < (1ll << 24) && p.tiles_x < 65536 && p.tiles_y < 65536,
             "conv: too many tiles for one launch (%d x %d x %d)", p.tiles_x, p.tiles_y, d->B);
  p.mg_n = div_magic(1);
  p.mg_x = div_magic(p.tiles_x);
  p.mg_y = div_magic(p.tiles_y);
  p.in_stride = 1;
  p.taps = 9;
  p.kchunks = 1;
  p.bias = d->bias;
  p.Cout_pad = 64;
  p.relu = d->relu;
  p.pool = 0;                       // (the pooled variant has its own epilogue; the generic one sees plain tiles)
  p.store_full = 1;

  Rs3Extra x;
  const int out_bytes = POOL ? 2 * kR3PoolBytes : 2 * kStageOutBytes;
  x.sa = (232448 - 1024 - (int)kR3WBytes - out_bytes - 1024) / kR3SlabBytes;
  if (x.sa > 4) x.sa = 4;
  DB_REQUIRE(x.sa >= 2, "conv_rs3: shared memory budget too small");
  const int smem_bytes = 1024 + (int)kR3WBytes + x.sa * kR3SlabBytes + out_bytes + 1024;

  CUtensorMap tmA, tmB, tmC0, tmC1, tmP;
  memset(&tmC0, 0, sizeof(tmC0));
  memset(&tmC1, 0, sizeof(tmC1));
  memset(&tmP, 0, sizeof(tmP));
  const uint32_t es4[4] = {1, 1, 1, 1};
  {
    uint64_t dims[4] = {64, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t str[3] = {128, (uint64_t)d->W * 128, (uint64_t)d->H * d->W * 128};
    uint32_t box[4] = {64, 10u, (uint32_t)kR3Rows, 1};
    if (make_tensor_map_f16(&tmA, d->x, 4, dims, str, box, es4, "rs3 activation")) return -1;
  }
  {
    uint64_t dims[3] = {64, 64, 9};
    uint64_t str[2] = {128, 64 * 128};
    uint32_t box[3] = {64, 32, 1};
    uint32_t es[3] = {1, 1, 1};
    if (make_tensor_map_f16(&tmB, d->w, 3, dims, str, box, es, "rs3 weights")) return -1;
  }
  if (POOL) {
    const uint64_t Wp = (uint64_t)(d->Wo / 2), Hp = (uint64_t)(d->Ho / 2);
    uint64_t dims[4] = {64, Wp, Hp, (uint64_t)d->B};
    uint64_t str[3] = {128, Wp * 128, Hp * Wp * 128};
    uint32_t box[4] = {64, kR3Tw / 2, kR3J, 1};
    if (make_tensor_map_f16(&tmP, d->y_pool, 4, dims, str, box, es4, "rs3 pooled output")) return -1;
  } else {
    // even / odd output rows as two views of the output tensor
    for (int c = 0; c < 2; ++c) {
      const uint64_t rows = (uint64_t)((d->Ho + 1 - c) / 2);
      if (rows == 0) { memcpy(&tmC1, &tmC0, sizeof(tmC0)); continue; }    // one-row map: the odd view is never stored in range
      uint64_t dims[4] = {64, (uint64_t)d->Wo, rows, (uint64_t)d->B};
      uint64_t str[3] = {(uint64_t)d->y_stride_w * 2, (uint64_t)d->y_stride_h * 4, (uint64_t)d->y_stride_b * 2};
      uint32_t box[4] = {64, kR3Tw, kR3J, 1};
      const __half* base = reinterpret_cast<const __half*>(d->y) + (size_t)c * d->y_stride_h;
      if (make_tensor_map_f16(c == 0 ? &tmC0 : &tmC1, base, 4, dims, str, box, es4, "rs3 output")) return -1;
    }
  }
  auto kern = conv_rs3_kernel<POOL>;
  static bool attr_set = false;
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int sms = device_sm_count() & ~1;
  int grid = 2 * p.total_tiles < sms ? 2 * p.total_tiles : sms;
  kern<<<grid, kR3Threads, smem_bytes, stream>>>(tmA, tmB, tmC0, tmC1, tmP, p, x);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// Returns 1 and launches when the layer qualifies for the two-row kernel, 0 otherwise, <0 on error.
int try_conv_rs3(const dreamb200_conv_desc* d, cudaStream_t stream) {
  const char* e = getenv("DREAMB200_RS3");            // 0 = off (conv_rs2 takes the layer), bit 0: pooled, bit 1: un-pooled
  const int mode = e ? atoi(e) : 3;                   // (read per call: the A/B tests flip it)
  if (mode == 0) return 0;
  if (d->out_mode != DREAMB200_OUT_NHWC_F16 || d->taps != 9 || d->in_stride != 1) return 0;
  if (d->Ho != d->H || d->Wo != d->W) return 0;
  if (d->Cout_pad != 64 || d->Cin != 64) return 0;
  if (d->Ho < 2 * kR3J) return 0;
  for (int t = 0; t < 9; ++t)
    if (d->tap_dy[t] != t / 3 - 1 || d->tap_dx[t] != t % 3 - 1) return 0;
  if (d->residual != nullptr || d->residual_f32 != nullptr || d->y_f32 != nullptr || d->gate != nullptr ||
      d->out_scale != nullptr || d->colsum != nullptr || d->absmax != nullptr)
    return 0;
  const bool pool = d->y_pool != nullptr;
  if (pool && d->y != nullptr) return 0;              // (training forward: both tensors -> conv_rs2)
  if (!pool && d->y == nullptr) return 0;
  if (pool ? !(mode & 1) : !(mode & 2)) return 0;
  if (!pool && ((d->y_stride_h * 4) % 16 != 0)) return 0;
  const double util = (double)d->Wo * d->Ho /
                      ((double)((d->Wo + 2 * kR3Tw - 1) / (2 * kR3Tw)) * ((d->Ho + 2 * kR3J - 1) / (2 * kR3J)) * 512.0);
  if (util < 0.65) return 0;
  const int rc = pool ? launch_rs3<true>(d, stream) : launch_rs3<false>(d, stream);
  return rc == 0 ? 1 : rc;
}

}  // namespace db200
