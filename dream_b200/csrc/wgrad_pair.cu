// wgrad_pair.cu -- CTA-pair (tcgen05 cta_group::2) variant of the row-shared 3x3 weight-gradient kernel
// (wgrad3x3_kernel<128> in train_kernels.cu) for layers with Cout % 256 == 0 and Cin % 128 == 0.
// On by default since round 2 (tools/wgrad_pair_check.py, tests/test_gpu_kernel_variants.py).
//
// Why: dW[tap][co][ci] = sum_px dY[px][co] * X[px + tap][ci] runs as M = 128 (co) x N = 128 (ci) MMAs with K = 16 pixels,
// both operands MN-major from shared memory: 32 + 32 operand wavefronts per 64-cycle MMA -- exactly the shared-memory
// bandwidth, plus the TMA writes of the same tiles (ncu: tensor pipe 79 % active, data pipe 77 %).  A CTA pair owns TWO
// co tiles (M = 256: rank r supplies the dY tile of co tile 2*cp + r) and splits the ci tile (N = 128: rank r supplies
// and loads only the X slab of its 64 input channels): 32 + 16 wavefronts per MMA and 52 KB instead of 72 KB per stage.
// Each CTA's TMEM half holds its co tile x all 128 ci for the three taps of the kernel row; the epilogue is unchanged.
// Pair protocol as in conv_rs2.cu (bytes counted on the leader's full barrier, multicast commits); the accumulators
// are drained once, after the split's last k-block, so there is no accumulator hand-back.
//
// Tile height.  A k-block is an 8-pixel-wide (one swizzle atom per image row), `th`-row tile of the map; pixels
// outside the image are TMA zero fill, i.e. MMAs spent on zeros.  With the fixed 16 rows of round 1 a 50x50 map was cut
// into 7 x 4 tiles = 3584 pixel slots for 2500 pixels (70 % useful, 1.19 PFLOP/s on the 512-channel layers against 1.40 at
// 100x100).  `th` is now chosen per launch (any even number up to 26; it only enters the TMA boxes, the stage layout
// and the MMA loop bound): 10 rows for 50x50 (89 %), 20 for 100x100 (96 %), 26 for 25x25 (75 %, was 61 %).
#include "common.cuh"
#include "dreamb200.h"

#include <stdlib.h>

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int device_sm_count();

struct WgradPairParams {
  int B, H, W;
  int tiles_x, tiles_y;
  int co_pairs, ci_tiles, splits;
  long long kblocks_total;
  float* dw;
  int Cout_pad, Cin_pad;
  int stages;
  int th;            // image rows per k-block (even)
};

constexpr int kWpThreads = 192;
constexpr int kWpN = 128;
// one pipeline stage for `th` image rows: dY = 2 chunks of [8 * th px][64 co] (this CTA's co tile), X = a slab of th
// image rows x 10 pixels x 64 ci (this CTA's half of the ci tile), padded to the 1024-byte swizzle period
__host__ __device__ constexpr uint32_t wp_chunk_bytes(int th) { return (uint32_t)th * 1024u; }
__host__ __device__ constexpr uint32_t wp_slab_tx(int th) { return (uint32_t)th * 1280u; }
__host__ __device__ constexpr uint32_t wp_slab_bytes(int th) { return (wp_slab_tx(th) + 1023u) & ~1023u; }
__host__ __device__ constexpr uint32_t wp_stage_bytes(int th) { return 2u * wp_chunk_bytes(th) + wp_slab_bytes(th); }

__host__ __device__ constexpr uint32_t umma_idesc_f16_m256_mn(uint32_t n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWpThreads, 1)
wgrad3x3_pair_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                     const __grid_constant__ WgradPairParams p) {
  constexpr uint32_t kIdesc = umma_idesc_f16_m256_mn(kWpN);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t kWpChunk = wp_chunk_bytes(p.th), kWpABytes = 2u * kWpChunk, kWpSlab = wp_slab_bytes(p.th);
  const uint32_t kWpStageBytes = wp_stage_bytes(p.th);
  const uint32_t bar_base = smem_base + stages * kWpStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * stages);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * stages + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // unit of this PAIR: (split, kernel row r, co pair, ci tile)
  int u = blockIdx.x >> 1;
  const int ci_t = u % p.ci_tiles; u /= p.ci_tiles;
  const int co_p = u % p.co_pairs; u /= p.co_pairs;
  const int r = u % 3;
  const int split = u / 3;
  const int co_t = co_p * 2 + (int)rank;
  const long long kb_lo = p.kblocks_total * split / p.splits;
  const long long kb_hi = p.kblocks_total * (split + 1) / p.splits;
  const int n_kb = (int)(kb_hi - kb_lo);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are counted on the leader's barriers) =====================
    int stage = 0;
    uint32_t phase = 0;
    const int tiles = p.tiles_x * p.tiles_y;
    for (long long kb = kb_lo; kb < kb_hi; ++kb) {
      const int b = (int)(kb / tiles);
      const int rr = (int)(kb - (long long)b * tiles);
      const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
      const int x0 = tx * 8, y0 = ty * p.th;
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t sa = smem_base + stage * kWpStageBytes;
      if (elect_one()) {
        if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * (kWpABytes + wp_slab_tx(p.th)));
        const uint32_t bar = mapa_cluster(full_bar(stage), 0);
#pragma unroll
        for (int m = 0; m < 2; ++m)
          tma_load_4d_2sm(sa + m * kWpChunk, &tmDY, bar, co_t * 128 + m * 64, x0, y0, b);
        tma_load_4d_2sm(sa + kWpABytes, &tmX, bar, ci_t * kWpN + (int)rank * 64, x0 - 1, y0 + r - 1, b);
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp, one elected lane) =====================
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kWpStageBytes;
        const uint32_t sb = sa + kWpABytes;
        if (elect_one()) {
          const int n_j = p.th >> 1;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
#pragma unroll 4
            for (int j = 0; j < n_j; ++j) {     // 16 pixels = 2 image rows of the tile per MMA
              const uint64_t adesc = umma_desc_mn_sw128(sa + (uint32_t)j * 2048u, kWpChunk);
              const uint64_t bdesc =
                  umma_desc_mn_sw128_sbo(sb + (uint32_t)(2 * j) * 1280u + (uint32_t)s * 128u, kWpSlab, 1280u);
              umma_f16_2sm(tmem_base + (uint32_t)(s * kWpN), adesc, bdesc, kIdesc, (kb | j) != 0 ? 1u : 0u);
            }
          }
          umma_commit_2sm(empty_bar(stage));
          if (kb == n_kb - 1) umma_commit_2sm(done_bar);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (n_kb > 0) {
    // ===================== epilogue: this CTA's co tile x 128 ci x 3 taps -> fp32 atomics =====================
    const int q = warp & 3;
    const int co = co_t * 128 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
      float* out = p.dw + ((size_t)(r * 3 + s) * p.Cout_pad + co) * p.Cin_pad + ci_t * kWpN;
#pragma unroll 1
      for (int c = 0; c < kWpN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kWpN + c * 32), v);
        tmem_wait_ld();
        if (co < p.Cout_pad) {
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(out + c * 32 + i, __uint_as_float(v[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// Returns 1 and launches when the layer qualifies (and the kernel is switched on), 0 otherwise, <0 on error.
int try_wgrad3x3_pair(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout_pad, int Cin_pad,
                      cudaStream_t stream) {
  // on by default since round 2 (tools/wgrad_pair_check.py green on a B200, 7-14 % faster on the 256 / 512-channel
  // layers: profiles/r02_ab_pair_kernels.txt); DREAMB200_WGRAD3_2SM=0 switches back.  Read per call.
  const char* e = getenv("DREAMB200_WGRAD3_2SM");
  if ((e && e[0] == '0') || Cout_pad % 256 != 0 || Cin_pad % 128 != 0) return 0;
  WgradPairParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  // rows per k-block: the even height <= 26 that wastes the fewest rows of the map (ties: the taller tile, fewer
  // k-blocks), while at least two stages fit; DREAMB200_WGRAD_TH pins it (A/B; 16 = round 1)
  int th = 16;
  {
    const char* e_th = getenv("DREAMB200_WGRAD_TH");
    int pinned = e_th ? atoi(e_th) : 0;
    if (pinned >= 2 && pinned <= 26 && pinned % 2 == 0) {
      th = pinned;
    } else {
      long long best = -1;
      for (int t = 4; t <= 26; t += 2) {
        const long long rows = (long long)((H + t - 1) / t) * t;
        if (best < 0 || rows < best || (rows == best && t > th)) { best = rows; th = t; }
      }
    }
  }
  p.th = th;
  p.tiles_x = (W + 7) / 8;
  p.tiles_y = (H + th - 1) / th;
  p.co_pairs = Cout_pad / 256;
  p.ci_tiles = Cin_pad / kWpN;
  p.kblocks_total = (long long)B * p.tiles_x * p.tiles_y;
  const int units = 3 * p.co_pairs * p.ci_tiles;            // pairs per split
  int splits = (device_sm_count() / 2) / units;             // the whole grid must be ONE wave (see wgrad3x3_impl)
  if (splits < 1) splits = 1;
  if ((long long)splits > p.kblocks_total) splits = (int)p.kblocks_total;
  p.splits = splits;
  p.dw = dw;
  p.Cout_pad = Cout_pad;
  p.Cin_pad = Cin_pad;
  CUtensorMap tmDY, tmX;
  const uint32_t es[4] = {1, 1, 1, 1};
  {
    const uint32_t box[4] = {64, 8, (uint32_t)th, 1};
    uint64_t dims[4] = {(uint64_t)Cout_pad, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cout_pad * 2, (uint64_t)W * Cout_pad * 2, (uint64_t)H * W * Cout_pad * 2};
    if (make_tensor_map_f16(&tmDY, dy, 4, dims, str, box, es, "wgrad pair dY")) return -1;
  }
  {
    const uint32_t box[4] = {64, 10, (uint32_t)th, 1};
    uint64_t dims[4] = {(uint64_t)Cin_pad, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin_pad * 2, (uint64_t)W * Cin_pad * 2, (uint64_t)H * W * Cin_pad * 2};
    if (make_tensor_map_f16(&tmX, x, 4, dims, str, box, es, "wgrad pair X")) return -1;
  }
  int stages = (232448 - 1024 - 512) / (int)wp_stage_bytes(th);
  if (stages > 6) stages = 6;
  p.stages = stages;
  const int smem_bytes = 1024 + stages * (int)wp_stage_bytes(th) + 512;
  static bool attr_set = false;
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(wgrad3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int grid = 2 * p.splits * units;
  wgrad3x3_pair_kernel<<<grid, kWpThreads, smem_bytes, stream>>>(tmDY, tmX, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 1;
}

}  // namespace db200
