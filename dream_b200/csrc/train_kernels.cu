// train_kernels.cu -- backward pass of the DREAM networks (reference: `loss.backward()` in
// dream/network.py:328-338, i.e. autograd of nn.Conv2d / MaxPool2d / Upsample / ReLU in
// dream/models.py:591-747).
//
//   data gradient  : dX = conv(dY, W rotated 180 deg, Cin<->Cout swapped) -> the forward tensor-core kernel
//                    (conv_tc.cu) with re-packed weights; nothing new here.
//   weight gradient: dW[tap][co][ci] = sum_p dY[p][co] * X[p + tap][ci]   (this file, wgrad_tc_kernel):
//                    a GEMM whose reduction axis is the pixel index.  Both operands are staged as
//                    K-major tiles by TMA from channel-major (NCHW, row pitch padded to 8) fp16 copies:
//                    a box of (bx x by) = 64 pixels x 128|N channels lands as rows of 128 B per channel,
//                    the same canonical SWIZZLE_128B layout the forward kernel feeds to tcgen05.mma.  The
//                    tap shift and the zero padding are TMA coordinates / out-of-bounds fill.  Split-K over
//                    pixel ranges across CTAs, fp32 accumulation in TMEM, fp32 atomics into dW.
//   plus the HBM-bound pieces: ReLU mask, 2x2 max-pool backward, nearest-upsample backward, bias
//   gradient (column sums) and the NHWC -> channel-major transposes.
#include "common.cuh"
#include "dreamb200.h"

#include <stdlib.h>

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int device_sm_count();

// ---------------------------------------------------------------------------------------------
// wgrad on tcgen05
// ---------------------------------------------------------------------------------------------
// k-block = one 16x8 pixel tile of one image (128 pixels = 8 MMAs of K=16).  Both operands are MN-major:
//   A (dY): 2 boxes [128 px][64 co]  -> M = 128 output channels (a half-empty tile is TMA zero fill)
//   B (X) : BLOCK_N/64 boxes [128 px][64 ci] fetched at the tap-shifted coordinates -> N = BLOCK_N
// (a channel-major staging was tried first: a one-pixel x shift would need a 2-byte-granular TMA start
//  address, which TMA does not allow -- shifts must be along non-contiguous dimensions, as they are in NHWC).
struct WgradParams {
  int B, H, W;
  int tiles_x, tiles_y;
  int taps;
  int8_t dy[DREAMB200_MAX_TAPS], dx[DREAMB200_MAX_TAPS];
  int co_tiles, ci_tiles;
  int splits;
  long long kblocks_total;  // B * tiles_y * tiles_x
  float* dw;                // fp32 [taps][Cout_pad][Cin_pad], accumulated with atomics
  int Cout_pad, Cin_pad;
  int stages;
  int a_stride;             // 1: conv (dY on X's grid); 2: ConvTranspose stride 2 (dY on the 2x finer grid)
  int shift_a;              // tap offset applied to the dY coordinates (ConvTranspose) instead of X's
  int b_stride;             // 2: stride-2 convolution (X on the 2x finer grid, tiles run over dY's grid)
};

constexpr int kWgThreads = 192;
constexpr int kWgTw = 16, kWgTh = 8;
constexpr int kChunkBytes = 128 * 128;   // [128 px][64 ch] fp16

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ WgradParams p) {
  constexpr int kABytes = 2 * kChunkBytes;
  constexpr int kBBytes = (BLOCK_N / 64) * kChunkBytes;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  constexpr uint32_t kIdesc = umma_idesc_f16_m128_mn(BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t bar_base = smem_base + stages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * stages);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * stages + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // work unit of this CTA: (split, tap, co tile, ci tile); CTAs with the same split are adjacent
  int u = blockIdx.x;
  const int ci_t = u % p.ci_tiles; u /= p.ci_tiles;
  const int co_t = u % p.co_tiles; u /= p.co_tiles;
  const int tap = u % p.taps;
  const int split = u / p.taps;
  const long long kb_lo = p.kblocks_total * split / p.splits;
  const long long kb_hi = p.kblocks_total * (split + 1) / p.splits;
  const int n_kb = (int)(kb_hi - kb_lo);

  if (warp == 0) {
    {   // whole warp, uniform control flow; one elected lane issues (see elect_one)
      int stage = 0;
      uint32_t phase = 0;
      const int tiles = p.tiles_x * p.tiles_y;
      for (long long kb = kb_lo; kb < kb_hi; ++kb) {
        const int b = (int)(kb / tiles);
        const int r = (int)(kb - (long long)b * tiles);
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int x0 = tx * kWgTw, y0 = ty * kWgTh;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * kStageBytes;
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), (uint32_t)(kABytes + kBBytes));
          const int ax = x0 * p.a_stride + (p.shift_a ? p.dx[tap] : 0), ay = y0 * p.a_stride + (p.shift_a ? p.dy[tap] : 0);
          const int bx = x0 * p.b_stride + (p.shift_a ? 0 : p.dx[tap]), by = y0 * p.b_stride + (p.shift_a ? 0 : p.dy[tap]);
#pragma unroll
          for (int m = 0; m < 2; ++m)
            tma_load_4d(sa + m * kChunkBytes, &tmDY, full_bar(stage), co_t * 128 + m * 64, ax, ay, b);
#pragma unroll
          for (int n = 0; n < BLOCK_N / 64; ++n)
            tma_load_4d(sa + kABytes + n * kChunkBytes, &tmX, full_bar(stage), ci_t * BLOCK_N + n * 64, bx, by, b);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp, uniform control flow; one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kStageBytes;
        const uint64_t adesc = umma_desc_mn_sw128(sa, kChunkBytes);
        const uint64_t bdesc = umma_desc_mn_sw128(sa + kABytes, kChunkBytes);
        if (elect_one()) {            // ONE election per k-block: the body is straight-line UTCHMMA + commits
#pragma unroll
          for (int k = 0; k < 8; ++k)   // 16 pixel rows (2 KB) per MMA
            umma_f16(tmem_base, adesc + 128u * k, bdesc + 128u * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (kb == n_kb - 1) umma_commit(done_bar);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (n_kb == 0 && elect_one()) umma_commit(done_bar);
    }
    __syncwarp();
  } else if (n_kb > 0) {
    // epilogue: 128 accumulator rows (co) x BLOCK_N columns (ci) -> fp32 atomics
    const int q = warp & 3;
    const int co = co_t * 128 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* out = p.dw + ((size_t)tap * p.Cout_pad + co) * p.Cin_pad + ci_t * BLOCK_N;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      tmem_wait_ld();
      if (co < p.Cout_pad) {
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(out + c * 32 + i, __uint_as_float(v[i]));
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

template <int BLOCK_N>
static int launch_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, WgradParams& p, cudaStream_t stream) {
  constexpr int kStageBytes = (2 + BLOCK_N / 64) * kChunkBytes;
  int stages = (232448 - 1024 - 512) / kStageBytes;
  if (stages > 6) stages = 6;
  p.stages = stages;
  const int smem_bytes = 1024 + stages * kStageBytes + 512;
  auto kern = wgrad_tc_kernel<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int grid = p.splits * p.taps * p.co_tiles * p.ci_tiles;
  kern<<<grid, kWgThreads, smem_bytes, stream>>>(tmDY, tmX, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// wgrad for the standard 3x3 / stride 1 / pad 1 convolution: one CTA owns a kernel ROW r (its three taps s).
// Per 8x16-pixel k-block it loads the dY tile once (it is the same for every tap) and ONE 10x16-pixel slab of X
// per 64-channel chunk; the B operand of tap (r, s) is that slab read from byte offset s*128 with 1280 B between
// image rows (the hardware swizzles on the absolute shared-memory address, see conv_rs.cu).  Three accumulators
// (s = 0,1,2) of BLOCK_N columns live in TMEM.  2.5x fewer bytes per MMA than the generic per-tap kernel above.
// ---------------------------------------------------------------------------------------------
struct Wgrad3Params {
  int B, H, W;
  int tiles_x, tiles_y;
  int co_tiles, ci_tiles, splits;
  long long kblocks_total;
  float* dw;
  int Cout_pad, Cin_pad;
  int stages;
};

constexpr int kW3SlabBytes = 16 * 1280;      // 16 image rows x 10 pixels x 128 B

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                const __grid_constant__ Wgrad3Params p) {
  constexpr int kABytes = 2 * kChunkBytes;
  constexpr int kBBytes = (BLOCK_N / 64) * kW3SlabBytes;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = 3 * BLOCK_N <= 256 ? 256 : 512;
  constexpr uint32_t kIdesc = umma_idesc_f16_m128_mn(BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t bar_base = smem_base + stages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * stages);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * stages + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // unit of this CTA: (split, kernel row r, co tile, ci tile); CTAs of one split are adjacent (shared L2 lines)
  int u = blockIdx.x;
  const int ci_t = u % p.ci_tiles; u /= p.ci_tiles;
  const int co_t = u % p.co_tiles; u /= p.co_tiles;
  const int r = u % 3;
  const int split = u / 3;
  const long long kb_lo = p.kblocks_total * split / p.splits;
  const long long kb_hi = p.kblocks_total * (split + 1) / p.splits;
  const int n_kb = (int)(kb_hi - kb_lo);

  if (warp == 0) {
    {   // whole warp, uniform control flow; one elected lane issues (see elect_one)
      int stage = 0;
      uint32_t phase = 0;
      const int tiles = p.tiles_x * p.tiles_y;
      for (long long kb = kb_lo; kb < kb_hi; ++kb) {
        const int b = (int)(kb / tiles);
        const int rr = (int)(kb - (long long)b * tiles);
        const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
        const int x0 = tx * 8, y0 = ty * 16;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * kStageBytes;
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), (uint32_t)(kABytes + kBBytes));
#pragma unroll
          for (int m = 0; m < 2; ++m)
            tma_load_4d(sa + m * kChunkBytes, &tmDY, full_bar(stage), co_t * 128 + m * 64, x0, y0, b);
#pragma unroll
          for (int n = 0; n < BLOCK_N / 64; ++n)
            tma_load_4d(sa + kABytes + n * kW3SlabBytes, &tmX, full_bar(stage), ci_t * BLOCK_N + n * 64, x0 - 1,
                        y0 + r - 1, b);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp, uniform control flow; one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kStageBytes;
        const uint32_t sb = sa + kABytes;
        if (elect_one()) {            // ONE election per k-block: the body is straight-line UTCHMMA + commits
#pragma unroll
          for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {       // 16 pixels = 2 image rows of the tile per MMA
              const uint64_t adesc = umma_desc_mn_sw128(sa + (uint32_t)j * 2048u, kChunkBytes);
              const uint64_t bdesc =
                  umma_desc_mn_sw128_sbo(sb + (uint32_t)(2 * j) * 1280u + (uint32_t)s * 128u, kW3SlabBytes, 1280u);
              umma_f16(tmem_base + (uint32_t)(s * BLOCK_N), adesc, bdesc, kIdesc, (kb | j) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (kb == n_kb - 1) umma_commit(done_bar);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (n_kb == 0 && elect_one()) umma_commit(done_bar);
    }
    __syncwarp();
  } else if (n_kb > 0) {
    const int q = warp & 3;
    const int co = co_t * 128 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int s = 0; s < 3; ++s) {
      float* out = p.dw + ((size_t)(r * 3 + s) * p.Cout_pad + co) * p.Cin_pad + ci_t * BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * BLOCK_N + c * 32), v);
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int BLOCK_N>
static int launch_wgrad3(const CUtensorMap& tmDY, const CUtensorMap& tmX, Wgrad3Params& p, cudaStream_t stream) {
  constexpr int kStageBytes = 2 * kChunkBytes + (BLOCK_N / 64) * kW3SlabBytes;
  int stages = (232448 - 1024 - 512) / kStageBytes;
  if (stages > 6) stages = 6;
  p.stages = stages;
  const int smem_bytes = 1024 + stages * kStageBytes + 512;
  auto kern = wgrad3x3_kernel<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int grid = p.splits * 3 * p.co_tiles * p.ci_tiles;
  kern<<<grid, kWgThreads, smem_bytes, stream>>>(tmDY, tmX, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static int wgrad3x3_impl(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout_pad, int Cin_pad,
                         cudaStream_t stream) {
  Wgrad3Params p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + 7) / 8;
  p.tiles_y = (H + 15) / 16;
  const int block_n = Cin_pad % 128 == 0 ? 128 : 64;
  p.co_tiles = (Cout_pad + 127) / 128;
  p.ci_tiles = Cin_pad / block_n;
  p.kblocks_total = (long long)B * p.tiles_x * p.tiles_y;
  const int units = 3 * p.co_tiles * p.ci_tiles;
  int splits = device_sm_count() / units;          // floor: the whole grid must be ONE wave (1 CTA per SM) --
  if (splits < 1) splits = 1;                      // a few CTAs spilling into a second wave double the kernel time
  if ((long long)splits > p.kblocks_total) splits = (int)p.kblocks_total;
  p.splits = splits;
  p.dw = dw;
  p.Cout_pad = Cout_pad;
  p.Cin_pad = Cin_pad;
  CUtensorMap tmDY, tmX;
  const uint32_t es[4] = {1, 1, 1, 1};
  {
    const uint32_t box[4] = {64, 8, 16, 1};
    uint64_t dims[4] = {(uint64_t)Cout_pad, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cout_pad * 2, (uint64_t)W * Cout_pad * 2, (uint64_t)H * W * Cout_pad * 2};
    if (make_tensor_map_f16(&tmDY, dy, 4, dims, str, box, es, "wgrad3 dY")) return -1;
  }
  {
    const uint32_t box[4] = {64, 10, 16, 1};
    uint64_t dims[4] = {(uint64_t)Cin_pad, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin_pad * 2, (uint64_t)W * Cin_pad * 2, (uint64_t)H * W * Cin_pad * 2};
    if (make_tensor_map_f16(&tmX, x, 4, dims, str, box, es, "wgrad3 X")) return -1;
  }
  if (block_n == 128) return launch_wgrad3<128>(tmDY, tmX, p, stream);
  return launch_wgrad3<64>(tmDY, tmX, p, stream);
}

// ---------------------------------------------------------------------------------------------
// wgrad for 3x3 / stride 1 / pad 1 with at most 64 output channels: ALL nine taps in one CTA, operands read ONCE.
// The dY operand needs only 64 of the MMA's 128 M rows, so the other 64 rows carry a second VIEW of the same
// shared-memory slab -- dY one image row further down (leading-dimension offset = one tile row = 1024 B): accumulator
// rows 0..63 then see tap row (rho), rows 64..127 tap row (rho - 1) of whatever X view the B operand presents.  B
// presents three column shifts at once (N = 3 x 64, chunk stride 128 B = one pixel of the halo slab) and the image
// row shift rho through its start address.  Two MMAs per 16 pixels (rho = 0, 1) produce taps
//   rho = 0: rows 0..63 -> dy =  0, rows 64..127 -> dy = -1;   rho = 1: rows 0..63 -> dy = +1 (rows 64..127: dy = 0
// again, discarded): 9 of 12 computed products are used, every byte of dY and X is fetched once instead of three
// times (the row kernel above splits the kernel rows across CTAs) and no MMA row multiplies TMA zero fill.
// The tile grid starts at image row -1 so that the shifted view covers row 0.
// ---------------------------------------------------------------------------------------------
constexpr int kC64ABytes = 18 * 1024;        // 17 image rows x 8 pixels x 128 B (17408), padded to the swizzle period
constexpr int kC64BBytes = 22 * 1024;        // 17 image rows x 10 pixels x 128 B (21760), padded
constexpr int kC64StageBytes = kC64ABytes + kC64BBytes;

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_c64_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                    const __grid_constant__ Wgrad3Params p) {
  constexpr uint32_t kIdesc = umma_idesc_f16_m128_mn(192);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t bar_base = smem_base + stages * kC64StageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * stages);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * stages + 1);
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
  if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int ci_t = blockIdx.x % p.ci_tiles;
  const int split = blockIdx.x / p.ci_tiles;
  const long long kb_lo = p.kblocks_total * split / p.splits;
  const long long kb_hi = p.kblocks_total * (split + 1) / p.splits;
  const int n_kb = (int)(kb_hi - kb_lo);

  if (warp == 0) {
    {   // whole warp, uniform control flow; one elected lane issues (see elect_one)
      int stage = 0;
      uint32_t phase = 0;
      const int tiles = p.tiles_x * p.tiles_y;
      for (long long kb = kb_lo; kb < kb_hi; ++kb) {
        const int b = (int)(kb / tiles);
        const int rr = (int)(kb - (long long)b * tiles);
        const int ty = rr / p.tiles_x, tx = rr - ty * p.tiles_x;
        const int x0 = tx * 8, y0 = ty * 16 - 1;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * kC64StageBytes;
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), (uint32_t)(17 * 8 * 128 + 17 * 10 * 128));
          tma_load_4d(sa, &tmDY, full_bar(stage), 0, x0, y0, b);
          tma_load_4d(sa + kC64ABytes, &tmX, full_bar(stage), ci_t * 64, x0 - 1, y0, b);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp, uniform control flow; one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * kC64StageBytes;
        const uint32_t sb = sa + kC64ABytes;
        if (elect_one()) {            // ONE election per k-block: the body is straight-line UTCHMMA + commits
#pragma unroll
          for (int rho = 0; rho < 2; ++rho) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {       // 16 pixels = 2 image rows of the tile per MMA
              const uint64_t adesc = umma_desc_mn_sw128(sa + (uint32_t)j * 2048u, 1024u);
              const uint64_t bdesc = umma_desc_mn_sw128_sbo(sb + (uint32_t)(2 * j + rho) * 1280u, 128u, 1280u);
              umma_f16(tmem_base + (uint32_t)(rho * 192), adesc, bdesc, kIdesc, (kb | j) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(stage));
          if (kb == n_kb - 1) umma_commit(done_bar);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
      if (n_kb == 0 && elect_one()) umma_commit(done_bar);
    }
    __syncwarp();
  } else if (n_kb > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int co = row & 63, half = row >> 6;          // half 1 = the view shifted one image row down
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int rho = 0; rho < 2; ++rho) {
      if (rho == 1 && half == 1) continue;             // (dy = 0 a second time)
      const int tap_dy = rho - half;                   // -1, 0, +1
#pragma unroll 1
      for (int sg = 0; sg < 3; ++sg) {
        float* out = p.dw + ((size_t)((tap_dy + 1) * 3 + sg) * p.Cout_pad + co) * p.Cin_pad + ci_t * 64;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(rho * 192 + sg * 64 + c * 32), v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) atomicAdd(out + c * 32 + i, __uint_as_float(v[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static int wgrad3x3_c64_impl(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin_pad,
                             cudaStream_t stream) {
  Wgrad3Params p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + 7) / 8;
  p.tiles_y = (H + 1 + 15) / 16;                   // tile rows start at image row -1
  p.co_tiles = 1;
  p.ci_tiles = Cin_pad / 64;
  p.kblocks_total = (long long)B * p.tiles_x * p.tiles_y;
  int splits = device_sm_count() / p.ci_tiles;     // one wave, see wgrad3x3_impl
  if (splits < 1) splits = 1;
  if ((long long)splits > p.kblocks_total) splits = (int)p.kblocks_total;
  p.splits = splits;
  p.dw = dw;
  p.Cout_pad = 64;
  p.Cin_pad = Cin_pad;
  CUtensorMap tmDY, tmX;
  const uint32_t es[4] = {1, 1, 1, 1};
  {
    const uint32_t box[4] = {64, 8, 17, 1};
    uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    if (make_tensor_map_f16(&tmDY, dy, 4, dims, str, box, es, "wgrad3 c64 dY")) return -1;
  }
  {
    const uint32_t box[4] = {64, 10, 17, 1};
    uint64_t dims[4] = {(uint64_t)Cin_pad, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin_pad * 2, (uint64_t)W * Cin_pad * 2, (uint64_t)H * W * Cin_pad * 2};
    if (make_tensor_map_f16(&tmX, x, 4, dims, str, box, es, "wgrad3 c64 X")) return -1;
  }
  int stages = (232448 - 1024 - 512) / kC64StageBytes;
  if (stages > 6) stages = 6;
  p.stages = stages;
  const int smem_bytes = 1024 + stages * kC64StageBytes + 512;
  static bool attr_set = false;
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(wgrad3x3_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int grid = p.splits * p.ci_tiles;
  wgrad3x3_c64_kernel<<<grid, kWgThreads, smem_bytes, stream>>>(tmDY, tmX, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// streaming kernels
// ---------------------------------------------------------------------------------------------
// dy = dy * scale * (y > 0): ReLU backward fused with the dynamic loss re-scaling; y and scale optional
__global__ void scale_mask_kernel(uint4* __restrict__ dy, const uint4* __restrict__ y, const float* __restrict__ scale,
                                  long long n8) {
  const __half2 zero = __float2half2_rn(0.0f);
  const __half2 sc = __float2half2_rn(scale != nullptr ? __ldg(scale) : 1.0f);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    uint4 g = dy[idx];
    __half2* gh = reinterpret_cast<__half2*>(&g);
    if (y != nullptr) {
      const uint4 a = __ldg(y + idx);
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) gh[j] = __hmul2(__hmul2(gh[j], sc), __hgt2(ah[j], zero));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) gh[j] = __hmul2(gh[j], sc);
    }
    dy[idx] = g;
  }
}

// dy = dy * scale * (y > 0) AND db[c] += sum_rows dy (after masking): ReLU backward, loss re-scaling and the bias
// gradient in ONE pass over dY.  dy / y fp16 [rows, C]; block = 32 channel pairs x 8 row lanes like bias_grad_kernel.
__global__ void scale_mask_bias_kernel(__half* __restrict__ dy, const __half* __restrict__ y,
                                       const float* __restrict__ scale, float* __restrict__ db, long long rows, int C) {
  const int cg = blockIdx.y;
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int c = cg * 64 + tc * 2;
  const float sc = scale != nullptr ? __ldg(scale) : 1.0f;
  float a0 = 0.0f, a1 = 0.0f;
  for (long long r = (long long)blockIdx.x * 8 + tr; r < rows; r += (long long)gridDim.x * 8) {
    __half2* gp = reinterpret_cast<__half2*>(dy + (size_t)r * C + c);
    float2 g = __half22float2(*gp);
    g.x *= sc; g.y *= sc;
    if (y != nullptr) {
      const float2 yv = __half22float2(*reinterpret_cast<const __half2*>(y + (size_t)r * C + c));
      if (!(yv.x > 0.0f)) g.x = 0.0f;
      if (!(yv.y > 0.0f)) g.y = 0.0f;
    }
    const __half2 gh = __floats2half2_rn(g.x, g.y);
    *gp = gh;
    const float2 gr = __half22float2(gh);          // sum what the consumers will read
    a0 += gr.x; a1 += gr.y;
  }
  __shared__ float sh[8][64];
  sh[tr][tc * 2] = a0;
  sh[tr][tc * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += sh[i][threadIdx.x];
    atomicAdd(db + cg * 64 + threadIdx.x, v);
  }
}

// out = max(out, max |x|) over an fp16 tensor; `out` holds the bits of a non-negative float (zeroed by the caller)
__global__ void absmax_kernel(const uint4* __restrict__ x, long long n8, int* __restrict__ out) {
  float m = 0.0f;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(x + idx);
    const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(__habs2(vh[j]));
      m = fmaxf(m, fmaxf(f.x, f.y));
    }
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, sh[i]);
    if (!(m <= 3.0e38f)) m = 3.0e38f;      // inf / nan -> huge, so the caller scales down
    atomicMax(out, __float_as_int(m));
  }
}

// 2x2/s2 max-pool backward (floor mode): gradient goes to the first maximum of each window (ATen order).
// One thread per window and 8 channels: reads the 4 inputs + dy once, writes the 4 gradients.  With `gate` the pooled
// tensor is a ReLU output and the gradient continues through that ReLU: a window whose maximum is 0 passes nothing
// (fuses the mask pass of the preceding conv layer, autograd of models.py:589 + the trunk's nn.ReLU).
__global__ void maxpool2_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                    int B, int H, int W, int C8, int Ho, int Wo, int gate) {
  const int Hc = (H + 1) >> 1, Wc = (W + 1) >> 1;      // windows incl. the partial ones of an odd edge (all-zero grad)
  const long long total = (long long)B * Hc * Wc * C8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    long long pix = idx / C8;
    const int ox = (int)(pix % Wc);
    pix /= Wc;
    const int oy = (int)(pix % Hc);
    const int b = (int)(pix / Hc);
    const size_t base = ((size_t)((size_t)b * H + 2 * oy) * W + 2 * ox) * C8 + c;
    const size_t off[4] = {0, (size_t)C8, (size_t)W * C8, (size_t)W * C8 + C8};
    uint4 out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = make_uint4(0, 0, 0, 0);
    if (oy < Ho && ox < Wo) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = __ldg(x + base + off[k]);
      const uint4 g = __ldg(dy + ((size_t)((size_t)b * Ho + oy) * Wo + ox) * C8 + c);
      const __half* gh = reinterpret_cast<const __half*>(&g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int arg = 0;
        float best = __half2float(reinterpret_cast<const __half*>(&v[0])[j]);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          const float f = __half2float(reinterpret_cast<const __half*>(&v[k])[j]);
          if (f > best) { best = f; arg = k; }
        }
        if (!gate || best > 0.0f) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k == arg) reinterpret_cast<__half*>(&out[k])[j] = gh[j];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int iy = 2 * oy + (k >> 1), ix = 2 * ox + (k & 1);
      if (iy < H && ix < W) dx[base + off[k]] = out[k];
    }
  }
}

// nearest x2 upsample backward: dx = sum of the 2x2 block of dy (fp32 sum, fp16 store)
__global__ void upsample2_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int B, int H, int W,
                                     int C8) {
  const long long total = (long long)B * H * W * C8;
  const int Wo = 2 * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    long long pix = idx / C8;
    const int ix = (int)(pix % W);
    pix /= W;
    const int iy = (int)(pix % H);
    const int b = (int)(pix / H);
    const size_t base = ((size_t)((size_t)b * 2 * H + 2 * iy) * Wo + 2 * ix) * C8 + c;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint4 v = __ldg(dy + base + (size_t)(k >> 1) * Wo * C8 + (size_t)(k & 1) * C8);
      const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(vh[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    uint4 o;
    __half2 h;
    h = __floats2half2_rn(acc[0], acc[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(acc[2], acc[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(acc[4], acc[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2half2_rn(acc[6], acc[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    dx[idx] = o;
  }
}

// bias gradient: db[c] += sum over rows of dy[row][c]  (dy fp16 [rows, C], db fp32 zeroed by the caller)
__global__ void bias_grad_kernel(const __half* __restrict__ dy, float* __restrict__ db, long long rows, int C) {
  // block = 256 threads: 32 channel-pairs (64 channels) x 8 row lanes; grid.y = channel groups of 64
  const int cg = blockIdx.y;
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int c = cg * 64 + tc * 2;
  float a0 = 0.0f, a1 = 0.0f;
  for (long long r = (long long)blockIdx.x * 8 + tr; r < rows; r += (long long)gridDim.x * 8) {
    const __half2 v = *reinterpret_cast<const __half2*>(dy + (size_t)r * C + c);
    const float2 f = __half22float2(v);
    a0 += f.x;
    a1 += f.y;
  }
  __shared__ float sh[8][64];
  sh[tr][tc * 2] = a0;
  sh[tr][tc * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][threadIdx.x];
    atomicAdd(db + cg * 64 + threadIdx.x, s);
  }
}

// ---- BatchNorm2d in training mode (torchvision ResNet trunk + decoder BNs, models.py:22-32,46-76) ----
// Deterministic two-stage column reduction shared by bn_stats and bn_bwd_reduce: every block writes its 2 x 64 partial
// sums to `partials[which][block][C]`, takes a ticket on `tickets[channel group]`, and the block that draws the LAST
// ticket adds the partials in block order and WRITES the result (no floating-point atomics: the summation order, and
// with it the batch statistics and every ReLU decision downstream, is the same on every run -- the reference gets
// this from cudnn.deterministic, dream/utilities.py:15-26).  The winner also re-arms the ticket for the next launch.
__device__ __forceinline__ void two_stage_finish(float (*sh)[8][64], float* __restrict__ partials,
                                                 unsigned* __restrict__ tickets, float* __restrict__ out0,
                                                 float* __restrict__ out1, int C) {
  const int cg = blockIdx.y;
  const int nb = gridDim.x;
  __shared__ bool last;
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += sh[which][i][ch];
    partials[((size_t)which * nb + blockIdx.x) * C + cg * 64 + ch] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tickets + cg, 1u) == (unsigned)nb - 1u;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;
    const float* p = partials + (size_t)which * nb * C + cg * 64 + ch;
    float v = 0.f;
    for (int b = 0; b < nb; ++b) v += __ldcg(p + (size_t)b * C);
    (which ? out1 : out0)[cg * 64 + ch] = v;
  }
  if (threadIdx.x == 0) tickets[cg] = 0u;
}

// per-channel sum and sum of squares of z [rows, C] fp16 (fp32 accumulation, fixed summation order)
__global__ void bn_stats_kernel(const __half* __restrict__ z, float* __restrict__ sum, float* __restrict__ sumsq,
                                long long rows, int C, float* __restrict__ partials, unsigned* __restrict__ tickets) {
  const int cg = blockIdx.y;
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int c = cg * 64 + tc * 2;
  float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (long long r = (long long)blockIdx.x * 8 + tr; r < rows; r += (long long)gridDim.x * 8) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(z + (size_t)r * C + c));
    a0 += f.x; a1 += f.y; q0 += f.x * f.x; q1 += f.y * f.y;
  }
  __shared__ float sh[2][8][64];
  sh[0][tr][tc * 2] = a0; sh[0][tr][tc * 2 + 1] = a1;
  sh[1][tr][tc * 2] = q0; sh[1][tr][tc * 2 + 1] = q1;
  __syncthreads();
  two_stage_finish(sh, partials, tickets, sum, sumsq, C);
}

// y = relu?( z * scale[c] + shift[c] (+ residual) ), fp16 in / out, 8 channels per thread
__global__ void bn_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ scale,
                                const float* __restrict__ shift, const uint4* __restrict__ residual,
                                uint4* __restrict__ y, long long n8, int C8, int relu) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8) * 8;
    const uint4 v = __ldg(z + idx);
    const __half2* vh = reinterpret_cast<const __half2*>(&v);
    uint4 rv = make_uint4(0, 0, 0, 0);
    if (residual != nullptr) rv = __ldg(residual + idx);
    const __half2* rh = reinterpret_cast<const __half2*>(&rv);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(vh[j]);
      const float2 r = __half22float2(rh[j]);
      float a = f.x * __ldg(scale + c + 2 * j) + __ldg(shift + c + 2 * j) + r.x;
      float b = f.y * __ldg(scale + c + 2 * j + 1) + __ldg(shift + c + 2 * j + 1) + r.y;
      if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
      oh[j] = __floats2half2_rn(a, b);
    }
    y[idx] = o;
  }
}

// BN backward reductions: sum_dy[c] = sum dy, sum_dyz[c] = sum dy * z   (dy already ReLU-masked; fixed order)
__global__ void bn_bwd_reduce_kernel(const __half* __restrict__ dy, const __half* __restrict__ z,
                                     float* __restrict__ sum_dy, float* __restrict__ sum_dyz, long long rows, int C,
                                     float* __restrict__ partials, unsigned* __restrict__ tickets) {
  const int cg = blockIdx.y;
  const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
  const int c = cg * 64 + tc * 2;
  float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (long long r = (long long)blockIdx.x * 8 + tr; r < rows; r += (long long)gridDim.x * 8) {
    const float2 g = __half22float2(*reinterpret_cast<const __half2*>(dy + (size_t)r * C + c));
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(z + (size_t)r * C + c));
    a0 += g.x; a1 += g.y; q0 += g.x * f.x; q1 += g.y * f.y;
  }
  __shared__ float sh[2][8][64];
  sh[0][tr][tc * 2] = a0; sh[0][tr][tc * 2 + 1] = a1;
  sh[1][tr][tc * 2] = q0; sh[1][tr][tc * 2 + 1] = q1;
  __syncthreads();
  two_stage_finish(sh, partials, tickets, sum_dy, sum_dyz, C);
}

// dz = a[c] * dy + b[c] * z + c0[c]   (BN backward, coefficients precomputed per channel), in place on dy
__global__ void bn_bwd_apply_kernel(uint4* __restrict__ dy, const uint4* __restrict__ z, const float* __restrict__ a,
                                    const float* __restrict__ b, const float* __restrict__ c0, long long n8, int C8) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8) * 8;
    uint4 g = dy[idx];
    const uint4 v = __ldg(z + idx);
    __half2* gh = reinterpret_cast<__half2*>(&g);
    const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 gf = __half22float2(gh[j]);
      const float2 zf = __half22float2(vh[j]);
      const int c0i = c + 2 * j;
      gh[j] = __floats2half2_rn(__ldg(a + c0i) * gf.x + __ldg(b + c0i) * zf.x + __ldg(c0 + c0i),
                                __ldg(a + c0i + 1) * gf.y + __ldg(b + c0i + 1) * zf.y + __ldg(c0 + c0i + 1));
    }
    dy[idx] = g;
  }
}

// 3x3 / stride 2 / pad 1 max-pool backward (resnet stem, models.py:27): every input pixel gathers from the <= 4
// windows that contain it; a window's gradient goes to its first maximum in (row, col) scan order, like ATen.
__global__ void maxpool3_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                    int B, int H, int W, int C8, int Ho, int Wo) {
  const long long total = (long long)B * H * W * C8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8);
    long long pix = idx / C8;
    const int ix = (int)(pix % W);
    pix /= W;
    const int iy = (int)(pix % H);
    const int b = (int)(pix / H);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint4 mine = __ldg(x + idx);
    const __half* mh = reinterpret_cast<const __half*>(&mine);
    // windows (oy, ox) with 2*oy-1 <= iy <= 2*oy+1
    for (int oy = (iy) / 2; oy <= (iy + 1) / 2; ++oy) {
      if (oy < 0 || oy >= Ho) continue;
      for (int ox = (ix) / 2; ox <= (ix + 1) / 2; ++ox) {
        if (ox < 0 || ox >= Wo) continue;
        const uint4 g = __ldg(dy + ((size_t)((size_t)b * Ho + oy) * Wo + ox) * C8 + c);
        const __half* gh = reinterpret_cast<const __half*>(&g);
        // position of (iy, ix) in the window's scan order
        const int my_r = iy - (2 * oy - 1), my_c = ix - (2 * ox - 1);
        bool win[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = true;
        for (int r = 0; r < 3; ++r) {
          const int yy = 2 * oy - 1 + r;
          if (yy < 0 || yy >= H) continue;
          for (int q = 0; q < 3; ++q) {
            const int xx = 2 * ox - 1 + q;
            if (xx < 0 || xx >= W) continue;
            if (r == my_r && q == my_c) continue;
            const uint4 v = __ldg(x + ((size_t)((size_t)b * H + yy) * W + xx) * C8 + c);
            const __half* vh = reinterpret_cast<const __half*>(&v);
            const bool before = (r < my_r) || (r == my_r && q < my_c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float o = __half2float(vh[j]), m = __half2float(mh[j]);
              // an earlier element wins ties (>=), a later one only if strictly greater
              if (before ? (o >= m) : (o > m)) win[j] = false;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (win[j]) acc[j] += __half2float(gh[j]);
      }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    dx[idx] = o;
  }
}

static int grid_cap(long long work, int threads) {
  long long blocks = (work + threads - 1) / threads;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace db200

using namespace db200;

// H, W: the grid the pixel tiles run over (the coarser of the two operands); Hd x Wd: dY extent; Hx x Wx: X extent
static int wgrad_impl(const void* dy, const void* x, float* dw, int B, int H, int W, int Hd, int Wd, int Hx, int Wx,
                      int a_stride, int b_stride, int shift_a, int Cout_pad, int Cin_pad, int taps, const int8_t* tap_dy, const int8_t* tap_dx,
                      cudaStream_t stream) {
  DB_REQUIRE(dy && x && dw && tap_dy && tap_dx, "wgrad: null pointer");
  DB_REQUIRE(Cout_pad % 64 == 0 && Cin_pad % 64 == 0, "wgrad: channel counts must be multiples of 64 (%d, %d)",
             Cout_pad, Cin_pad);
  DB_REQUIRE(taps >= 1 && taps <= DREAMB200_MAX_TAPS, "wgrad: taps=%d out of range", taps);
  DB_REQUIRE(B > 0 && H > 0 && W > 0 && Hd > 0 && Wd > 0, "wgrad: empty input");
  DB_REQUIRE(a_stride == 1 || a_stride == 2, "wgrad: a_stride=%d unsupported", a_stride);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + kWgTw - 1) / kWgTw;
  p.tiles_y = (H + kWgTh - 1) / kWgTh;
  p.taps = taps;
  memcpy(p.dy, tap_dy, taps);
  memcpy(p.dx, tap_dx, taps);
  p.a_stride = a_stride;
  p.b_stride = b_stride;
  p.shift_a = shift_a;
  const int block_n = Cin_pad % 128 == 0 ? 128 : 64;
  p.co_tiles = (Cout_pad + 127) / 128;   // a half-empty last tile is zero-filled by TMA
  p.ci_tiles = Cin_pad / block_n;
  p.kblocks_total = (long long)B * p.tiles_x * p.tiles_y;
  const int units = taps * p.co_tiles * p.ci_tiles;
  int splits = device_sm_count() / units;          // floor: the whole grid must be ONE wave (1 CTA per SM) --
  if (splits < 1) splits = 1;                      // a few CTAs spilling into a second wave double the kernel time
  if ((long long)splits > p.kblocks_total) splits = (int)p.kblocks_total;
  p.splits = splits;
  p.dw = dw;
  p.Cout_pad = Cout_pad;
  p.Cin_pad = Cin_pad;

  CUtensorMap tmDY, tmX;
  {
    const uint32_t s = (uint32_t)a_stride;
    const uint32_t box[4] = {64, kWgTw * s, kWgTh * s, 1};
    const uint32_t es[4] = {1, s, s, 1};
    uint64_t dims[4] = {(uint64_t)Cout_pad, (uint64_t)Wd, (uint64_t)Hd, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cout_pad * 2, (uint64_t)Wd * Cout_pad * 2, (uint64_t)Hd * Wd * Cout_pad * 2};
    if (make_tensor_map_f16(&tmDY, dy, 4, dims, str, box, es, "wgrad dY")) return -1;
  }
  {
    const uint32_t s = (uint32_t)b_stride;
    const uint32_t box[4] = {64, kWgTw * s, kWgTh * s, 1};
    const uint32_t es[4] = {1, s, s, 1};
    uint64_t dims[4] = {(uint64_t)Cin_pad, (uint64_t)Wx, (uint64_t)Hx, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin_pad * 2, (uint64_t)Wx * Cin_pad * 2, (uint64_t)Hx * Wx * Cin_pad * 2};
    if (make_tensor_map_f16(&tmX, x, 4, dims, str, box, es, "wgrad X")) return -1;
  }
  if (block_n == 128) return launch_wgrad<128>(tmDY, tmX, p, stream);
  return launch_wgrad<64>(tmDY, tmX, p, stream);
}

namespace db200 {
int try_wgrad3x3_pair(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout_pad, int Cin_pad,
                      cudaStream_t stream);   // wgrad_pair.cu (opt-in CTA-pair kernel)
}
extern "C" int dreamb200_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout_pad,
                               int Cin_pad, int taps, const int8_t* tap_dy, const int8_t* tap_dx, void* stream_v) {
  static int use_row_kernel = -1;
  if (use_row_kernel < 0) {
    const char* e = getenv("DREAMB200_WGRAD3");       // 0 disables the row-shared 3x3 kernel (A/B measurements)
    use_row_kernel = e ? atoi(e) : 1;
  }
  bool std3x3 = use_row_kernel && taps == 9 && dy && x && dw && tap_dy && tap_dx && Cout_pad % 64 == 0 &&
                Cin_pad % 64 == 0 && B > 0 && H > 0 && W > 0;
  for (int t = 0; std3x3 && t < 9; ++t) std3x3 = tap_dy[t] == t / 3 - 1 && tap_dx[t] == t % 3 - 1;
  static int use_c64 = -1;
  if (use_c64 < 0) {
    const char* e = getenv("DREAMB200_WGRAD_C64");    // 0 disables the all-taps kernel for <= 64 output channels
    use_c64 = e ? atoi(e) : 1;
  }
  if (std3x3 && use_c64 && Cout_pad == 64)
    return wgrad3x3_c64_impl(dy, x, dw, B, H, W, Cin_pad, (cudaStream_t)stream_v);
  if (std3x3) {
    const int r = try_wgrad3x3_pair(dy, x, dw, B, H, W, Cout_pad, Cin_pad, (cudaStream_t)stream_v);   // opt-in
    if (r != 0) return r > 0 ? 0 : r;
    return wgrad3x3_impl(dy, x, dw, B, H, W, Cout_pad, Cin_pad, (cudaStream_t)stream_v);
  }
  return wgrad_impl(dy, x, dw, B, H, W, H, W, H, W, 1, 1, 0, Cout_pad, Cin_pad, taps, tap_dy, tap_dx,
                    (cudaStream_t)stream_v);
}

extern "C" int dreamb200_wgrad_deconv(const void* dy, const void* x, float* dw, int B, int H, int W, int Cout_pad,
                                      int Cin_pad, int taps, const int8_t* tap_dy, const int8_t* tap_dx,
                                      void* stream_v) {
  return wgrad_impl(dy, x, dw, B, H, W, 2 * H, 2 * W, H, W, 2, 1, 1, Cout_pad, Cin_pad, taps, tap_dy, tap_dx,
                    (cudaStream_t)stream_v);
}

// weight gradient of a stride-2 convolution: dy [B,Ho,Wo,Cout_pad], x [B,Hx,Wx,Cin_pad] with Ho=(Hx-1)/2+1:
// dW[tap][co][ci] += sum_p dY[p][co] * X[2p + tap][ci]
extern "C" int dreamb200_wgrad_strided(const void* dy, const void* x, float* dw, int B, int Ho, int Wo, int Hx, int Wx,
                                       int Cout_pad, int Cin_pad, int taps, const int8_t* tap_dy, const int8_t* tap_dx,
                                       void* stream_v) {
  return wgrad_impl(dy, x, dw, B, Ho, Wo, Ho, Wo, Hx, Wx, 1, 2, 0, Cout_pad, Cin_pad, taps, tap_dy, tap_dx,
                    (cudaStream_t)stream_v);
}

extern "C" int dreamb200_scale_mask_f16(void* dy, const void* y, const float* scale, long long n, void* stream) {
  DB_REQUIRE(dy && n > 0 && n % 8 == 0, "scale_mask: bad arguments");
  scale_mask_kernel<<<grid_cap(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<uint4*>(dy), reinterpret_cast<const uint4*>(y), scale, n / 8);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_scale_mask_bias_f16(void* dy, const void* y, const float* scale, float* db, long long rows,
                                             int C, void* stream) {
  DB_REQUIRE(dy && db && rows > 0 && C % 64 == 0, "scale_mask_bias: bad arguments");
  long long bx = (rows + 8 * 16 - 1) / (8 * 16);
  if (bx > device_sm_count() * 8) bx = device_sm_count() * 8;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)(C / 64));
  scale_mask_bias_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<__half*>(dy),
                                                                 reinterpret_cast<const __half*>(y), scale, db, rows, C);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// One step of the backward pass' loss re-scaling, entirely on the device: f = 2^floor(log2(target / max|dY|))
// (clamped to 2^+-20), cum *= f, inv = 1 / cum.  Replaces ~10 single-element torch kernels per layer.
__global__ void loss_scale_step_kernel(const float* __restrict__ amax, float* __restrict__ cum, float* __restrict__ f_out,
                                       float* __restrict__ inv_out, float target) {
  const float a = fmaxf(*amax, 1e-30f);
  float f = exp2f(floorf(log2f(target / a)));
  f = fminf(fmaxf(f, 9.5367431640625e-07f), 1048576.0f);
  const float c = *cum * f;
  *cum = c;
  *f_out = f;
  *inv_out = 1.0f / c;
}

extern "C" int dreamb200_loss_scale_step(const float* amax, float* cum, float* f_out, float* inv_out, float target,
                                         void* stream) {
  DB_REQUIRE(amax && cum && f_out && inv_out && target > 0.0f, "loss_scale_step: bad arguments");
  loss_scale_step_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(amax, cum, f_out, inv_out, target);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_absmax_f16(const void* x, long long n, float* out, void* stream) {
  DB_REQUIRE(x && out && n > 0 && n % 8 == 0, "absmax: bad arguments");
  absmax_kernel<<<grid_cap(n / 8, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), n / 8,
                                                                       reinterpret_cast<int*>(out));
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_maxpool2_bwd_nhwc(const void* x, const void* dy, void* dx, int B, int H, int W, int C,
                                           int relu_gate, void* stream) {
  DB_REQUIRE(x && dy && dx && C % 8 == 0, "maxpool2_bwd: bad arguments");
  const long long total = (long long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  maxpool2_bwd_kernel<<<grid_cap(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), B, H, W,
      C / 8, H / 2, W / 2, relu_gate);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_upsample2_bwd_nhwc(const void* dy, void* dx, int B, int H, int W, int C, void* stream) {
  DB_REQUIRE(dy && dx && C % 8 == 0, "upsample2_bwd: bad arguments");
  const long long total = (long long)B * H * W * (C / 8);
  upsample2_bwd_kernel<<<grid_cap(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), B, H, W, C / 8);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_bias_grad(const void* dy, float* db, long long rows, int C, void* stream) {
  DB_REQUIRE(dy && db && rows > 0 && C % 64 == 0, "bias_grad: bad arguments");
  long long bx = (rows + 8 * 64 - 1) / (8 * 64);
  if (bx > device_sm_count() * 4) bx = device_sm_count() * 4;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)(C / 64));
  bias_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(dy), db, rows, C);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

static dim3 reduce_grid(long long rows, int C) {
  long long bx = (rows + 8 * 64 - 1) / (8 * 64);
  if (bx > device_sm_count() * 4) bx = device_sm_count() * 4;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)(C / 64));
}

extern "C" int dreamb200_bn_reduce_workspace(long long rows, int C, long long* partial_floats, int* n_tickets) {
  DB_REQUIRE(partial_floats && n_tickets && rows > 0 && C % 64 == 0, "bn_reduce_workspace: bad arguments");
  *partial_floats = 2LL * reduce_grid(rows, C).x * C;
  *n_tickets = C / 64;
  return 0;
}

extern "C" int dreamb200_bn_stats_f16(const void* z, float* sum, float* sumsq, long long rows, int C, float* partials,
                                      unsigned* tickets, void* stream) {
  DB_REQUIRE(z && sum && sumsq && partials && tickets && rows > 0 && C % 64 == 0, "bn_stats: bad arguments");
  bn_stats_kernel<<<reduce_grid(rows, C), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __half*>(z), sum,
                                                                          sumsq, rows, C, partials, tickets);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_bn_apply_f16(const void* z, const float* scale, const float* shift, const void* residual,
                                      void* y, long long rows, int C, int relu, void* stream) {
  DB_REQUIRE(z && scale && shift && y && rows > 0 && C % 8 == 0, "bn_apply: bad arguments");
  const long long n8 = rows * C / 8;
  bn_apply_kernel<<<grid_cap(n8, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(z), scale, shift, reinterpret_cast<const uint4*>(residual),
      reinterpret_cast<uint4*>(y), n8, C / 8, relu);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_bn_bwd_reduce_f16(const void* dy, const void* z, float* sum_dy, float* sum_dyz, long long rows,
                                           int C, float* partials, unsigned* tickets, void* stream) {
  DB_REQUIRE(dy && z && sum_dy && sum_dyz && partials && tickets && rows > 0 && C % 64 == 0,
             "bn_bwd_reduce: bad arguments");
  bn_bwd_reduce_kernel<<<reduce_grid(rows, C), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __half*>(dy), reinterpret_cast<const __half*>(z), sum_dy, sum_dyz, rows, C, partials,
      tickets);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_bn_bwd_apply_f16(void* dy, const void* z, const float* a, const float* b, const float* c0,
                                          long long rows, int C, void* stream) {
  DB_REQUIRE(dy && z && a && b && c0 && rows > 0 && C % 8 == 0, "bn_bwd_apply: bad arguments");
  const long long n8 = rows * C / 8;
  bn_bwd_apply_kernel<<<grid_cap(n8, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<uint4*>(dy), reinterpret_cast<const uint4*>(z), a, b, c0, n8, C / 8);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int dreamb200_maxpool3_bwd_nhwc(const void* x, const void* dy, void* dx, int B, int H, int W, int C,
                                           void* stream) {
  DB_REQUIRE(x && dy && dx && C % 8 == 0, "maxpool3_bwd: bad arguments");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)B * H * W * (C / 8);
  maxpool3_bwd_kernel<<<grid_cap(total, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), B, H, W,
      C / 8, Ho, Wo);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
