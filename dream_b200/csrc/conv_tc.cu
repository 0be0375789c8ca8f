// conv_tc.cu -- im2col-free implicit-GEMM convolution on the sm_100a tensor cores.
//
// One kernel covers every conv / deconv-phase / 1x1 layer of the DREAM networks
// (reference: dream/models.py:591-747 vgg path, :22-136 resnet path).  A convolution is
// evaluated as   Y[p, co] = sum_tap sum_ci X[p*stride + (dy,dx)_tap, ci] * Wt[tap][co][ci]
// i.e. a sum over "taps" of shifted 1x1 GEMMs:
//   * A operand  : a (tw x th) patch of output pixels -> tw*th <= 128 rows of 64 input channels,
//                  fetched by ONE 4-D TMA box per (tap, channel chunk) straight from the NHWC
//                  activation; the tap shift is a coordinate offset and the zero padding is the
//                  TMA out-of-bounds fill, so no im2col buffer ever exists.
//   * B operand  : weights [tap][Cout][Cin] fp16, a 3-D TMA box of BLOCK_N x 64.
//   * D          : fp32 accumulator in TMEM (128 lanes x BLOCK_N columns), double buffered so the
//                  epilogue of tile i overlaps the MMAs of tile i+1.
//   * epilogue   : tcgen05.ld -> +bias (+residual) (ReLU) -> fp16 -> 128B-swizzled smem -> TMA store
//                  (the store box clips partial tiles), or fp32 NCHW for the network head.
// Warp roles (192 threads): warp0 = TMA producer, warp1 = MMA issuer + TMEM owner, warps2-5 = epilogue.
// Persistent: grid = #SMs, tiles round-robin.
#include "common.cuh"
#include "conv_common.cuh"
#include "dreamb200.h"

#include <mutex>

namespace db200 {

// NHWC mode runs 8 epilogue warps (two per TMEM lane quarter, 32 of each chunk's 64 columns each): for layers with a
// short K loop (1x1 convs, few input channels) the epilogue, not the MMA, paces a tile.  The fp32 NCHW head keeps 4.
template <int OUT_MODE>
__host__ __device__ constexpr int kTcThreads() { return OUT_MODE == DREAMB200_OUT_NHWC_F16 ? 64 + 256 : 64 + 128; }
constexpr int kTcSplit = 2;

template <int BLOCK_N, int OUT_MODE, bool PLAIN = false>
__global__ void __launch_bounds__(kTcThreads<OUT_MODE>(), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP,
               const __grid_constant__ ConvParams p) {
  constexpr int kBBytes = BLOCK_N * 128;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                            : (2 * BLOCK_N <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = umma_idesc_f16_m128(BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int stages = p.stages;
  const uint32_t smem_ab = smem_base;                                   // stages * kStageBytes
  const uint32_t smem_out = smem_ab + stages * kStageBytes;             // 2 * 16 KB (NHWC mode)
  const uint32_t out_bytes =
      (OUT_MODE == DREAMB200_OUT_NHWC_F16) ? 2 * kStageOutBytes + (p.pool ? 2 * kPoolBytes : 0) : 0;
  const uint32_t smem_pool = smem_out + 2 * kStageOutBytes;             // 2 * 4 KB pooled staging (pool only)
  const uint32_t bar_base = smem_out + out_bytes;                       // barriers (8 B each)
  // full[s], empty[s], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * stages + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * stages + 4);
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));
  const uint32_t smem_bias = bar_base + 256u;                           // BLOCK_N fp32 (NHWC mode)
  float* smem_bias_gen = reinterpret_cast<float*>(smem_gen + (smem_bias - smem_base));
  if (OUT_MODE == DREAMB200_OUT_NHWC_F16) stage_bias(p, smem_bias_gen, BLOCK_N);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (OUT_MODE == DREAMB200_OUT_NHWC_F16) {
      tma_prefetch_desc(&tmC);
      if (p.pool) tma_prefetch_desc(&tmP);
    }
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), OUT_MODE == DREAMB200_OUT_NHWC_F16 ? 4 * kTcSplit : 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const int num_kb = p.taps * p.kchunks;

  if (warp == 0) {
    // ===================== TMA producer (whole warp in uniform control flow, one elected lane issues) ==========
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int n, tx, ty, b;
        decode_tile(p, tile, n, tx, ty, b);
        const int x0 = tx * p.tw * p.in_stride, y0 = ty * p.th * p.in_stride;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int xi = x0 + p.dx[tap], yi = y0 + p.dy[tap];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_ab + stage * kStageBytes;
            if (elect_one()) {
              mbar_expect_tx(full_bar(stage), (uint32_t)(p.tw * p.th * 128 + kBBytes));
              tma_load_4d(sa, &tmA, full_bar(stage), kc * 64, xi, yi, b);
              tma_load_3d(sa + kABytes, &tmB, full_bar(stage), kc * 64, n * BLOCK_N, tap);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in uniform control flow, one elected lane issues) ==========
    {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_ab + stage * kStageBytes;
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
          if (elect_one()) {      // ONE election per k-block: the body is straight-line UTCHMMA + commits
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 16 fp16 = 32 B along K inside the 128 B swizzle row: +2 in 16 B units
              umma_f16(d_tmem, adesc + 2u * k, bdesc + 2u * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));
            if (kb == num_kb - 1) umma_commit(tfull_bar(as));
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9 NHWC / 2..5 NCHW head) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int hsel = (warp - 2) >> 2;       // which 32 of a chunk's 64 columns (NHWC mode)
    const int row = q * 32 + lane;          // accumulator row == output pixel inside the tile
    const int epi_tid = threadIdx.x - 64;
    const int ly = row / p.tw, lx = row - ly * p.tw;
    int as = 0;
    uint32_t aphase = 0;
    uint32_t chunk_ctr = 0;
    float csum[kCsumSize<BLOCK_N, kTcSplit>];
#pragma unroll
    for (int i = 0; i < kCsumSize<BLOCK_N, kTcSplit>; ++i) csum[i] = 0.0f;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int n, tx, ty, b;
      decode_tile(p, tile, n, tx, ty, b);
      const int ox = tx * p.tw + lx, oy = ty * p.th + ly;
      const bool valid = (ly < p.th) && (ox < p.Wo) && (oy < p.Ho);
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N);

      if constexpr (OUT_MODE == DREAMB200_OUT_NHWC_F16) {
        epilogue_nhwc_tile<BLOCK_N, kTcSplit, false, PLAIN>(p, &tmC, &tmP, t_row, smem_out, smem_pool, smem_bias, smem_bias_gen,
                                              tempty_bar(as), n, tx, ty, b, ox, oy, valid, row, lane, epi_tid,
                                              chunk_ctr, hsel, csum, nullptr, nullptr, tfull_bar(as), aphase);
      } else {
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        // fp32 NCHW head: BLOCK_N == 16 accumulator columns, first cout_real are real channels
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row, v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
        if (valid) {
          const size_t plane = (size_t)p.Ho * p.Wo;
          float* o = p.out_f32 + (size_t)b * p.cout_real * plane + (size_t)oy * p.Wo + ox;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (k < p.cout_real) {
              float f = __uint_as_float(v[k]) + (p.bias != nullptr ? __ldg(p.bias + k) : 0.0f);
              if (p.relu) f = fmaxf(f, 0.0f);
              o[(size_t)k * plane] = f;
            }
          }
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if constexpr (OUT_MODE == DREAMB200_OUT_NHWC_F16) flush_colsum<BLOCK_N, kTcSplit>(p, csum, lane, hsel);
    if (OUT_MODE == DREAMB200_OUT_NHWC_F16 && epi_tid < 32) {
      if (elect_one()) tma_store_wait_read<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box,
                        const uint32_t* estride, const char* what) {
  PFN_encodeTiled enc = get_encode();
  DB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return 0;
}

// fp32 tensor, SWIZZLE_128B (inner box = 32 floats = one 128-byte row): residual tiles of conv_tc2.cu.
int make_tensor_map_f32_sw128(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                              const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, const char* what) {
  PFN_encodeTiled enc = get_encode();
  DB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  uint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return 0;
}

// Un-swizzled map over a plain fp32 / uint8 tensor (input staging of first_conv.cu). kind: 0 = fp32, 1 = uint8.
int make_tensor_map_plain(CUtensorMap* tm, int kind, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, const char* what) {
  PFN_encodeTiled enc = get_encode();
  DB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  uint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank,
                   const_cast<void*>(base), dims, strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return 0;
}

// Pick the output patch (tw x th <= 128 pixels) that wastes the fewest accumulator rows.
// `even` forces even patch sides (needed when a 2x2 pool is fused on top of the patch).
static double choose_tile(int Wo, int Ho, int in_stride, bool even, int* tw_out, int* th_out) {
  double best = -1.0;
  int btw = 1, bth = 1;
  const int max_side = 256 / in_stride;  // TMA box side limit (in input elements)
  const int step = even ? 2 : 1;
  for (int tw = step; tw <= 128 && tw <= Wo + (even ? 1 : 0) && tw <= max_side; tw += step) {
    int th = 128 / tw;
    if (th > Ho + (even ? 1 : 0)) th = Ho + (even ? 1 : 0);
    if (th > max_side) th = max_side;
    if (even) th &= ~1;
    if (th < 1) continue;
    const long tiles = (long)((Wo + tw - 1) / tw) * ((Ho + th - 1) / th);
    const double util = (double)Wo * Ho / (double)(tiles * 128);
    // on ties prefer the wider patch (longer contiguous runs per TMA row)
    if (util > best + 1e-9 || (util > best - 1e-9 && tw > btw)) {
      best = util;
      btw = tw;
      bth = th;
    }
  }
  *tw_out = btw;
  *th_out = bth;
  return best;
}

double conv_choose_tile(int Wo, int Ho, int in_stride, bool even, int* tw_out, int* th_out) {   // for conv_tc2.cu
  return choose_tile(Wo, Ho, in_stride, even, tw_out, th_out);
}

int device_sm_count();
int try_conv_tc2(const dreamb200_conv_desc* d, cudaStream_t stream);  // conv_tc2.cu
int try_conv_tc2_phases(const dreamb200_conv_desc* descs, int n_phases, cudaStream_t stream);
int try_conv_rs(const dreamb200_conv_desc* d, cudaStream_t stream);   // conv_rs.cu
int try_conv_rs2(const dreamb200_conv_desc* d, cudaStream_t stream);  // conv_rs2.cu
int try_conv_rs3(const dreamb200_conv_desc* d, cudaStream_t stream);  // conv_rs3.cu

template <int BLOCK_N, int OUT_MODE>
static int launch(const dreamb200_conv_desc* d, cudaStream_t stream, int num_sms) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  choose_tile(d->Wo, d->Ho, d->in_stride, d->y_pool != nullptr, &p.tw, &p.th);
  p.pool = d->y_pool != nullptr ? 1 : 0;
  p.store_full = (d->y != nullptr) ? 1 : 0;
  p.tiles_x = (d->Wo + p.tw - 1) / p.tw;
  p.tiles_y = (d->Ho + p.th - 1) / p.th;
  p.n_tiles = d->Cout_pad / BLOCK_N;
  p.B = d->B;
  p.Ho = d->Ho;
  p.Wo = d->Wo;
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
  p.in_stride = d->in_stride;
  p.taps = d->taps;
  p.kchunks = d->Cin / 64;
  memcpy(p.dy, d->tap_dy, sizeof(p.dy));
  memcpy(p.dx, d->tap_dx, sizeof(p.dx));
  p.bias = d->bias;
  p.residual = reinterpret_cast<const __half*>(d->residual);
  p.residual_f32 = d->residual_f32;
  p.y_f32 = d->y_f32;
  p.Cout_pad = d->Cout_pad;
  p.relu = d->relu;
  p.out_f32 = (OUT_MODE == DREAMB200_OUT_NCHW_F32) ? reinterpret_cast<float*>(d->y) : nullptr;
  p.cout_real = d->cout_real;

  constexpr int kStageBytes = kABytes + BLOCK_N * 128;
  const int out_bytes =
      (OUT_MODE == DREAMB200_OUT_NHWC_F16) ? 2 * kStageOutBytes + (p.pool ? 2 * kPoolBytes : 0) : 0;
  const int budget = 232448 - 1024 - out_bytes - 256 - BLOCK_N * 4;
  int stages = budget / kStageBytes;
  if (stages > 8) stages = 8;
  DB_REQUIRE(stages >= 2, "conv: not enough shared memory for 2 stages");
  p.stages = stages;
  const int smem_bytes = 1024 + stages * kStageBytes + out_bytes + 256 + BLOCK_N * 4;

  // A: activation NHWC (C, W, H, B); box (64, tw, th, 1); element stride = conv stride
  CUtensorMap tmA, tmB, tmC, tmP;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmP, 0, sizeof(tmP));
  {
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->Cin * 2, (uint64_t)d->W * d->Cin * 2, (uint64_t)d->H * d->W * d->Cin * 2};
    const uint32_t s = (uint32_t)d->in_stride;
    uint32_t box[4] = {64, (uint32_t)p.tw * s, (uint32_t)p.th * s, 1};
    uint32_t es[4] = {1, s, s, 1};
    DB_REQUIRE(box[1] <= 256 && box[2] <= 256, "conv: TMA box too large (%u x %u)", box[1], box[2]);
    if (make_tensor_map_f16(&tmA, d->x, 4, dims, str, box, es, "activation")) return -1;
  }
  {
    uint64_t dims[3] = {(uint64_t)d->Cin, (uint64_t)d->Cout_pad, (uint64_t)d->taps};
    uint64_t str[2] = {(uint64_t)d->Cin * 2, (uint64_t)d->Cout_pad * d->Cin * 2};
    uint32_t box[3] = {64, (uint32_t)BLOCK_N, 1};
    uint32_t es[3] = {1, 1, 1};
    if (make_tensor_map_f16(&tmB, d->w, 3, dims, str, box, es, "weights")) return -1;
  }
  if (OUT_MODE == DREAMB200_OUT_NHWC_F16 && d->y_pool != nullptr) {
    const uint64_t Wp = (uint64_t)(d->Wo / 2), Hp = (uint64_t)(d->Ho / 2), C = (uint64_t)d->Cout_pad;
    uint64_t dims[4] = {C, Wp, Hp, (uint64_t)d->B};
    uint64_t str[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw / 2, (uint32_t)p.th / 2, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (make_tensor_map_f16(&tmP, d->y_pool, 4, dims, str, box, es, "pooled output")) return -1;
  }
  if (OUT_MODE == DREAMB200_OUT_NHWC_F16 && d->y != nullptr) {
    uint64_t dims[4] = {(uint64_t)d->Cout_pad, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->B};
    uint64_t str[3] = {(uint64_t)d->y_stride_w * 2, (uint64_t)d->y_stride_h * 2, (uint64_t)d->y_stride_b * 2};
    uint32_t box[4] = {64, (uint32_t)p.tw, (uint32_t)p.th, 1};
    uint32_t es[4] = {1, 1, 1, 1};
    if (make_tensor_map_f16(&tmC, d->y, 4, dims, str, box, es, "output")) return -1;
  }

  constexpr bool kNhwc = OUT_MODE == DREAMB200_OUT_NHWC_F16;
  const bool plain = kNhwc && d->residual == nullptr && d->residual_f32 == nullptr && d->y_f32 == nullptr &&
                     d->gate == nullptr && d->out_scale == nullptr && d->colsum == nullptr && d->absmax == nullptr;
  auto kern = plain ? conv_tc_kernel<BLOCK_N, OUT_MODE, kNhwc> : conv_tc_kernel<BLOCK_N, OUT_MODE, false>;
  static bool attr_set[2] = {false, false};  // per template instantiation
  if (!attr_set[plain]) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set[plain] = true;
  }
  int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  kern<<<grid, kTcThreads<OUT_MODE>(), smem_bytes, stream>>>(tmA, tmB, tmC, tmP, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;   // query failed (no device yet): B200 default, re-queried never -- launches fail anyway
  }
  return n;
}

}  // namespace db200

using namespace db200;

extern "C" double dreamb200_conv_tile_utilization(int Wo, int Ho, int in_stride, int even) {
  int tw, th;
  return choose_tile(Wo, Ho, in_stride, even != 0, &tw, &th);
}

extern "C" int dreamb200_conv2d_fwd(const dreamb200_conv_desc* d, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  DB_REQUIRE(d != nullptr, "conv: null descriptor");
  DB_REQUIRE(d->x && d->w && (d->y || d->y_pool), "conv: null tensor pointer");
  DB_REQUIRE(d->y_pool == nullptr || (d->out_mode == DREAMB200_OUT_NHWC_F16 && d->Ho >= 2 && d->Wo >= 2),
             "conv: fused pooling needs the NHWC_F16 mode and an output of at least 2x2");
  DB_REQUIRE(d->Cin > 0 && d->Cin % 64 == 0, "conv: Cin=%d must be a positive multiple of 64", d->Cin);
  DB_REQUIRE(d->taps >= 1 && d->taps <= DREAMB200_MAX_TAPS, "conv: taps=%d out of range", d->taps);
  DB_REQUIRE(d->in_stride == 1 || d->in_stride == 2, "conv: stride %d unsupported", d->in_stride);
  DB_REQUIRE(d->B > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0, "conv: empty tensor");
  DB_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->w & 15) == 0 && ((uintptr_t)d->y & 15) == 0 &&
                 ((uintptr_t)d->y_pool & 15) == 0,
             "conv: tensors must be 16-byte aligned");
  const int sms = device_sm_count();
  if (d->out_mode == DREAMB200_OUT_NCHW_F32) {
    DB_REQUIRE(d->Cout_pad == 16, "conv: NCHW_F32 head needs Cout_pad == 16 (got %d)", d->Cout_pad);
    DB_REQUIRE(d->cout_real >= 1 && d->cout_real <= 16, "conv: cout_real=%d out of range", d->cout_real);
    DB_REQUIRE(d->residual == nullptr && d->residual_f32 == nullptr && d->y_f32 == nullptr,
               "conv: residual / fp32 stream unsupported for the NCHW_F32 head");
    {
      const int r = try_conv_rs(d, stream);    // slab kernel when the head is a plain 3x3 over 64 (padded) channels
      if (r != 0) return r > 0 ? 0 : r;
    }
    return launch<16, DREAMB200_OUT_NCHW_F32>(d, stream, sms);
  }
  DB_REQUIRE(d->out_mode == DREAMB200_OUT_NHWC_F16, "conv: unknown out_mode %d", d->out_mode);
  DB_REQUIRE(d->Cout_pad > 0 && d->Cout_pad % 64 == 0, "conv: Cout_pad=%d must be a multiple of 64",
             d->Cout_pad);
  DB_REQUIRE((d->y_stride_w * 2) % 16 == 0 && (d->y_stride_h * 2) % 16 == 0 && (d->y_stride_b * 2) % 16 == 0,
             "conv: output strides must be multiples of 16 bytes");
  {
    int r = try_conv_tc2(d, stream);           // CTA-pair kernel for the wide layers (DREAMB200_TC2=0 switches it off)
    if (r != 0) return r > 0 ? 0 : r;
    r = try_conv_rs3(d, stream);               // two-output-rows slab kernel for 64 -> 64 channels (DREAMB200_RS3)
    if (r != 0) return r > 0 ? 0 : r;
    r = try_conv_rs2(d, stream);               // CTA-pair slab kernel (DREAMB200_RS2 bit mask)
    if (r != 0) return r > 0 ? 0 : r;
    r = try_conv_rs(d, stream);                // row-shared kernel for the narrow 3x3 layers
    if (r != 0) return r > 0 ? 0 : r;
  }
  if (d->Cout_pad % 256 == 0) return launch<256, DREAMB200_OUT_NHWC_F16>(d, stream, sms);
  if (d->Cout_pad % 128 == 0) return launch<128, DREAMB200_OUT_NHWC_F16>(d, stream, sms);
  return launch<64, DREAMB200_OUT_NHWC_F16>(d, stream, sms);
}

extern "C" int dreamb200_conv2d_fwd_phases(const dreamb200_conv_desc* descs, int n_phases, void* stream_v) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  DB_REQUIRE(descs != nullptr && n_phases >= 1 && n_phases <= 4, "conv phases: bad arguments (n_phases=%d)", n_phases);
  if (n_phases > 1) {
    bool ok = true;
    for (int ph = 0; ph < n_phases && ok; ++ph) {
      const dreamb200_conv_desc* d = descs + ph;
      ok = d->x && d->w && d->y && d->Cin > 0 && d->Cin % 64 == 0 && d->taps >= 1 && d->Cout_pad % 64 == 0 &&
           ((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->w & 15) == 0 && ((uintptr_t)d->y & 15) == 0 &&
           (d->y_stride_w * 2) % 16 == 0 && (d->y_stride_h * 2) % 16 == 0 && (d->y_stride_b * 2) % 16 == 0 &&
           (d->in_stride == 1 || d->in_stride == 2) && d->B > 0 && d->H > 0 && d->W > 0 && d->Ho > 0 && d->Wo > 0;
    }
    if (ok) {
      const int r = try_conv_tc2_phases(descs, n_phases, stream);
      if (r != 0) return r > 0 ? 0 : r;
    }
  }
  for (int ph = 0; ph < n_phases; ++ph) {          // not a group the pair kernel takes: one launch per phase
    const int rc = dreamb200_conv2d_fwd(descs + ph, stream_v);
    if (rc != 0) return rc;
  }
  return 0;
}
