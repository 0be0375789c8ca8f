// wgrad_first.cu -- weight gradient of the network's first convolution (3 -> 64 channels, 3x3, pad 1; reference:
// autograd of `layer_0_1_down.0`, dream/models.py:591-599, reached from `loss.backward()` in network.py:333):
//
//     dW[co][(r,s,c)] = sum over pixels p of dY[p][co] * x[p + (r-1, s-1)][c]
//
// a 64 x 27 result reduced over B*H*W pixels.  A materialised patch tensor (im2col) costs 128 B written + read per
// pixel; here, as in first_conv.cu, the 10x24 fp32 input patch of every 16x8 pixel tile is staged by TMA and four
// producer warps turn it into the [128 pixels][64 k-slots] fp16 tile (27 real slots) directly in shared memory --
// exactly the MN-major SWIZZLE_128B B operand the generic wgrad kernel (train_kernels.cu) would have fetched from
// the im2col tensor.  A = the dY tile [128 pixels][64 co] by TMA; M = 128 with the upper 64 rows pointed at a
// shared all-zero chunk.  fp32 accumulation in TMEM over this CTA's pixel range, fp32 atomics into dW at the end.
// HBM traffic: 128 B (dY) + 12 B (x) per pixel.
//
// Warps: 0 = TMA (dY tiles + input patches), 1 = MMA issuer + TMEM owner, 2-5 = producers, then the epilogue.
#include "common.cuh"
#include "dreamb200.h"

namespace db200 {

int make_tensor_map_f16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estride,
                        const char* what);
int make_tensor_map_plain(CUtensorMap* tm, int kind, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, const char* what);
int device_sm_count();

struct WgradFirstParams {
  int B, H, W;
  int tiles_x, tiles_y;
  long long kblocks_total;
  float* dw;            // fp32 [64][64]: dw[co][k], k = (r*3+s)*3+c < 27
  int stages;
};

constexpr int kWfThreads = 6 * 32;
constexpr int kWfTw = 16, kWfTh = 8;
constexpr int kWfChunk = 128 * 128;                       // [128 px][64 ch] fp16
constexpr int kWfPatchW = 24, kWfPatchH = kWfTh + 2;
constexpr uint32_t kWfPatchBytes = 3 * kWfPatchH * kWfPatchW * 4;   // 2880
constexpr uint32_t kWfPatchStride = 3072;
constexpr int kWfPatchStages = 4;

__global__ void __launch_bounds__(kWfThreads, 1)
wgrad_first_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ WgradFirstParams p) {
  constexpr uint32_t kIdesc = umma_idesc_f16_m128_mn(64);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int stages = p.stages;
  const uint32_t smem_zero = smem_base;                                  // 16 KB of zeros: rows 64..127 of A
  const uint32_t smem_ab = smem_zero + kWfChunk;                         // stages x (A 16 KB + B 16 KB)
  const uint32_t smem_in = smem_ab + stages * 2 * kWfChunk;              // input patches
  const uint32_t bar_base = smem_in + kWfPatchStages * kWfPatchStride;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };              // dY TMA transaction + 4 producer warps
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };  // MMA commit
  auto in_full = [&](int s) { return bar_base + 8u * (2 * stages + s); };
  auto in_empty = [&](int s) { return bar_base + 8u * (2 * stages + kWfPatchStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * stages + 2 * kWfPatchStages);
  const uint32_t tmem_ptr_smem = done_bar + 8u;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_smem - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero the shared zero chunk and every B buffer once: the producers only ever write k-slots 0..31 of a row
  {
    uint4* z = reinterpret_cast<uint4*>(smem_gen);
    const int n16 = (kWfChunk + stages * 2 * kWfChunk) / 16;
    for (int i = threadIdx.x; i < n16; i += kWfThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1 + 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < kWfPatchStages; ++s) {
      mbar_init(in_full(s), 1);
      mbar_init(in_empty(s), 4);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<64>(tmem_ptr_smem);
  fence_proxy_async_smem();                       // the zero fill must be visible to the tensor core's reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  const long long kb_lo = p.kblocks_total * blockIdx.x / gridDim.x;
  const long long kb_hi = p.kblocks_total * (blockIdx.x + 1) / gridDim.x;
  const int n_kb = (int)(kb_hi - kb_lo);
  const int tiles = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    {   // whole warp in uniform control flow, one elected lane issues (see elect_one in common.cuh)
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (long long kb = kb_lo; kb < kb_hi; ++kb, ++it) {
        const int b = (int)(kb / tiles);
        const int r = (int)(kb - (long long)b * tiles);
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int is = it % kWfPatchStages;
        mbar_wait(in_empty(is), (uint32_t)(((it / kWfPatchStages) & 1) ^ 1));
        if (elect_one()) {
          mbar_expect_tx(in_full(is), kWfPatchBytes);
          tma_load_4d(smem_in + is * kWfPatchStride, &tmX, in_full(is), tx * kWfTw - 4, ty * kWfTh - 1, 0, b);
        }
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), (uint32_t)kWfChunk);
          tma_load_4d(smem_ab + stage * 2 * kWfChunk, &tmDY, full_bar(stage), 0, tx * kWfTw, ty * kWfTh, b);
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    {   // whole warp in uniform control flow, one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_ab + stage * 2 * kWfChunk;
        // M chunk 1 (rows 64..127) = the zero chunk in front of the ring: leading-dimension offset is negative in
        // address terms, so describe A starting AT the zero chunk instead: rows 0..63 <- zeros, rows 64..127 <- dY
        const uint64_t adesc = umma_desc_mn_sw128(smem_zero, sa - smem_zero);
        const uint64_t bdesc = umma_desc_mn_sw128(sa + kWfChunk, kWfChunk);
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
  } else {
    // ===================== producers (warps 2..5): input patch -> B tile =====================
    // lanes 0-15 / 16-31 of a warp take tile rows two apart (see first_conv.cu: disjoint shared-memory banks)
    const int pw = (warp - 2) & 3, lx = lane & 15;
    const int ly = (pw >> 1) * 4 + (pw & 1) + 2 * ((lane >> 4) & 1);
    const int t = ly * kWfTw + lx;             // row of the B tile == pixel of the tile
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < n_kb; ++it) {
      const int is = it % kWfPatchStages;
      mbar_wait(in_full(is), (uint32_t)((it / kWfPatchStages) & 1));
      const float* patch = reinterpret_cast<const float*>(smem_gen + (smem_in - smem_base) + is * kWfPatchStride);
      float v[32];
#pragma unroll
      for (int k = 27; k < 32; ++k) v[k] = 0.0f;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int ss = 0; ss < 3; ++ss)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            v[(rr * 3 + ss) * 3 + c] = patch[(c * kWfPatchH + ly + rr) * kWfPatchW + lx + ss + 3];
      __syncwarp();
      if (lane == 0) mbar_arrive(in_empty(is));
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t row_addr = smem_ab + stage * 2 * kWfChunk + kWfChunk + (uint32_t)t * 128u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __half2 h0 = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]);
        __half2 h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
        __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
        __half2 h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
        const uint32_t dst = row_addr + (((uint32_t)j ^ (uint32_t)(t & 7)) * 16u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                     "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
                     "r"(*reinterpret_cast<uint32_t*>(&h3))
                     : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    // ===================== epilogue: rows 64..127 of the accumulator = co 0..63, columns = k-slots =====================
    if (n_kb > 0) {
      const int q = warp & 3;                    // TMEM lane quarter of this warp
      mbar_wait(done_bar, 0);
      tc_fence_after();
      if (q >= 2) {
        const int co = (q - 2) * 32 + lane;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16), v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 27; ++i) atomicAdd(p.dw + co * 64 + i, __uint_as_float(v[i]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}

}  // namespace db200

using namespace db200;

// dy fp16 NHWC [B,H,W,64] (the masked, scaled gradient of the first layer's output); x fp32 NCHW [B,3,H,W];
// dw fp32 [64][64], zeroed by the caller: dw[co][(r*3+s)*3+c] += sum_p dy[p][co] * x[p+(r-1,s-1)][c]
extern "C" int dreamb200_wgrad_first3x3(const void* dy, const float* x, float* dw, int B, int H, int W, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  DB_REQUIRE(dy && x && dw, "wgrad_first: null pointer");
  DB_REQUIRE(B > 0 && H > 0 && W > 0, "wgrad_first: empty input");
  DB_REQUIRE(W % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0,
             "wgrad_first: needs W %% 4 == 0 and 16-byte aligned tensors (use the im2col path otherwise)");
  WgradFirstParams p;
  p.B = B; p.H = H; p.W = W;
  p.tiles_x = (W + kWfTw - 1) / kWfTw;
  p.tiles_y = (H + kWfTh - 1) / kWfTh;
  p.kblocks_total = (long long)B * p.tiles_x * p.tiles_y;
  p.dw = dw;
  CUtensorMap tmDY, tmX;
  {
    const uint32_t es[4] = {1, 1, 1, 1};
    const uint32_t box[4] = {64, kWfTw, kWfTh, 1};
    uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    if (make_tensor_map_f16(&tmDY, dy, 4, dims, str, box, es, "wgrad_first dY")) return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)H * W * 12};
    uint32_t box[4] = {kWfPatchW, kWfPatchH, 3, 1};
    if (make_tensor_map_plain(&tmX, 0, x, 4, dims, str, box, "wgrad_first input")) return -1;
  }
  const int tail = kWfPatchStages * (int)kWfPatchStride + 512;
  int stages = (232448 - 1024 - kWfChunk - tail) / (2 * kWfChunk);
  if (stages > 6) stages = 6;
  p.stages = stages;
  const int smem_bytes = 1024 + kWfChunk + stages * 2 * kWfChunk + tail;
  static bool attr_set = false;
  if (!attr_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(wgrad_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  long long grid = device_sm_count();
  if (grid > p.kblocks_total) grid = p.kblocks_total;
  wgrad_first_kernel<<<(int)grid, kWfThreads, smem_bytes, stream>>>(tmDY, tmX, p);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
