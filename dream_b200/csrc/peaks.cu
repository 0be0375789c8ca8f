// peaks.cu -- device-side keypoint extraction from belief maps.
//
// dreamb200_peaks reproduces dream/image_proc.py:914-1018 (peaks_from_belief_maps) for all B*K maps
// of a batch in three launches, bit-for-bit on the integer peak set:
//   1. gauss_pass<0>: scipy.ndimage.gaussian_filter axis-0 pass (image_proc.py:935).  fp64 accumulation
//      in scipy's exact order -- centre tap first, then symmetric pairs from the farthest inwards,
//      (lo+hi)*w added with separate rounding (no FMA: scipy's C is built without contraction) --
//      "reflect" boundary, result rounded to fp32 exactly once, like scipy's fp32 output array.
//   2. gauss_pass<1>: the axis-1 pass on the fp32 intermediate.
//   3. collect_peaks: one CTA per map.  4-neighbour >= test against zero-padded shifts and > 0.01f
//      (image_proc.py:936-954), ordered (raster) compaction, 5x5 weighted centroid on the UNsmoothed
//      map in fp64 with numpy's pairwise summation order (image_proc.py:961-998), plus the
//      best / second-best scores needed by DreamNetwork.inference (network.py:548-577).
// All three are HBM/L2 streaming kernels over 4*B*K*h*w bytes.  They remain as the path for maps too large for
// shared memory (full-resolution decoders) or a Gaussian radius other than 12.
//
// peaks_fused_kernel (the default for sigma = 3 and maps up to ~160x160): ONE launch, one CTA per map.  The map is read
// from HBM once into shared memory; both Gaussian passes, the peak test, the ordered compaction and the centroids run
// there.  Same fp64 operations in the same order as the three-kernel path (bit-identical results), but every thread
// filters a STRIP of 10 consecutive outputs from a register window of 34 inputs, so each input is read from shared
// memory and converted to fp64 once per 10 outputs instead of once per tap (the three-kernel path was bound by a
// 25-long dependent chain with two global loads and two fp32->fp64 conversions per tap).  The intermediate and the
// smoothed map are stored transposed with an odd pitch so that both passes and the peak test are bank-conflict free.
// What bounds it: the fp64 pipe (2 * 37 fp64 operations per pixel), not HBM -- see DESIGN.md 3.4.
//
// dreamb200_softargmax: SoftArgmaxPavlo.forward (dream/spatial_softmax.py:24-95), one CTA per map.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "dreamb200.h"

namespace db200 {

int device_sm_count();

constexpr int kMaxRadius = 32;
struct GaussW {
  double w[kMaxRadius + 1];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i >= n ? period - 1 - i : i;
}

template <int AXIS>
__global__ void gauss_pass_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_maps, int h,
                                  int w, const __grid_constant__ GaussW gw, int radius) {
  const long long total = n_maps * h * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % w);
    const long long r = idx / w;
    const int y = (int)(r % h);
    const float* m = in + (r / h) * (long long)h * w;
    const int n = AXIS == 0 ? h : w;
    const int c = AXIS == 0 ? y : x;
    double acc = __dmul_rn((double)m[(size_t)y * w + x], gw.w[0]);
    for (int d = radius; d >= 1; --d) {
      int lo = c - d, hi = c + d;
      if (lo < 0 || hi >= n) {
        lo = reflect_idx(lo, n);
        hi = reflect_idx(hi, n);
      }
      const double a = AXIS == 0 ? (double)m[(size_t)lo * w + x] : (double)m[(size_t)y * w + lo];
      const double b = AXIS == 0 ? (double)m[(size_t)hi * w + x] : (double)m[(size_t)y * w + hi];
      acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), gw.w[d]));
    }
    out[idx] = __double2float_rn(acc);
  }
}

// numpy pairwise sum of a contiguous 25-vector: 8 partials over the first 24, tree combine, tail.
__device__ __forceinline__ double pairwise25(const double* v) {
  double r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = v[j];
#pragma unroll
  for (int i = 8; i < 24; i += 8)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], v[i + j]);
  const double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  return __dadd_rn(res, v[24]);
}

struct Top2 {
  float s0, s1;      // best and second-best score (NaN-free inputs assumed)
  int i0;            // raster index of the best (ties -> earliest), -1 if none
  int n;             // how many candidates merged (0, 1, 2+)
  double x0, y0;
};

__device__ __forceinline__ void top2_insert(Top2& t, float s, int idx, double x, double y) {
  if (t.n == 0) {
    t.s0 = s; t.i0 = idx; t.x0 = x; t.y0 = y; t.n = 1;
  } else if (s > t.s0 || (s == t.s0 && idx < t.i0)) {
    t.s1 = t.s0; t.s0 = s; t.i0 = idx; t.x0 = x; t.y0 = y; t.n = 2;
  } else if (t.n == 1 || s > t.s1) {
    t.s1 = s; t.n = 2;
  }
}

constexpr int kPeakThreads = 256;

__global__ void __launch_bounds__(kPeakThreads)
collect_peaks_kernel(const float* __restrict__ ori, const float* __restrict__ smooth, int h, int w,
                     double offset, int cap, double* __restrict__ peak_xy, float* __restrict__ peak_score,
                     int32_t* __restrict__ peak_ij, int32_t* __restrict__ counts, double* __restrict__ summary) {
  const long long map = blockIdx.x;
  const float* mo = ori + map * (long long)h * w;
  const float* ms = smooth + map * (long long)h * w;
  __shared__ int warp_cnt[kPeakThreads / 32];
  __shared__ int base_s;
  __shared__ Top2 tops[kPeakThreads];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) base_s = 0;
  Top2 mine;
  mine.n = 0; mine.i0 = -1; mine.s0 = 0.f; mine.s1 = 0.f; mine.x0 = 0.0; mine.y0 = 0.0;
  __syncthreads();
  const int total = h * w;
  for (int start = 0; start < total; start += kPeakThreads) {
    const int idx = start + tid;
    bool is_peak = false;
    int px = 0, py = 0;
    if (idx < total) {
      py = idx / w;
      px = idx - py * w;
      const float v = ms[idx];
      const float up = py > 0 ? ms[idx - w] : 0.0f;
      const float dn = py < h - 1 ? ms[idx + w] : 0.0f;
      const float lf = px > 0 ? ms[idx - 1] : 0.0f;
      const float rt = px < w - 1 ? ms[idx + 1] : 0.0f;
      is_peak = (v >= up) && (v >= dn) && (v >= lf) && (v >= rt) && (v > 0.01f);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, is_peak);
    if (lane == 0) warp_cnt[wid] = __popc(ballot);
    __syncthreads();
    int before = base_s;
    int chunk_total = 0;
#pragma unroll
    for (int i = 0; i < kPeakThreads / 32; ++i) {
      const int c = warp_cnt[i];
      if (i < wid) before += c;
      chunk_total += c;
    }
    if (is_peak) {
      const int slot = before + __popc(ballot & ((1u << lane) - 1u));
      double wts[25], xv[25], yv[25];
#pragma unroll
      for (int k = 0; k < 25; ++k) { wts[k] = 0.0; xv[k] = 0.0; yv[k] = 0.0; }
#pragma unroll
      for (int i = -2; i <= 2; ++i) {       // row offset
#pragma unroll
        for (int j = -2; j <= 2; ++j) {     // column offset
          const int yy = py + i, xx = px + j;
          if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
          const int k = (j + 2) * 5 + (i + 2);
          wts[k] = (double)mo[(size_t)yy * w + xx];
          xv[k] = (double)xx;
          yv[k] = (double)yy;
        }
      }
      const double scl = pairwise25(wts);
      double cx, cy;
      if (scl == 0.0) {
        cx = (double)px + offset;
        cy = (double)py + offset;
      } else {
#pragma unroll
        for (int k = 0; k < 25; ++k) { xv[k] = __dmul_rn(xv[k], wts[k]); yv[k] = __dmul_rn(yv[k], wts[k]); }
        cx = __dadd_rn(__ddiv_rn(pairwise25(xv), scl), offset);
        cy = __dadd_rn(__ddiv_rn(pairwise25(yv), scl), offset);
      }
      const float score = mo[idx];
      if (slot < cap) {
        const size_t o = (size_t)map * cap + slot;
        peak_xy[2 * o] = cx;
        peak_xy[2 * o + 1] = cy;
        peak_score[o] = score;
        peak_ij[2 * o] = px;
        peak_ij[2 * o + 1] = py;
      }
      top2_insert(mine, score, idx, cx, cy);
    }
    __syncthreads();
    if (tid == 0) base_s += chunk_total;
    __syncthreads();
  }
  tops[tid] = mine;
  __syncthreads();
  if (tid == 0) {
    Top2 t = tops[0];
    for (int i = 1; i < kPeakThreads; ++i) {
      const Top2& o = tops[i];
      if (o.n == 0) continue;
      const float o_s1 = o.s1;
      const int o_n = o.n;
      top2_insert(t, o.s0, o.i0, o.x0, o.y0);
      if (o_n >= 2) {
        // second of the other set can only become our second (index irrelevant for the runner-up)
        if (t.n == 1 || o_s1 > t.s1) { t.s1 = o_s1; t.n = 2; }
      }
    }
    counts[map] = base_s;
    summary[4 * map + 0] = t.x0;
    summary[4 * map + 1] = t.y0;
    summary[4 * map + 2] = (double)t.s0;
    summary[4 * map + 3] = (double)t.s1;
  }
}

// ---------------------------------------------------------------------------------------------
// fused path: one CTA per map, everything in shared memory
// ---------------------------------------------------------------------------------------------
constexpr int kFusedRadius = 12;
constexpr int kStrip = 10;
constexpr int kFusedThreads = 256;
constexpr int kMaxMaskWords = 1024;          // h*w <= 32768 pixels

// fp32 <-> fp64 conversions on the integer pipe.  On sm_100 F2F.F64.F32 / F2F.F32.F64 issue on the XU pipe at a few
// threads per clock per SM: ncu showed the first version of this kernel with the XU pipe saturated (118 % of its
// nominal peak) and the fp64 pipe 24 % busy -- the conversions, not the 74 fp64 operations per pixel, set its 190 us.
// A second version branched to the hardware instruction per value for the rare non-normal cases: 44 branch /
// reconvergence pairs per task and a 160 KB unrolled body made it instruction-fetch bound (stall_no_inst on top).
// Now both conversions are BRANCH-FREE bit manipulation, exact for normal numbers and zero (widening is always
// exact; narrowing rounds to nearest even like the hardware), and each accumulates a flag when it meets anything else
// (fp32 subnormal, inf, nan; a result outside the normal fp32 range).  A flagged task is simply recomputed by
// gauss_task_exact with the hardware conversions -- one well-predicted branch per 10 outputs.
__device__ __forceinline__ double f2d_fast(float f, uint32_t& flag) {
  const uint32_t b = __float_as_uint(f);
  const uint32_t a = b & 0x7fffffffu;
  const bool normal = a - 0x00800000u < 0x7f000000u;                       // exponent field 1..254
  flag |= normal ? 0u : a;                                                 // non-zero and not normal -> exact path
  const uint32_t hi = (b & 0x80000000u) | (normal ? (a >> 3) + 0x38000000u : 0u);    // re-bias 127 -> 1023
  return __hiloint2double((int)hi, (int)(b << 29));
}
__device__ __forceinline__ float d2f_fast(double d, uint32_t& flag) {
  const uint32_t hi = (uint32_t)__double2hiint(d), lo = (uint32_t)__double2loint(d);
  const uint32_t a = hi & 0x7fffffffu;
  const bool normal = a - 0x38100000u < 0x0fe00000u;                       // exponent field 897..1150
  flag |= normal ? 0u : (a | lo);                                          // non-zero and outside -> exact path
  uint32_t f = ((a - 0x38000000u) << 3) | (lo >> 29);
  const uint32_t rem = lo & 0x1fffffffu;
  f += (rem > 0x10000000u || (rem == 0x10000000u && (f & 1u))) ? 1u : 0u; // a carry into the exponent is correct
  return __uint_as_float((hi & 0x80000000u) | (normal ? f : 0u));
}

// Reference form of one task (the arithmetic of gauss_pass_kernel: hardware conversions, general reflection): used for
// flagged tasks and for axes shorter than the register window.
static __device__ __noinline__ void gauss_task_exact(const float* __restrict__ line, int s_axis,
                                                     float* __restrict__ out_line, int d_axis, int c0, int c_end,
                                                     int n_axis, const GaussW& gw) {
  for (int j = 0; j < kStrip && c0 + j < c_end; ++j) {
    const int c = c0 + j;
    double acc = __dmul_rn((double)line[c * s_axis], gw.w[0]);
    for (int d = kFusedRadius; d >= 1; --d) {
      const int lo = reflect_idx(c - d, n_axis), hi = reflect_idx(c + d, n_axis);
      acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)line[lo * s_axis], (double)line[hi * s_axis]), gw.w[d]));
    }
    out_line[c * d_axis] = __double2float_rn(acc);
  }
}

// One separable pass over a [n_other lines] x [n_axis samples] map.  Element (line o, sample i) of the source is
// src[o*s_other + i*s_axis]; thread tasks are (strip, line) with the line index fastest, so that consecutive lanes
// touch consecutive lines (the caller picks layouts where that is conflict free).  Arithmetic = gauss_pass_kernel's.
// Outputs are produced for the axis positions [out_begin, out_begin + out_count) and stored at dst index
// (position - out_begin); the whole-map kernel asks for all of them, a band only for its rows (its halo rows are part
// of the source, so nothing is reflected there).  Out of line: both passes share one copy of the unrolled body.
static __device__ __noinline__ void gauss_strips(const float* __restrict__ src, int s_axis, int s_other,
                                                 float* __restrict__ dst, int d_axis, int d_other, int n_axis,
                                                 int n_other, int out_begin, int out_count, const GaussW& gw) {
  const int strips = (out_count + kStrip - 1) / kStrip;
  const int tasks = strips * n_other;
  const int out_end = out_begin + out_count;
  const uint32_t magic = 0xffffffffu / (uint32_t)n_other + 1u;      // t / n_other == umulhi(t, magic) for t*n_other < 2^32
  // the register window reflects an index at most once per side: enough when every index it can touch, -12 on the
  // left and (last output position + 9 + 12) on the right, lands inside [0, n_axis) after one reflection
  const bool windowed = n_axis > kFusedRadius && 2 * n_axis - 1 >= out_end + kStrip - 1 + kFusedRadius;
  for (int t = threadIdx.x; t < tasks; t += kFusedThreads) {
    const int strip = (int)__umulhi((uint32_t)t, magic);
    const int o = t - strip * n_other;
    const int c0 = out_begin + strip * kStrip;
    const float* line = src + o * s_other;
    float* out_line = dst + o * d_other - out_begin * d_axis;       // indexed by the axis position
    uint32_t flag = windowed ? 0u : 1u;
    if (windowed) {
      double v[kStrip + 2 * kFusedRadius];
#pragma unroll
      for (int j = 0; j < kStrip + 2 * kFusedRadius; ++j) {
        int i = c0 - kFusedRadius + j;
        i = i < 0 ? -1 - i : i;                                     // "reflect": d c b a | a b c d | d c b a
        i = i >= n_axis ? 2 * n_axis - 1 - i : i;
        v[j] = f2d_fast(line[i * s_axis], flag);
      }
#pragma unroll
      for (int j = 0; j < kStrip; ++j) {
        double acc = __dmul_rn(v[j + kFusedRadius], gw.w[0]);
#pragma unroll
        for (int d = kFusedRadius; d >= 1; --d)
          acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v[j + kFusedRadius - d], v[j + kFusedRadius + d]), gw.w[d]));
        const float r = d2f_fast(acc, flag);
        if (c0 + j < out_end) out_line[(c0 + j) * d_axis] = r;
      }
    }
    if (flag != 0u) gauss_task_exact(line, s_axis, out_line, d_axis, c0, out_end, n_axis, gw);
  }
}

// 5x5 weighted centroid around (px, py) on the unsmoothed map (image_proc.py:961-998); out of line: peaks are rare and
// its 75 fp64 temporaries must not squeeze the register budget of the filter loops.
static __device__ __noinline__ void peak_centroid(const float* __restrict__ mo, int h, int w, int px, int py,
                                                  double offset, double* cx_out, double* cy_out) {
  double wts[25], xv[25], yv[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) { wts[k] = 0.0; xv[k] = 0.0; yv[k] = 0.0; }
#pragma unroll
  for (int i = -2; i <= 2; ++i) {       // row offset
#pragma unroll
    for (int j = -2; j <= 2; ++j) {     // column offset
      const int yy = py + i, xx = px + j;
      if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
      const int k = (j + 2) * 5 + (i + 2);
      wts[k] = (double)__ldg(mo + yy * w + xx);
      xv[k] = (double)xx;
      yv[k] = (double)yy;
    }
  }
  const double scl = pairwise25(wts);
  double cx, cy;
  if (scl == 0.0) {
    cx = (double)px + offset;
    cy = (double)py + offset;
  } else {
#pragma unroll
    for (int k = 0; k < 25; ++k) { xv[k] = __dmul_rn(xv[k], wts[k]); yv[k] = __dmul_rn(yv[k], wts[k]); }
    cx = __dadd_rn(__ddiv_rn(pairwise25(xv), scl), offset);
    cy = __dadd_rn(__ddiv_rn(pairwise25(yv), scl), offset);
  }
  *cx_out = cx;
  *cy_out = cy;
}

// combined best / runner-up of two candidate sets (commutative: multiset semantics, ties -> earliest raster index)
__device__ __forceinline__ void top2_merge(Top2& t, const Top2& o) {
  if (o.n == 0) return;
  const float o_s1 = o.s1;
  const int o_n = o.n;
  top2_insert(t, o.s0, o.i0, o.x0, o.y0);
  if (o_n >= 2 && (t.n == 1 || o_s1 > t.s1)) { t.s1 = o_s1; t.n = 2; }
}

__device__ __forceinline__ Top2 top2_shfl_xor(const Top2& t, int m) {
  Top2 o;
  o.s0 = __shfl_xor_sync(0xffffffffu, t.s0, m);
  o.s1 = __shfl_xor_sync(0xffffffffu, t.s1, m);
  o.i0 = __shfl_xor_sync(0xffffffffu, t.i0, m);
  o.n = __shfl_xor_sync(0xffffffffu, t.n, m);
  o.x0 = __shfl_xor_sync(0xffffffffu, t.x0, m);
  o.y0 = __shfl_xor_sync(0xffffffffu, t.y0, m);
  return o;
}

__global__ void __launch_bounds__(kFusedThreads, 2)
peaks_fused_kernel(const float* __restrict__ maps, int n_maps, int h, int w, const __grid_constant__ GaussW gw,
                   double offset, int cap, double* __restrict__ peak_xy, float* __restrict__ peak_score,
                   int32_t* __restrict__ peak_ij, int32_t* __restrict__ counts, double* __restrict__ summary,
                   float* __restrict__ smooth_out) {
  extern __shared__ __align__(16) float fsm[];
  const int hp = h | 1;                        // odd pitch of the transposed buffers
  const int total = h * w;
  float* bufA = fsm;                           // the map, row major [h][w]; later the smoothed map, transposed [w][hp]
  float* bufB = fsm + w * hp;                  // axis-0 result, transposed [w][hp]
  __shared__ uint32_t mask[kMaxMaskWords];
  __shared__ int pref[kMaxMaskWords];
  __shared__ Top2 warp_top[kFusedThreads / 32];
  __shared__ int total_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int words = (total + 31) >> 5;

  for (int map = blockIdx.x; map < n_maps; map += gridDim.x) {
    const float* mo = maps + (long long)map * total;
    // ---- HBM -> shared, once
    if ((total & 3) == 0 && (reinterpret_cast<uintptr_t>(mo) & 15) == 0) {
      const float4* m4 = reinterpret_cast<const float4*>(mo);
      float4* a4 = reinterpret_cast<float4*>(bufA);
      for (int i = tid; i < (total >> 2); i += kFusedThreads) a4[i] = __ldg(m4 + i);
    } else {
      for (int i = tid; i < total; i += kFusedThreads) bufA[i] = __ldg(mo + i);
    }
    __syncthreads();
    // ---- axis 0 (along y): lines = columns x; reads bufA[y*w + x], writes bufB[x*hp + y]
    gauss_strips(bufA, w, 1, bufB, 1, hp, h, w, 0, h, gw);
    __syncthreads();
    // ---- axis 1 (along x): lines = rows y; reads bufB[x*hp + y], writes bufA[x*hp + y]
    gauss_strips(bufB, hp, 1, bufA, hp, 1, w, h, 0, w, gw);
    __syncthreads();
    // ---- peak test -> bitmask in raster order
    const float* sm = bufA;
    const uint32_t wmagic = 0xffffffffu / (uint32_t)w + 1u;   // idx / w == umulhi(idx, wmagic): idx * w < 2^30
    if (smooth_out != nullptr) {                              // dreamb200_gaussian_smooth: dump the filtered map, done
      for (int idx = tid; idx < total; idx += kFusedThreads) {
        const int py = (int)__umulhi((uint32_t)idx, wmagic), px = idx - py * w;
        smooth_out[(long long)map * total + idx] = sm[px * hp + py];
      }
      __syncthreads();
      continue;
    }
    for (int base = wid * 32; base < words * 32; base += kFusedThreads) {
      const int idx = base + lane;
      bool is_peak = false;
      if (idx < total) {
        const int py = (int)__umulhi((uint32_t)idx, wmagic), px = idx - py * w;
        const float* c = sm + px * hp + py;
        const float v = c[0];
        const float up = py > 0 ? c[-1] : 0.0f;
        const float dn = py < h - 1 ? c[1] : 0.0f;
        const float lf = px > 0 ? c[-hp] : 0.0f;
        const float rt = px < w - 1 ? c[hp] : 0.0f;
        is_peak = (v >= up) && (v >= dn) && (v >= lf) && (v >= rt) && (v > 0.01f);
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, is_peak);
      if (lane == 0) mask[base >> 5] = ballot;
    }
    __syncthreads();
    // ---- exclusive prefix of the per-word counts (one warp)
    if (wid == 0) {
      int running = 0;
      for (int b = 0; b < words; b += 32) {
        const int c = b + lane < words ? __popc(mask[b + lane]) : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += n;
        }
        if (b + lane < words) pref[b + lane] = running + inc - c;
        running += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) total_s = running;
    }
    __syncthreads();
    // ---- centroids of the peaks (few per map): one thread per mask word
    Top2 mine;
    mine.n = 0; mine.i0 = -1; mine.s0 = 0.f; mine.s1 = 0.f; mine.x0 = 0.0; mine.y0 = 0.0;
    for (int wd = tid; wd < words; wd += kFusedThreads) {
      uint32_t bits = mask[wd];
      int slot = pref[wd];
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int idx = wd * 32 + b;
        const int py = (int)__umulhi((uint32_t)idx, wmagic), px = idx - py * w;
        double cx, cy;
        peak_centroid(mo, h, w, px, py, offset, &cx, &cy);
        const float score = __ldg(mo + idx);
        if (slot < cap) {
          const size_t o = (size_t)map * cap + slot;
          peak_xy[2 * o] = cx;
          peak_xy[2 * o + 1] = cy;
          peak_score[o] = score;
          peak_ij[2 * o] = px;
          peak_ij[2 * o + 1] = py;
        }
        ++slot;
        top2_insert(mine, score, idx, cx, cy);
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const Top2 o = top2_shfl_xor(mine, m);
      top2_merge(mine, o);
    }
    if (lane == 0) warp_top[wid] = mine;
    __syncthreads();
    if (tid == 0) {
      Top2 t = warp_top[0];
      for (int i = 1; i < kFusedThreads / 32; ++i) top2_merge(t, warp_top[i]);
      counts[map] = total_s;
      summary[4 * map + 0] = t.x0;
      summary[4 * map + 1] = t.y0;
      summary[4 * map + 2] = (double)t.s0;
      summary[4 * map + 3] = (double)t.s1;
    }
    __syncthreads();                           // bufA / mask / warp_top are reused by the next map
  }
}

// ---------------------------------------------------------------------------------------------
// banded path: maps too large for one CTA's shared memory (resnet-H 208x208, full-resolution decoders 400x400 /
// 480x640) are cut into bands of `hb` rows; a CTA owns (map, band).
// ---------------------------------------------------------------------------------------------
// It stages the band's rows plus a halo of 13 rows above and below (reflected at the image border exactly like
// scipy's line extension), filters rows y0-1 .. y1 along y and then along x in shared memory (the extra row on
// either side is what the 4-neighbour test of the band's first / last row looks at), runs the same peak test /
// ordered compaction / centroids as the whole-map kernel on its rows and leaves its peaks in a staging record.
// The CTA that draws the LAST ticket of a map concatenates the bands' records in band order (= raster order)
// into the peak table and merges their best / runner-up summaries.  One launch, fixed order, no sorting.
struct BandHeader {
  int count;        // peaks found in the band (may exceed the staged capacity)
  int pad;
  Top2 top;         // best / runner-up of the band (raster index i0 is the global one)
};
struct BandPeak {
  double x, y;
  float score;
  int px, py;
  int pad;
};
struct BandParams {
  int n_maps, h, w, hb, n_bands, cap;
  double offset;
  unsigned char* stage;      // [n_maps][n_bands] records of (BandHeader + cap * BandPeak)
  unsigned* tickets;         // [n_maps], zero on entry
  double* peak_xy;
  float* peak_score;
  int32_t* peak_ij;
  int32_t* counts;
  double* summary;
};

__global__ void __launch_bounds__(kFusedThreads, 2)
peaks_banded_kernel(const float* __restrict__ maps, const __grid_constant__ GaussW gw,
                    const __grid_constant__ BandParams P) {
  extern __shared__ __align__(16) float fsm[];
  const int h = P.h, w = P.w;
  const int map = blockIdx.x / P.n_bands, band = blockIdx.x - map * P.n_bands;
  const int y0 = band * P.hb;
  const int rows = min(P.hb, h - y0);
  const int rows_in = rows + 2 * kFusedRadius + 2;            // global rows y0-13 .. y0+rows+12
  const int rows_out = rows + 2;                              // smoothed rows y0-1 .. y0+rows
  const int hp = rows_out | 1;
  float* bufA = fsm;                                          // [rows_in][w]; later the smoothed rows, transposed [w][hp]
  float* bufB = fsm + max(rows_in * w, w * hp);               // axis-0 result, transposed [w][hp]
  __shared__ uint32_t mask[kMaxMaskWords];
  __shared__ int pref[kMaxMaskWords];
  __shared__ Top2 warp_top[kFusedThreads / 32];
  __shared__ int total_s;
  __shared__ bool last_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float* mo = maps + (long long)map * h * w;
  // ---- rows (reflected at the image border) -> shared
  for (int r = wid; r < rows_in; r += kFusedThreads / 32) {
    const int gy = reflect_idx(y0 - kFusedRadius - 1 + r, h);
    const float* srow = mo + (long long)gy * w;
    for (int x = lane; x < w; x += 32) bufA[r * w + x] = __ldg(srow + x);
  }
  __syncthreads();
  // ---- axis 0: staged row r is global row y0-13+r, so output row l (global y0-1+l) is centred on staged row l+12
  gauss_strips(bufA, w, 1, bufB, 1, hp, rows_in, w, kFusedRadius, rows_out, gw);
  __syncthreads();
  // ---- axis 1: full rows, reflected at x = 0 / w like the whole-map kernel
  gauss_strips(bufB, hp, 1, bufA, hp, 1, w, rows_out, 0, w, gw);
  __syncthreads();
  // ---- peak test on the band's rows (local row l = 1 .. rows), bitmask in raster order
  const float* sm = bufA;
  const int total = rows * w;
  const int words = (total + 31) >> 5;
  const uint32_t wmagic = 0xffffffffu / (uint32_t)w + 1u;
  for (int base = wid * 32; base < words * 32; base += kFusedThreads) {
    const int idx = base + lane;
    bool is_peak = false;
    if (idx < total) {
      const int ly = (int)__umulhi((uint32_t)idx, wmagic), px = idx - ly * w;
      const int py = y0 + ly;
      const float* c = sm + px * hp + ly + 1;
      const float v = c[0];
      const float up = py > 0 ? c[-1] : 0.0f;
      const float dn = py < h - 1 ? c[1] : 0.0f;
      const float lf = px > 0 ? c[-hp] : 0.0f;
      const float rt = px < w - 1 ? c[hp] : 0.0f;
      is_peak = (v >= up) && (v >= dn) && (v >= lf) && (v >= rt) && (v > 0.01f);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, is_peak);
    if (lane == 0) mask[base >> 5] = ballot;
  }
  __syncthreads();
  if (wid == 0) {
    int running = 0;
    for (int b = 0; b < words; b += 32) {
      const int c = b + lane < words ? __popc(mask[b + lane]) : 0;
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      if (b + lane < words) pref[b + lane] = running + inc - c;
      running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) total_s = running;
  }
  __syncthreads();
  const size_t rec_bytes = sizeof(BandHeader) + (size_t)P.cap * sizeof(BandPeak);
  unsigned char* rec = P.stage + ((size_t)map * P.n_bands + band) * rec_bytes;
  BandPeak* staged = reinterpret_cast<BandPeak*>(rec + sizeof(BandHeader));
  Top2 mine;
  mine.n = 0; mine.i0 = -1; mine.s0 = 0.f; mine.s1 = 0.f; mine.x0 = 0.0; mine.y0 = 0.0;
  for (int wd = tid; wd < words; wd += kFusedThreads) {
    uint32_t bits = mask[wd];
    int slot = pref[wd];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int idx = wd * 32 + b;
      const int ly = (int)__umulhi((uint32_t)idx, wmagic), px = idx - ly * w;
      const int py = y0 + ly;
      double cx, cy;
      peak_centroid(mo, h, w, px, py, P.offset, &cx, &cy);
      const float score = __ldg(mo + py * w + px);
      if (slot < P.cap) {
        BandPeak bp;
        bp.x = cx; bp.y = cy; bp.score = score; bp.px = px; bp.py = py; bp.pad = 0;
        staged[slot] = bp;
      }
      ++slot;
      top2_insert(mine, score, py * w + px, cx, cy);
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const Top2 o = top2_shfl_xor(mine, m);
    top2_merge(mine, o);
  }
  if (lane == 0) warp_top[wid] = mine;
  __syncthreads();
  if (tid == 0) {
    Top2 t = warp_top[0];
    for (int i = 1; i < kFusedThreads / 32; ++i) top2_merge(t, warp_top[i]);
    BandHeader* hd = reinterpret_cast<BandHeader*>(rec);
    hd->count = total_s;
    hd->pad = 0;
    hd->top = t;
  }
  // ---- the last band of the map to finish assembles the map's rows of the peak table
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = atomicAdd(P.tickets + map, 1u) == (unsigned)P.n_bands - 1u;
  __syncthreads();
  if (!last_s) return;
  __threadfence();
  const unsigned char* recs = P.stage + (size_t)map * P.n_bands * rec_bytes;
  if (wid == 0) {
    // exclusive prefix of the band counts into pref[] (n_bands <= kMaxMaskWords), total into total_s
    int running = 0;
    for (int b = 0; b < P.n_bands; b += 32) {
      int c = 0;
      if (b + lane < P.n_bands) c = __ldcg(&reinterpret_cast<const BandHeader*>(recs + (size_t)(b + lane) * rec_bytes)->count);
      int inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      if (b + lane < P.n_bands) pref[b + lane] = running + inc - c;
      running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) total_s = running;
  }
  __syncthreads();
  for (int b = wid; b < P.n_bands; b += kFusedThreads / 32) {
    const unsigned char* r = recs + (size_t)b * rec_bytes;
    const int cnt = min(__ldcg(&reinterpret_cast<const BandHeader*>(r)->count), P.cap);
    const BandPeak* sp = reinterpret_cast<const BandPeak*>(r + sizeof(BandHeader));
    for (int i = lane; i < cnt; i += 32) {
      const int slot = pref[b] + i;
      if (slot >= P.cap) break;
      const size_t o = (size_t)map * P.cap + slot;
      const double x = __ldcg(&sp[i].x), y = __ldcg(&sp[i].y);
      P.peak_xy[2 * o] = x;
      P.peak_xy[2 * o + 1] = y;
      P.peak_score[o] = __ldcg(&sp[i].score);
      P.peak_ij[2 * o] = __ldcg(&sp[i].px);
      P.peak_ij[2 * o + 1] = __ldcg(&sp[i].py);
    }
  }
  if (tid == 0) {
    Top2 t;
    t.n = 0; t.i0 = -1; t.s0 = 0.f; t.s1 = 0.f; t.x0 = 0.0; t.y0 = 0.0;
    for (int b = 0; b < P.n_bands; ++b) {
      const BandHeader* hd = reinterpret_cast<const BandHeader*>(recs + (size_t)b * rec_bytes);
      Top2 o;
      o.s0 = __ldcg(&hd->top.s0); o.s1 = __ldcg(&hd->top.s1); o.i0 = __ldcg(&hd->top.i0); o.n = __ldcg(&hd->top.n);
      o.x0 = __ldcg(&hd->top.x0); o.y0 = __ldcg(&hd->top.y0);
      top2_merge(t, o);
    }
    P.counts[map] = total_s;
    P.summary[4 * map + 0] = t.x0;
    P.summary[4 * map + 1] = t.y0;
    P.summary[4 * map + 2] = (double)t.s0;
    P.summary[4 * map + 3] = (double)t.s1;
  }
}

// band height for a map of h x w: as tall as ~100 KB of shared memory allows (two CTAs per SM), but short enough that
// the launch has at least two CTAs per SM when the batch is small; 0 = the map is too wide for the banded kernel
static int band_rows(int n_maps, int h, int w, size_t* smem_out) {
  auto smem = [&](int hb) {
    const int rows_in = hb + 2 * kFusedRadius + 2, hp = (hb + 2) | 1;
    return (size_t)(std::max(rows_in * w, w * hp) + w * hp) * sizeof(float);
  };
  int hb = 0;
  for (int c = 1; c <= h && c <= 96; ++c)
    if (smem(c) <= 100 * 1024 && (long long)c * w <= 32LL * kMaxMaskWords) hb = c;
  if (hb == 0) return 0;
  const int want_ctas = 2 * device_sm_count();
  while (hb >= 16 && (long long)n_maps * ((h + hb - 1) / hb) < want_ctas) hb = (hb + 1) / 2;
  if ((h + hb - 1) / hb > kMaxMaskWords) return 0;
  *smem_out = smem(hb);
  return hb;
}
static size_t band_stage_bytes(int n_maps, int n_bands, int cap) {
  return (size_t)n_maps * n_bands * (sizeof(BandHeader) + (size_t)cap * sizeof(BandPeak)) + (size_t)n_maps * 4 + 64;
}

static size_t fused_smem_bytes(int h, int w) { return (size_t)2 * w * (h | 1) * sizeof(float); }
static int fused_set_smem(size_t smem) {
  static size_t smem_set = 0;
  if (smem > smem_set) {
    DB_CHECK_CUDA(cudaFuncSetAttribute(peaks_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  return 0;
}
static bool fused_ok(int h, int w, int radius) {
  return radius == kFusedRadius && (long long)h * w <= 32LL * kMaxMaskWords && fused_smem_bytes(h, w) <= 200 * 1024;
}

// ---------------------------------------------------------------------------------------------
// SoftArgmaxPavlo: 7x7 average pool (zero padded, /49) -> max -> exp(beta*(v-max)) -> expected x,y
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, sh[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256)
softargmax_kernel(const float* __restrict__ maps, const float* __restrict__ beta, float* __restrict__ out_xy,
                  int K, int H, int W, float* __restrict__ scratch) {
  const long long map = blockIdx.x;
  const float* m = maps + map * (long long)H * W;
  float* pooled = scratch + map * (long long)H * W;
  __shared__ float shf[8];
  __shared__ double shd[8];
  const float b = beta[map % K];
  float vmax = -INFINITY;
  for (int idx = threadIdx.x; idx < H * W; idx += blockDim.x) {
    const int y = idx / W, x = idx - y * W;
    float s = 0.0f;
    for (int dy = -3; dy <= 3; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      for (int dx = -3; dx <= 3; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        s += m[(size_t)yy * W + xx];
      }
    }
    s = s / 49.0f;
    pooled[idx] = s;
    vmax = fmaxf(vmax, s);
  }
  vmax = block_reduce_max(vmax, shf);
  double se = 0.0, sx = 0.0, sy = 0.0;
  for (int idx = threadIdx.x; idx < H * W; idx += blockDim.x) {
    const int y = idx / W, x = idx - y * W;
    const float e = expf(b * (pooled[idx] - vmax));
    se += (double)e;
    sx += (double)e * x;
    sy += (double)e * y;
  }
  se = block_reduce_sum(se, shd);
  sx = block_reduce_sum(sx, shd);
  sy = block_reduce_sum(sy, shd);
  if (threadIdx.x == 0) {
    const double den = se + 1e-8;
    out_xy[2 * map] = (float)(sx / den);
    out_xy[2 * map + 1] = (float)(sy / den);
  }
}

// 0 = whole map per CTA (peaks_fused_kernel), 1 = bands (peaks_banded_kernel), 2 = the three generic kernels.
// DREAMB200_PEAKS_UNFUSED forces 2, DREAMB200_PEAKS_BANDED prefers 1 even where 0 would fit (tests).
static int peaks_mode(int n_maps, int h, int w, int radius, int* hb, size_t* band_smem) {
  if (getenv("DREAMB200_PEAKS_UNFUSED") || radius != kFusedRadius) return 2;
  const bool want_banded = getenv("DREAMB200_PEAKS_BANDED") != nullptr;
  if (fused_ok(h, w, radius) && !want_banded) return 0;
  *hb = band_rows(n_maps, h, w, band_smem);
  if (*hb > 0) return 1;
  return fused_ok(h, w, radius) ? 0 : 2;
}

}  // namespace db200

using namespace db200;

extern "C" int dreamb200_peaks(const float* maps, int n_maps, int h, int w, const double* gauss_w, int radius,
                               double offset, float* scratch, int cap, double* peak_xy, float* peak_score,
                               int32_t* peak_ij, int32_t* counts, double* summary, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  DB_REQUIRE(maps && gauss_w && peak_xy && peak_score && peak_ij && counts && summary,
             "peaks: null pointer");
  DB_REQUIRE(n_maps > 0 && h > 0 && w > 0, "peaks: empty input (n_maps=%d h=%d w=%d)", n_maps, h, w);
  DB_REQUIRE(radius >= 0 && radius <= kMaxRadius, "peaks: radius %d out of range", radius);
  DB_REQUIRE(cap >= 1, "peaks: cap must be >= 1");
  DB_REQUIRE((long long)h * w < (1ll << 30), "peaks: map too large");
  GaussW gw;
  for (int i = 0; i <= kMaxRadius; ++i) gw.w[i] = i <= radius ? gauss_w[i] : 0.0;
  size_t band_smem = 0;
  int hb = 0;
  const int mode = peaks_mode(n_maps, h, w, radius, &hb, &band_smem);
  if (mode == 0) {
    const size_t smem = fused_smem_bytes(h, w);
    if (fused_set_smem(smem)) return -2;
    peaks_fused_kernel<<<n_maps, kFusedThreads, smem, stream>>>(maps, n_maps, h, w, gw, offset, cap, peak_xy,
                                                                 peak_score, peak_ij, counts, summary, nullptr);
    DB_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
  }
  if (mode == 1) {
    DB_REQUIRE(scratch, "peaks: the banded kernel needs the scratch buffer (dreamb200_peaks_plan)");
    DB_REQUIRE(((uintptr_t)scratch & 7) == 0, "peaks: scratch must be 8-byte aligned");
    BandParams P;
    P.n_maps = n_maps; P.h = h; P.w = w; P.hb = hb; P.n_bands = (h + hb - 1) / hb; P.cap = cap; P.offset = offset;
    const size_t rec_bytes = sizeof(BandHeader) + (size_t)cap * sizeof(BandPeak);
    P.stage = reinterpret_cast<unsigned char*>(scratch);
    P.tickets = reinterpret_cast<unsigned*>(P.stage + (size_t)n_maps * P.n_bands * rec_bytes);
    P.peak_xy = peak_xy; P.peak_score = peak_score; P.peak_ij = peak_ij; P.counts = counts; P.summary = summary;
    static size_t band_smem_set = 0;
    if (band_smem > band_smem_set) {
      DB_CHECK_CUDA(cudaFuncSetAttribute(peaks_banded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)band_smem));
      band_smem_set = band_smem;
    }
    DB_CHECK_CUDA(cudaMemsetAsync(P.tickets, 0, (size_t)n_maps * sizeof(unsigned), stream));
    peaks_banded_kernel<<<n_maps * P.n_bands, kFusedThreads, band_smem, stream>>>(maps, gw, P);
    DB_CHECK_CUDA(cudaGetLastError());
    count_launch(1);
    return 0;
  }
  DB_REQUIRE(scratch, "peaks: this map size / radius takes the three-kernel path and needs the scratch buffer");
  const long long total = (long long)n_maps * h * w;
  float* tmp = scratch;
  float* smooth = scratch + total;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)device_sm_count() * 32) blocks = (long long)device_sm_count() * 32;
  gauss_pass_kernel<0><<<(int)blocks, 256, 0, stream>>>(maps, tmp, n_maps, h, w, gw, radius);
  gauss_pass_kernel<1><<<(int)blocks, 256, 0, stream>>>(tmp, smooth, n_maps, h, w, gw, radius);
  collect_peaks_kernel<<<n_maps, kPeakThreads, 0, stream>>>(maps, smooth, h, w, offset, cap, peak_xy, peak_score,
                                                          peak_ij, counts, summary);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch(3);
  return 0;
}

extern "C" int dreamb200_gaussian_smooth(const float* maps, int n_maps, int h, int w, const double* gauss_w, int radius,
                                         float* scratch, float* out, void* stream_v) {
  cudaStream_t stream = (cudaStream_t)stream_v;
  DB_REQUIRE(maps && gauss_w && out, "gaussian_smooth: null pointer");
  DB_REQUIRE(n_maps > 0 && h > 0 && w > 0, "gaussian_smooth: empty input (n_maps=%d h=%d w=%d)", n_maps, h, w);
  DB_REQUIRE(radius >= 0 && radius <= kMaxRadius, "gaussian_smooth: radius %d out of range", radius);
  GaussW gw;
  for (int i = 0; i <= kMaxRadius; ++i) gw.w[i] = i <= radius ? gauss_w[i] : 0.0;
  if (fused_ok(h, w, radius) && !getenv("DREAMB200_PEAKS_UNFUSED") && !getenv("DREAMB200_PEAKS_BANDED")) {
    const size_t smem = fused_smem_bytes(h, w);
    if (fused_set_smem(smem)) return -2;
    peaks_fused_kernel<<<n_maps, kFusedThreads, smem, stream>>>(maps, n_maps, h, w, gw, 0.0, 1, nullptr, nullptr,
                                                                 nullptr, nullptr, nullptr, out);
  } else {
    DB_REQUIRE(scratch, "gaussian_smooth: this map size / radius needs n_maps*h*w floats of scratch");
    const long long total = (long long)n_maps * h * w;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)device_sm_count() * 32) blocks = (long long)device_sm_count() * 32;
    gauss_pass_kernel<0><<<(int)blocks, 256, 0, stream>>>(maps, scratch, n_maps, h, w, gw, radius);
    gauss_pass_kernel<1><<<(int)blocks, 256, 0, stream>>>(scratch, out, n_maps, h, w, gw, radius);
  }
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch(1);
  return 0;
}

extern "C" int dreamb200_peaks_plan(int n_maps, int h, int w, int radius, int cap, long long* scratch_floats,
                                    int* mode_out) {
  DB_REQUIRE(scratch_floats && mode_out && n_maps > 0 && h > 0 && w > 0 && cap >= 1, "peaks_plan: bad arguments");
  size_t band_smem = 0;
  int hb = 0;
  const int mode = peaks_mode(n_maps, h, w, radius, &hb, &band_smem);
  *mode_out = mode;
  if (mode == 0) *scratch_floats = 0;
  else if (mode == 1) *scratch_floats = (long long)((band_stage_bytes(n_maps, (h + hb - 1) / hb, cap) + 3) / 4);
  else *scratch_floats = 2LL * n_maps * h * w;
  return 0;
}

extern "C" int dreamb200_softargmax(const float* maps, const float* beta, float* out_xy, int B, int K, int H,
                                    int W, float* scratch, void* stream_v) {
  DB_REQUIRE(maps && beta && out_xy && scratch, "softargmax: null pointer");
  DB_REQUIRE(B > 0 && K > 0 && H > 0 && W > 0, "softargmax: empty input");
  softargmax_kernel<<<B * K, 256, 0, (cudaStream_t)stream_v>>>(maps, beta, out_xy, K, H, W, scratch);
  DB_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
