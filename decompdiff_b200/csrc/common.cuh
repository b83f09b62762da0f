// Shared device helpers for the decompdiff_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddb {

constexpr int H = 128;        // hidden_dim (configs/training.yml:44)
constexpr int NH = 16;        // n_heads
constexpr int DH = 8;         // H / NH
constexpr int KNN = 32;       // max neighbours per node (one warp lane per neighbour slot)
constexpr int NG = 20;        // Gaussian smearing width (models/common.py:18)
constexpr int NANG = 13;      // AngularEncoding width (models/common.py:43)
constexpr float LN_EPS = 1e-5f;
constexpr unsigned FULL = 0xffffffffu;

// Fixed Gaussian offsets of models/common.py:18 (fix_offset=True ignores start/stop).
__constant__ float c_gauss_offset[NG] = {0.f, 1.f, 1.25f, 1.5f, 1.75f, 2.f, 2.25f, 2.5f, 2.75f, 3.f,
                                         3.5f, 4.f, 4.5f, 5.f, 5.5f, 6.f, 7.f, 8.f, 9.f, 10.f};

__device__ __forceinline__ float gauss_feat(float d, int g) {
  float t = d - c_gauss_offset[g];
  return expf(-0.5f * t * t);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, m));
  return v;
}

// Sum-all-reduce of 4 independent values across the warp with 9 shuffles
// (reduce-scatter 4->1, 3 butterfly steps, all-gather 1->4).
__device__ __forceinline__ void warp_allreduce4(float (&v)[4], int lane) {
  const bool u16 = lane & 16, u8 = lane & 8;
  float k0 = u16 ? v[2] : v[0], k1 = u16 ? v[3] : v[1];
  float s0 = u16 ? v[0] : v[2], s1 = u16 ? v[1] : v[3];
  k0 += __shfl_xor_sync(FULL, s0, 16);
  k1 += __shfl_xor_sync(FULL, s1, 16);
  float k = u8 ? k1 : k0, s = u8 ? k0 : k1;
  k += __shfl_xor_sync(FULL, s, 8);
  k += __shfl_xor_sync(FULL, k, 4);
  k += __shfl_xor_sync(FULL, k, 2);
  k += __shfl_xor_sync(FULL, k, 1);
  // lane now holds total of index (u16*2 + u8); gather back
  float o = __shfl_xor_sync(FULL, k, 8);
  float p0 = u8 ? o : k, p1 = u8 ? k : o;     // pair (u16*2+0, u16*2+1)
  float q0 = __shfl_xor_sync(FULL, p0, 16), q1 = __shfl_xor_sync(FULL, p1, 16);
  v[0] = u16 ? q0 : p0; v[1] = u16 ? q1 : p1;
  v[2] = u16 ? p0 : q0; v[3] = u16 ? p1 : q1;
}

// Butterfly reduce-scatter: every lane holds N partial values; afterwards lane l holds the
// warp-wide sums of indices [l*N/32, (l+1)*N/32) in v[0 .. N/32).  N/2 + N/4 + ... shuffles.
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N], int lane) {
  static_assert(N >= 32 && (N & (N - 1)) == 0, "N must be a power of two >= 32");
  int n = N;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    const bool upper = lane & m;
    const int half = n >> 1;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      if (i < half) {
        float keep = upper ? v[i + half] : v[i];
        float send = upper ? v[i] : v[i + half];
        v[i] = keep + __shfl_xor_sync(FULL, send, m);
      }
    }
    n = half;
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 fma4(float s, float4 w, float4 acc) {
  return make_float4(fmaf(s, w.x, acc.x), fmaf(s, w.y, acc.y), fmaf(s, w.z, acc.z), fmaf(s, w.w, acc.w));
}
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// LayerNorm(128, eps 1e-5, affine) + ReLU on BLK rows at once; each lane owns 4 channels
// (lane*4 .. lane*4+3) of every row.  Two-pass statistics (mean, then centred variance).
template <int BLK>
__device__ __forceinline__ void ln_relu_rows(float4 (&z)[BLK], float4 gamma, float4 beta, int lane) {
  static_assert(BLK == 4 || BLK == 1, "BLK");
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < BLK; ++r) s[r] = (z[r].x + z[r].y) + (z[r].z + z[r].w);
  if (BLK == 4) warp_allreduce4(s, lane); else s[0] = warp_sum(s[0]);
  float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < BLK; ++r) {
    float mu = s[r] * (1.0f / H);
    z[r].x -= mu; z[r].y -= mu; z[r].z -= mu; z[r].w -= mu;
    q[r] = (z[r].x * z[r].x + z[r].y * z[r].y) + (z[r].z * z[r].z + z[r].w * z[r].w);
  }
  if (BLK == 4) warp_allreduce4(q, lane); else q[0] = warp_sum(q[0]);
#pragma unroll
  for (int r = 0; r < BLK; ++r) {
    float rstd = 1.0f / sqrtf(q[r] * (1.0f / H) + LN_EPS);
    z[r].x = fmaxf(fmaf(z[r].x * rstd, gamma.x, beta.x), 0.f);
    z[r].y = fmaxf(fmaf(z[r].y * rstd, gamma.y, beta.y), 0.f);
    z[r].z = fmaxf(fmaf(z[r].z * rstd, gamma.z, beta.z), 0.f);
    z[r].w = fmaxf(fmaf(z[r].w * rstd, gamma.w, beta.w), 0.f);
  }
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: one flag per (call site, device).  A thread that
// races the first caller at worst sets the attribute a second time; it never launches before the attribute is set.
struct DeviceOnce {
  volatile bool flag[64] = {};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool done() const { return flag[dev()]; }
  void mark() { flag[dev()] = true; }
};

// cooperative copy of `n_floats` (multiple of 4) from global to shared by the whole CTA
__device__ __forceinline__ void cta_copy_f4(float* dst, const float* __restrict__ src, int n_floats) {
  for (int i = threadIdx.x * 4; i < n_floats; i += blockDim.x * 4) st4(dst + i, ldg4(src + i));
}

}  // namespace ddb
