#include "gemm.cuh"

namespace ddb {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float shifted_softplus(float x) {
  // F.softplus (beta 1, threshold 20) - ln 2   (models/common.py:66-72)
  float sp = x > 20.f ? x : log1pf(expf(x));
  return sp - 0.69314718055994530942f;
}

// A tile element (k, m) lives at k*128 + (((m>>2) ^ ((k>>2)&31))<<2) + (m&3): the XOR keeps float4
// reads along m intact while spreading the transposing scalar stores of a row over 8 bank groups.
__device__ __forceinline__ int a_swz(int k, int m) { return k * GEMM_BM + ((((m >> 2) ^ ((k >> 2) & 31)) << 2) | (m & 3)); }

__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm128_kernel(const GemmArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                       // [128 k][128 m] swizzled
  float* Ws = smem + H * GEMM_BM;         // [2][GEMM_BK][GEMM_BN]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * GEMM_BM, col0 = blockIdx.y * GEMM_BN;
  const int M = a.M_dev ? min(__ldg(a.M_dev), a.M) : a.M;
  if (row0 >= M) return;

  auto load_w_stage = [&](int stage, int k0) {
    float* dst = Ws + stage * GEMM_BK * GEMM_BN;
#pragma unroll
    for (int i = 0; i < (GEMM_BK * GEMM_BN / 4) / GEMM_THREADS; ++i) {
      int f = tid + i * GEMM_THREADS;          // float4 index in the stage
      int kk = f >> 5, n4 = f & 31;
      cp_async16(dst + kk * GEMM_BN + n4 * 4, a.Wt + (size_t)(k0 + kk) * a.ldw + col0 + n4 * 4);
    }
    cp_async_commit();
  };
  load_w_stage(0, 0);
  load_w_stage(1, GEMM_BK);

  // ---- stage the A tile: one warp per row, full 512-byte rows, prologue on complete rows
  float4 gam = make_float4(1, 1, 1, 1), bet = make_float4(0, 0, 0, 0);
  const bool do_ln = a.ln_gamma != nullptr;
  if (do_ln) { gam = ldg4(a.ln_gamma + lane * 4); bet = ldg4(a.ln_beta + lane * 4); }
  for (int r = warp; r < GEMM_BM; r += GEMM_THREADS / 32) {
    int m = row0 + r;
    float4 z[1] = {make_float4(0, 0, 0, 0)};
    const int ar = m < M ? (a.a_rows ? a.a_rows[m] : m) : -1;
    if (ar >= 0) {
      z[0] = ld4(a.A + (size_t)ar * a.lda + lane * 4);
      if (a.A2) {
        int r2 = a.a2_rows[m];
        if (r2 >= 0) z[0] = add4(z[0], ld4(a.A2 + (size_t)r2 * a.lda2 + lane * 4));
      }
      if (do_ln) ln_relu_rows<1>(z, gam, bet, lane);
    }
    int k = lane * 4;
    As[a_swz(k + 0, r)] = z[0].x; As[a_swz(k + 1, r)] = z[0].y;
    As[a_swz(k + 2, r)] = z[0].z; As[a_swz(k + 3, r)] = z[0].w;
  }

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  constexpr int NKT = H / GEMM_BK;
  for (int kt = 0; kt < NKT; ++kt) {
    if (kt + 1 < NKT) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const float* ws = Ws + (kt & 1) * GEMM_BK * GEMM_BN;
#pragma unroll 8
    for (int kk = 0; kk < GEMM_BK; ++kk) {
      int k = kt * GEMM_BK + kk;
      int sw = (k >> 2) & 31;
      float4 a0 = ld4(As + k * GEMM_BM + ((ty ^ sw) << 2));
      float4 a1 = ld4(As + k * GEMM_BM + (((16 + ty) ^ sw) << 2));
      float4 b0 = ld4(ws + kk * GEMM_BN + tx * 4);
      float4 b1 = ld4(ws + kk * GEMM_BN + 64 + tx * 4);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
    if (kt + 2 < NKT) load_w_stage(kt & 1, (kt + 2) * GEMM_BK);
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    int m = row0 + r;
    if (m >= M) continue;
    int cr = a.c_rows ? a.c_rows[m] : m;
    if (cr < 0) continue;
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      int n = col0 + jb * 64 + tx * 4;
      float4 o = make_float4(acc[i][jb * 4 + 0], acc[i][jb * 4 + 1], acc[i][jb * 4 + 2], acc[i][jb * 4 + 3]);
      if (a.bias) o = add4(o, ldg4(a.bias + n));
      if (a.R) o = add4(o, ld4(a.R + (size_t)cr * a.ldr + n));
      if (a.act == 1) {
        o.x = shifted_softplus(o.x); o.y = shifted_softplus(o.y);
        o.z = shifted_softplus(o.z); o.w = shifted_softplus(o.w);
      }
      st4(a.C + (size_t)cr * a.ldc + n, o);
    }
  }
}

void launch_gemm128(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return;
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    cudaFuncSetAttribute(gemm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    attr_set.mark();
  }
  dim3 grid((a.M + GEMM_BM - 1) / GEMM_BM, a.N / GEMM_BN);
  gemm128_kernel<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(a);
}

}  // namespace ddb
