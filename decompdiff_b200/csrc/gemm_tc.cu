// Tensor-core projection GEMM for sm_100a:  C = act(prologue(A)[M,128] @ W[128,N] + bias) (+ R)
// tcgen05.mma (kind::tf32), A operand AND accumulators in TMEM, B (weights) streamed through shared memory.
//
// fp32 parity: plain TF32 inputs break the rtol 1e-4 contract, so every product is evaluated as
// the 3xTF32 split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (a_hi = rna_tf32(a), a_lo = rna_tf32(a - a_hi)), three
// MMAs into the same fp32 accumulator; the dropped a_lo*b_lo term is O(2^-22).
//
// Structure (one CTA = 16 worker warps + 1 MMA-issuing warp per 128-row tile x a slice of the N columns):
//   * worker thread (row r = 32q + lane, channel slice s) of warp w = 4s + q stages 32 channels of one row: gather / add /
//     LayerNorm+ReLU prologue (row statistics exchanged through smem between the 4 warps of a quadrant), hi/lo split,
//     tcgen05.st into TMEM (A_hi columns 0..127, A_lo columns 128..255; row -> lane, k -> column)
//   * B is pre-split and pre-swizzled on the host into 32-column chunks of 32 KB (hi | lo, K-major, 128B swizzle); a chunk
//     is ONE cp.async.bulk (TMA engine, mbarrier complete_tx) into a 4-stage ring
//   * the issuer warp waits for a chunk and a free accumulator, issues 48 tcgen05.mma (M128 N32 K8, A from TMEM) and
//     commits to an mbarrier - no CTA-wide barrier inside the chunk loop
//   * chunk c is drained by the 4 warps of slice c % 4: tcgen05.ld, release the accumulator, refill the ring stage with
//     chunk c + 4, smem transpose, bias / residual / activation, coalesced global stores
#include <cstring>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace ddb {

constexpr int TC_BM = 128;              // rows per CTA (UMMA M)
constexpr int TC_BN = 32;               // columns per chunk (UMMA N)
constexpr int TC_STAGES = 4;            // B ring depth == accumulator ring depth == number of epilogue warp groups
constexpr int TC_B_KB = TC_BN * 128;                    // 4 KB per K-block (32 tf32 along K) of a B chunk
constexpr int TC_B_PART = 4 * TC_B_KB;                  // 16 KB (hi or lo)
constexpr int TC_B_CHUNK = 2 * TC_B_PART;               // 32 KB per chunk (hi | lo)
constexpr int TC_WORKERS = 512;                         // 16 staging / epilogue warps
constexpr int TC_THREADS = TC_WORKERS + 32;             // + 1 MMA-issuing warp
constexpr int TC_EPI_LD = 36;                           // padded row of a per-warp 32x32 transpose tile (floats)
constexpr int TC_EPI_BYTES = 16 * 32 * TC_EPI_LD * 4;   // 72 KB
constexpr int TC_STAT_BYTES = 2 * 128 * 4 * 4;          // LayerNorm partial sums [row][slice] x {sum, centred squares}
constexpr int TC_SMEM = TC_STAGES * TC_B_CHUNK + TC_EPI_BYTES + TC_STAT_BYTES + 128 /*barriers*/;
constexpr int TC_COL_AHI = 0, TC_COL_ALO = 128, TC_COL_D = 256;   // TMEM column map (512 allocated)
constexpr int TC_BAR_A_READY = 5;
static_assert(TC_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ float ssp(float x) { return (x > 20.f ? x : log1pf(expf(x))) - 0.69314718055994530942f; }
__device__ __forceinline__ void tc_quad_barrier(int q) { asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(128) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1) gemm128_tc_kernel(const GemmArgs a, const float* __restrict__ Wtc,
                                                                   int chunks_per_cta) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;                                                // TC_STAGES x (hi | lo)
  float* sEpi = reinterpret_cast<float*>(sB + TC_STAGES * TC_B_CHUNK);   // per-warp 32 x 36 transpose tiles
  float* sStatA = sEpi + 16 * 32 * TC_EPI_LD;
  float* sStatB = sStatA + 128 * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStatB + 128 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * TC_BM;
  const int chunk0 = blockIdx.y * chunks_per_cta;
  const int n_chunks = min(chunks_per_cta, a.N / TC_BN - chunk0);
  auto bar_b = [&](int i) { return smem_u32(&bars[i]); };                     // B chunk landed in stage i
  auto bar_m = [&](int i) { return smem_u32(&bars[TC_STAGES + i]); };         // MMAs into accumulator i retired
  auto bar_f = [&](int i) { return smem_u32(&bars[2 * TC_STAGES + i]); };     // accumulator i drained (4 warp arrivals)

  if ((smem_u32(sB) & 1023u) != 0u) __trap();      // SWIZZLE_128B operands need a 1024-byte aligned base
  if (tid == 0) {
    for (int i = 0; i < 2 * TC_STAGES; ++i) mbar_init(smem_u32(&bars[i]), 1);
    for (int i = 0; i < TC_STAGES; ++i) mbar_init(bar_f(i), 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 16) {
    // ------------------------------------------------------------------------------------ MMA issuer warp
    if (lane == 0) {
      for (int c = 0; c < min(TC_STAGES, n_chunks); ++c) {      // fill the weight ring while the A tile is being staged
        mbar_expect_tx(bar_b(c), TC_B_CHUNK);
        bulk_g2s(smem_u32(sB + c * TC_B_CHUNK), Wtc + (size_t)(chunk0 + c) * (TC_B_CHUNK / 4), TC_B_CHUNK, bar_b(c));
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(TC_BAR_A_READY), "r"(TC_THREADS) : "memory");   // A tile is in TMEM
    if (lane == 0) {
      tc_fence_after();
      // instruction descriptor: D=F32, A=B=TF32, both K-major, N=32, M=128 (cute::UMMA::InstrDescriptor)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint64_t desc0 = umma_desc_sw128(smem_u32(sB));     // descriptors of other addresses differ in the low field only
      for (int c = 0; c < n_chunks; ++c) {
        const int st = c % TC_STAGES, use = c / TC_STAGES;
        mbar_wait(bar_b(st), use & 1);
        if (use > 0) mbar_wait(bar_f(st), (use - 1) & 1);       // accumulator st drained by the epilogue of chunk c - 4
        tc_fence_after();
        const uint64_t d_hi = desc0 + (uint64_t)((st * TC_B_CHUNK) >> 4), d_lo = d_hi + (uint64_t)(TC_B_PART >> 4);
        const uint32_t d = tmem_base + TC_COL_D + st * TC_BN;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {   // UMMA K = 8 tf32: 8 TMEM columns of A, 32 bytes inside the 128-byte swizzle atom of B
          const uint64_t bo = (uint64_t)(((kk >> 2) * TC_B_KB + (kk & 3) * 32) >> 4);
          umma_tf32_ts(d, tmem_base + TC_COL_AHI + kk * 8, d_hi + bo, idesc, kk ? 1u : 0u);
          umma_tf32_ts(d, tmem_base + TC_COL_ALO + kk * 8, d_hi + bo, idesc, 1u);
          umma_tf32_ts(d, tmem_base + TC_COL_AHI + kk * 8, d_lo + bo, idesc, 1u);
        }
        umma_commit(bar_m(st));             // implies tcgen05.fence::before_thread_sync
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------ worker warps
    const int q = warp & 3, s = warp >> 2;                  // TMEM lane quadrant, channel slice
    {
      // ---- stage A into TMEM: this thread owns channels [32s, 32s+32) of row 32q + lane
      const int r = q * 32 + lane, m = row0 + r;
      float z[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = 0.f;
      if (m < a.M) {
        const float* src = a.A + (size_t)(a.a_rows ? a.a_rows[m] : m) * a.lda + s * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) { float4 v = ld4(src + i * 4); z[4 * i] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w; }
        if (a.A2) {
          int r2 = a.a2_rows[m];
          if (r2 >= 0) {
            const float* s2 = a.A2 + (size_t)r2 * a.lda2 + s * 32;
#pragma unroll
            for (int i = 0; i < 8; ++i) { float4 v = ld4(s2 + i * 4); z[4 * i] += v.x; z[4 * i + 1] += v.y; z[4 * i + 2] += v.z; z[4 * i + 3] += v.w; }
          }
        }
      }
      if (a.ln_gamma != nullptr) {       // LayerNorm + ReLU over the full row: two-pass statistics via smem
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) p += z[i];
        sStatA[r * 4 + s] = p;
        tc_quad_barrier(q);
        float4 t = ld4(sStatA + r * 4);
        const float mu = ((t.x + t.y) + (t.z + t.w)) * (1.0f / H);
        p = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { z[i] -= mu; p = fmaf(z[i], z[i], p); }
        sStatB[r * 4 + s] = p;
        tc_quad_barrier(q);
        t = ld4(sStatB + r * 4);
        const float rstd = 1.0f / sqrtf(((t.x + t.y) + (t.z + t.w)) * (1.0f / H) + LN_EPS);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 g = ldg4(a.ln_gamma + s * 32 + i4 * 4), b = ldg4(a.ln_beta + s * 32 + i4 * 4);
          z[i4 * 4 + 0] = fmaxf(fmaf(z[i4 * 4 + 0] * rstd, g.x, b.x), 0.f);
          z[i4 * 4 + 1] = fmaxf(fmaf(z[i4 * 4 + 1] * rstd, g.y, b.y), 0.f);
          z[i4 * 4 + 2] = fmaxf(fmaf(z[i4 * 4 + 2] * rstd, g.z, b.z), 0.f);
          z[i4 * 4 + 3] = fmaxf(fmaf(z[i4 * 4 + 3] * rstd, g.w, b.w), 0.f);
        }
      }
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { float h = tf32_rna(z[i]); hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(tf32_rna(z[i] - h)); }
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      tmem_st32(lane_addr + TC_COL_AHI + s * 32, hi);
      tmem_st32(lane_addr + TC_COL_ALO + s * 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      asm volatile("bar.arrive %0, %1;" ::"r"(TC_BAR_A_READY), "r"(TC_THREADS) : "memory");
    }
    // ---- epilogue: slice group s drains chunks s, s+4, s+8, ...
    float* tile = sEpi + warp * 32 * TC_EPI_LD;
    for (int c = s; c < n_chunks; c += TC_STAGES) {
      const int st = s, use = c / TC_STAGES;                // c % TC_STAGES == s
      mbar_wait(bar_m(st), use & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + TC_COL_D + st * TC_BN, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_f(st));                             // accumulator st may be overwritten
        if (q == 0 && c + TC_STAGES < n_chunks) {           // MMAs of chunk c retired -> ring stage st is free: refill it
          mbar_expect_tx(bar_b(st), TC_B_CHUNK);
          bulk_g2s(smem_u32(sB + st * TC_B_CHUNK), Wtc + (size_t)(chunk0 + c + TC_STAGES) * (TC_B_CHUNK / 4), TC_B_CHUNK, bar_b(st));
        }
      }
      // thread = row -> transpose through smem so that 8 lanes write one 128-byte segment of a row
#pragma unroll
      for (int j = 0; j < TC_BN; j += 4)
        st4(tile + lane * TC_EPI_LD + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
      __syncwarp();
      const int n0 = (chunk0 + c) * TC_BN + (lane & 7) * 4;
      const float4 bias4 = a.bias ? ldg4(a.bias + n0) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rl = it * 4 + (lane >> 3);
        const int m = row0 + q * 32 + rl;
        if (m < a.M) {
          const int cr = a.c_rows ? a.c_rows[m] : m;
          float4 o = add4(ld4(tile + rl * TC_EPI_LD + (lane & 7) * 4), bias4);
          if (a.R) o = add4(o, ld4(a.R + (size_t)cr * a.ldr + n0));
          if (a.act == 1) { o.x = ssp(o.x); o.y = ssp(o.y); o.z = ssp(o.z); o.w = ssp(o.w); }
          st4(a.C + (size_t)cr * a.ldc + n0, o);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

void launch_gemm128_tc(const GemmArgs& a, const float* Wtc, int num_sms, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm128_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    attr_set = true;
  }
  const int row_tiles = (a.M + TC_BM - 1) / TC_BM, chunks = a.N / TC_BN;
  // Split the N chunks over `nsplit` CTAs per row tile.  Cost model in units of one chunk of MMA work: every CTA pays ~3
  // units to stage its A tile, then chunks/nsplit units; CTAs run in waves of num_sms (1 CTA per SM).
  int best = 1; double best_cost = 1e30;
  for (int ns = 1; ns <= chunks; ++ns) {
    if (chunks % ns) continue;
    long ctas = (long)row_tiles * ns;
    double waves = (double)((ctas + num_sms - 1) / num_sms);
    double cost = waves * (3.0 + (double)chunks / ns);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ns; }
  }
  const int per = chunks / best;
  dim3 grid(row_tiles, (chunks + per - 1) / per);
  gemm128_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(a, Wtc, per);
}

// host-side packing of a K-major weight Wt[128][N] into the chunked, hi/lo-split, 128B-swizzled image the kernel copies
void pack_gemm_tc(const float* Wt, int N, float* out /* N*128*2 floats */) {
  const int chunks = N / TC_BN;
  for (int c = 0; c < chunks; ++c) {
    float* hi = out + (size_t)c * (TC_B_CHUNK / 4);
    float* lo = hi + TC_B_PART / 4;
    for (int nl = 0; nl < TC_BN; ++nl) {
      for (int k = 0; k < H; ++k) {
        float w = Wt[(size_t)k * N + c * TC_BN + nl];
        float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
        int off = sw128_offset_bytes(nl, k, TC_BN) / 4;
        hi[off] = h;
        lo[off] = l;
      }
    }
  }
}

}  // namespace ddb
