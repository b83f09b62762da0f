// Tensor-core projection GEMM for sm_100a:  C = act(prologue(A)[M,128] @ W[128,N] + bias) (+ R)
// tcgen05.mma (kind::tf32) with the accumulator in TMEM, operands in 128B-swizzled shared memory.
//
// fp32 parity: plain TF32 inputs break the rtol 1e-4 contract, so every product is evaluated as
// the 3xTF32 split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (a_hi = rna_tf32(a), a_lo = rna_tf32(a - a_hi)), three
// MMAs into the same fp32 accumulator; the dropped a_lo*b_lo term is O(2^-22).
//
// Structure (one CTA per 128-row tile x a slice of the N columns):
//   * all threads stage the A tile once: full rows (gather / add / LayerNorm+ReLU prologue), split hi/lo, written
//     K-major into 4 K-blocks of [128 rows x 128 B] with the 128B swizzle the UMMA descriptor expects
//   * B (weights) is pre-split and pre-swizzled on the host into 32-column chunks of 32 KB (hi | lo); a chunk is ONE
//     cp.async.bulk (TMA engine, mbarrier complete_tx), double buffered
//   * one elected thread issues 48 tcgen05.mma (M128 N32 K8) per chunk and commits to an mbarrier; two 32-column TMEM
//     accumulators ping-pong so the epilogue of chunk c-1 (tcgen05.ld -> bias/residual/activation -> global) and the
//     load of chunk c+1 overlap the MMAs of chunk c
#include <cstring>

#include "gemm.cuh"

namespace ddb {

constexpr int TC_BM = 128;              // rows per CTA (UMMA M)
constexpr int TC_BN = 32;               // columns per chunk (UMMA N)
constexpr int TC_KB_BYTES = 128;        // one swizzle atom row: 32 tf32 along K
constexpr int TC_NKB = 4;               // K = 128 = 4 K-blocks of 32
constexpr int TC_A_KB = TC_BM * TC_KB_BYTES;            // 16 KB per K-block of A
constexpr int TC_A_PART = TC_NKB * TC_A_KB;             // 64 KB (hi or lo)
constexpr int TC_B_KB = TC_BN * TC_KB_BYTES;            // 4 KB per K-block of a B chunk
constexpr int TC_B_PART = TC_NKB * TC_B_KB;             // 16 KB (hi or lo)
constexpr int TC_B_CHUNK = 2 * TC_B_PART;               // 32 KB per chunk (hi | lo)
constexpr int TC_THREADS = 256;          // 8 warps stage A; warps 0-3 / 4-7 drain the even / odd accumulator
constexpr int TC_EPI_LD = 36;            // padded row of the per-warp 32x32 transpose tile (floats)
constexpr int TC_EPI_BYTES = 4 * 32 * TC_EPI_LD * 4;   // only one warp group drains at a time
constexpr int TC_SMEM = 2 * TC_A_PART + 2 * TC_B_CHUNK + TC_EPI_BYTES + 1024 /*alignment slack*/ + 64 /*barriers*/;
constexpr int TC_TMEM_COLS = 64;        // two 32-column fp32 accumulators

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
static_assert(TC_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1ull << 28)) __trap();     // turn a protocol bug into an error instead of a hang
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO<<16 | SBO<<32 |
// version 1 <<46 | layout SWIZZLE_128B (2) << 61.  SBO = 1024 B between 8-row groups; LBO is 1 for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float ssp(float x) { return (x > 20.f ? x : log1pf(expf(x))) - 0.69314718055994530942f; }

__global__ void __launch_bounds__(TC_THREADS, 1) gemm128_tc_kernel(const GemmArgs a, const float* __restrict__ Wtc,
                                                                   int chunks_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA_hi = smem;
  uint8_t* sA_lo = smem + TC_A_PART;
  uint8_t* sB = smem + 2 * TC_A_PART;                       // 2 buffers x (hi | lo)
  float* sEpi = reinterpret_cast<float*>(sB + 2 * TC_B_CHUNK);      // per-warp 32 x 36 transpose tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * TC_B_CHUNK + TC_EPI_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t bar_b0 = smem_u32(&bars[0]), bar_b1 = smem_u32(&bars[1]);        // B chunk landed
  const uint32_t bar_m0 = smem_u32(&bars[2]), bar_m1 = smem_u32(&bars[3]);        // MMAs of a chunk retired
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * TC_BM;
  const int chunk0 = blockIdx.y * chunks_per_cta;
  const int n_chunks = min(chunks_per_cta, a.N / TC_BN - chunk0);

  if (tid == 0) {
    mbar_init(bar_b0, 1); mbar_init(bar_b1, 1); mbar_init(bar_m0, 1); mbar_init(bar_m1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  // kick off the first two weight chunks while the A tile is being staged
  if (tid == 0) {
    for (int c = 0; c < min(2, n_chunks); ++c) {
      uint32_t bar = c ? bar_b1 : bar_b0;
      mbar_expect_tx(bar, TC_B_CHUNK);
      bulk_g2s(smem_u32(sB + c * TC_B_CHUNK), Wtc + (size_t)(chunk0 + c) * (TC_B_CHUNK / 4), TC_B_CHUNK, bar);
    }
  }

  // ---- stage A: one warp per row; lane l owns channels 4l..4l+3 = K-block l>>3, 16-byte chunk l&7 of that block
  {
    float4 gam = make_float4(1, 1, 1, 1), bet = make_float4(0, 0, 0, 0);
    const bool do_ln = a.ln_gamma != nullptr;
    if (do_ln) { gam = ldg4(a.ln_gamma + lane * 4); bet = ldg4(a.ln_beta + lane * 4); }
    const int kb = lane >> 3, ch = lane & 7;
    constexpr int ROWS_PER_WARP = TC_BM / (TC_THREADS / 32);
    for (int rb = 0; rb < ROWS_PER_WARP; rb += 4) {
      float4 z[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {         // all loads first: 4 rows in flight per warp
        const int m = row0 + warp * ROWS_PER_WARP + rb + i;
        z[i] = make_float4(0, 0, 0, 0);
        if (m < a.M) {
          int ar = a.a_rows ? a.a_rows[m] : m;
          z[i] = ld4(a.A + (size_t)ar * a.lda + lane * 4);
          if (a.A2) {
            int r2 = a.a2_rows[m];
            if (r2 >= 0) z[i] = add4(z[i], ld4(a.A2 + (size_t)r2 * a.lda2 + lane * 4));
          }
        }
      }
      if (do_ln) ln_relu_rows<4>(z, gam, bet, lane);    // rows past M are zeros: LN of zeros is finite, never stored
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = warp * ROWS_PER_WARP + rb + i;
        float4 hi = make_float4(tf32_rna(z[i].x), tf32_rna(z[i].y), tf32_rna(z[i].z), tf32_rna(z[i].w));
        float4 lo = make_float4(tf32_rna(z[i].x - hi.x), tf32_rna(z[i].y - hi.y), tf32_rna(z[i].z - hi.z), tf32_rna(z[i].w - hi.w));
        int off = kb * TC_A_KB + r * TC_KB_BYTES + ((ch ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(sA_hi + off) = hi;
        *reinterpret_cast<float4*>(sA_lo + off) = lo;
      }
    }
  }
  // generic-proxy writes -> visible to the async proxy (tensor core reads smem through it)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptor: D=F32, A=B=TF32, both K-major, N=32, M=128 (cute::UMMA::InstrDescriptor)
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  const uint32_t a_hi = smem_u32(sA_hi), a_lo = smem_u32(sA_lo);

  auto epilogue = [&](int c) {          // chunk c (32 columns) is drained by warps 0-3 (even c) or 4-7 (odd c)
    if ((warp >> 2) != (c & 1)) return;
    const int q = warp & 3;             // TMEM lane quadrant this warp may read: lanes 32q .. 32q+31
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((c & 1) * TC_BN);
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // thread = row -> transpose through a padded smem tile so that 8 lanes write one 128-byte row segment
    float* tile = sEpi + q * 32 * TC_EPI_LD;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < TC_BN; j += 4)
      st4(tile + lane * TC_EPI_LD + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
    __syncwarp();
    const int n0 = (chunk0 + c) * TC_BN + (lane & 7) * 4;
    float4 bias4 = a.bias ? ldg4(a.bias + n0) : make_float4(0, 0, 0, 0);
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
  };

  for (int c = 0; c < n_chunks; ++c) {
    const int buf = c & 1;
    if (tid == 0) {
      mbar_wait(buf ? bar_b1 : bar_b0, (c >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t b_hi = smem_u32(sB + buf * TC_B_CHUNK), b_lo = b_hi + TC_B_PART;
      const uint32_t d = tmem_base + (uint32_t)(buf * TC_BN);
#pragma unroll
      for (int kb = 0; kb < TC_NKB; ++kb) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {       // UMMA K = 8 tf32 = 32 bytes inside the 128-byte swizzle atom
          const uint32_t ao = kb * TC_A_KB + ks * 32, bo = kb * TC_B_KB + ks * 32;
          umma_tf32(d, umma_desc(a_hi + ao), umma_desc(b_hi + bo), idesc, (kb | ks) ? 1u : 0u);
          umma_tf32(d, umma_desc(a_lo + ao), umma_desc(b_hi + bo), idesc, 1u);
          umma_tf32(d, umma_desc(a_hi + ao), umma_desc(b_lo + bo), idesc, 1u);
        }
      }
      umma_commit(buf ? bar_m1 : bar_m0);      // implies tcgen05.fence::before_thread_sync
    }
    if (c > 0) {
      const int pb = (c - 1) & 1;
      mbar_wait(pb ? bar_m1 : bar_m0, ((c - 1) >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue(c - 1);
      // accumulator + weight buffer pb are free again: fetch chunk c+1 into it
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (tid == 0 && c + 1 < n_chunks) {
        uint32_t bar = pb ? bar_b1 : bar_b0;
        mbar_expect_tx(bar, TC_B_CHUNK);
        bulk_g2s(smem_u32(sB + pb * TC_B_CHUNK), Wtc + (size_t)(chunk0 + c + 1) * (TC_B_CHUNK / 4), TC_B_CHUNK, bar);
      }
    }
  }
  if (n_chunks > 0) {
    const int c = n_chunks - 1;
    mbar_wait((c & 1) ? bar_m1 : bar_m0, (c >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(c);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
}

void launch_gemm128_tc(const GemmArgs& a, const float* Wtc, int num_sms, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm128_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    attr_set = true;
  }
  const int row_tiles = (a.M + TC_BM - 1) / TC_BM, chunks = a.N / TC_BN;
  // Split the N chunks over `nsplit` CTAs per row tile.  Cost model in units of one chunk of MMA work: every CTA pays ~4
  // units to stage its A tile, then chunks/nsplit units; CTAs run in waves of num_sms (1 CTA per SM).
  int best = 1; double best_cost = 1e30;
  for (int ns = 1; ns <= chunks; ++ns) {
    if (chunks % ns) continue;
    long ctas = (long)row_tiles * ns;
    double waves = (double)((ctas + num_sms - 1) / num_sms);
    double cost = waves * (4.0 + (double)chunks / ns);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = ns; }
  }
  const int per = chunks / best;
  dim3 grid(row_tiles, (chunks + per - 1) / per);
  gemm128_tc_kernel<<<grid, TC_THREADS, TC_SMEM, stream>>>(a, Wtc, per);
}

// host-side packing of a K-major weight Wt[128][N] into the chunked, hi/lo-split, 128B-swizzled image the kernel copies
void pack_gemm_tc(const float* Wt, int N, float* out /* N*128*2 floats */) {
  auto rna = [](float x) {
    uint32_t u; memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u;
    u &= 0xffffe000u;
    float r; memcpy(&r, &u, 4);
    return r;
  };
  const int chunks = N / TC_BN;
  for (int c = 0; c < chunks; ++c) {
    float* hi = out + (size_t)c * (TC_B_CHUNK / 4);
    float* lo = hi + TC_B_PART / 4;
    for (int nl = 0; nl < TC_BN; ++nl) {
      for (int k = 0; k < H; ++k) {
        float w = Wt[(size_t)k * N + c * TC_BN + nl];
        float h = rna(w), l = rna(w - h);
        int kb = k >> 5, kk = k & 31;
        int off_bytes = kb * TC_B_KB + nl * TC_KB_BYTES + ((((kk >> 2) ^ (nl & 7))) << 4) + (kk & 3) * 4;
        hi[off_bytes / 4] = h;
        lo[off_bytes / 4] = l;
      }
    }
  }
}

}  // namespace ddb
