// Tensor-core attention kernels (sm_100a): the same math as attn_knn.cu / attn_bond.cu (fp32 cross-check kernels), restructured
// so that the second Linear of the key / value MLPs - and the feature term of the first Linear - run on tcgen05 as true GEMMs
// with shared weights.  Pieces shared by attn_tc_knn.cu (kNN edges), attn_tc_trip.cu (triplets) and attn_tc_bond.cu (bond edges).
//
//   rows   : one attention candidate each (a kNN edge of a destination node / a triplet k->j->i of a bond edge j->i / a bond edge
//            entering a ligand atom); 32 rows form one softmax group = one TMEM lane quadrant, 4 groups form a 128-row tile
//   threads: 640 = 16 worker warps + one warpgroup whose first warp issues every tcgen05.mma (the issuing thread is paced by the
//            tensor pipe, so it carries no row work; setmaxnreg hands the group's registers to the workers: 112 / 32).
//            Worker warp w = 4*s + q owns rows 32q..32q+31 (thread = row) and hidden channels 32s..32s+31
//   hidden : a = ReLU(LN(first-Linear pieces))  computed thread-per-row with packed f32x2 math: no cross-lane reductions, row
//            statistics are exchanged once between the 4 slice-warps of a quadrant through shared memory
//   GEMM   : a (hi/lo TF32 split) is written to TMEM with tcgen05.st and used as the A operand of 48 tcgen05.mma
//            (M128 N128 K8, 3xTF32) against W2 (hi/lo, K-major, 128B swizzle) resident in shared memory for the whole
//            persistent kernel; D (128 x 128 fp32) lives in TMEM
//   k pass : logits[row, head] = <q_group[head], D[row, head]>  (thread-local dot products), fused per-group softmax (one REDUX
//            per head for the max, a transposed all-reduce for the sums), times e_w -> wbuf
//   v pass : out[group, c] = sum_rows w[row, head(c)] D[row, c] + b2[c] sum_rows w[row, head(c)]  via one 32-value
//            butterfly reduce-scatter across the 32 rows
//   The main MMA of tile t overlaps the first Linear / LayerNorm of tile t+1; D is drained into registers before the next A is
//   written, so the tensor core restarts while the softmax / weighted sums of tile t are finished from registers.
#pragma once
#include "kernels.cuh"
#include "tc_common.cuh"

namespace ddb {

// Optional in-kernel timeline (compile with -DDDB_TIMELINE): CTA 0, lane 0 of warps 0 and 13 add the SM cycles spent in each
// phase of the tile loop to g_timeline[pass][warp slot][phase]; read with ddb_debug_timeline_{trip,knn}() (profiles/timeline.py).
#ifdef DDB_TIMELINE
static __device__ unsigned long long g_timeline[2][2][16];      // one copy per translation unit: [pass][warp slot][phase]
#ifndef TL_WA
#define TL_WA 0
#define TL_WB 13
#endif
#define TL_DECL unsigned long long tl_t = clock64(), tl_acc[12] = {0}; const bool tl_on = blockIdx.x == 0 && lane == 0 && (warp == TL_WA || warp == TL_WB);
#define TL_MARK(i) do { if (tl_on) { unsigned long long n_ = clock64(); tl_acc[i] += n_ - tl_t; tl_t = n_; } } while (0)
#define TL_FLUSH(kid) do { if (tl_on) { for (int i_ = 0; i_ < 12; ++i_) g_timeline[kid][warp == TL_WA ? 0 : 1][i_] = tl_acc[i_]; g_timeline[kid][warp == TL_WA ? 0 : 1][12] = it; } } while (0)
#else
#define TL_DECL
#define TL_MARK(i)
#define TL_FLUSH(kid)
#endif

constexpr int ATC_THREADS = 512;
constexpr int ATC_W2_BYTES = 2 * 128 * 128 * 4;                       // hi | lo image of W2
constexpr int ANG_LD = 20;              // padded row of the per-tile angular features (bank-conflict-free float4 reads)
constexpr int ATC_COL_AHI = 0, ATC_COL_ALO = 128, ATC_COL_D = 256;    // TMEM column map (512 allocated)

__device__ __forceinline__ void quad_barrier(int q) { asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(128) : "memory"); }

// hi = a truncated to TF32 (exact: hi + lo == a), lo = the remainder; the tensor core reads only the 19 MSBs of lo, so the
// split carries a relative error <= 2^-21 per element - 2 instructions instead of 4 for the round-to-nearest variant
__device__ __forceinline__ void tf32_split(float a, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(a) & 0xffffe000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
}

// D (+)= A[tmem hi/lo] * W2[smem hi/lo]  : 16 k-steps x 3 MMAs, N = 128.  The k loop is NOT unrolled: the issuing thread is
// paced by the tensor pipe (64 cycles per MMA) anyway, and 48 hoisted descriptors would only spill.
__device__ __forceinline__ void atc_issue_mma(uint32_t tmem_base, uint32_t w2_smem, uint32_t bar) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t d = tmem_base + ATC_COL_D;
  const uint64_t b_hi0 = umma_desc_sw128(w2_smem), b_lo0 = umma_desc_sw128(w2_smem + ATC_W2_BYTES / 2);
#pragma unroll 1
  for (int kk = 0; kk < 16; ++kk) {
    const uint64_t bo = (uint64_t)(((kk >> 2) * (128 * 128) + (kk & 3) * 32) >> 4);      // start-address field is in 16-byte units
    umma_tf32_ts(d, tmem_base + ATC_COL_AHI + kk * 8, b_hi0 + bo, idesc, kk ? 1u : 0u);
#ifndef DDB_EXP_MMA1      // timing experiment only (wrong results): one pass instead of the 3xTF32 split
    umma_tf32_ts(d, tmem_base + ATC_COL_ALO + kk * 8, b_hi0 + bo, idesc, 1u);
    umma_tf32_ts(d, tmem_base + ATC_COL_AHI + kk * 8, b_lo0 + bo, idesc, 1u);
#endif
  }
  umma_commit(bar);
}

// same with N output columns (W2 image: hi | lo, each 4 K-blocks of [N rows][128 B]); N = 16 serves the position value MLPs
template <int N>
__device__ __forceinline__ void atc_issue_mma_n(uint32_t tmem_base, uint32_t w2_smem, uint32_t bar) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t d = tmem_base + ATC_COL_D;
  const uint64_t b_hi0 = umma_desc_sw128(w2_smem), b_lo0 = umma_desc_sw128(w2_smem + 4 * N * 128);
#pragma unroll 1
  for (int kk = 0; kk < 16; ++kk) {
    const uint64_t bo = (uint64_t)(((kk >> 2) * (N * 128) + (kk & 3) * 32) >> 4);
    umma_tf32_ts(d, tmem_base + ATC_COL_AHI + kk * 8, b_hi0 + bo, idesc, kk ? 1u : 0u);
    umma_tf32_ts(d, tmem_base + ATC_COL_ALO + kk * 8, b_hi0 + bo, idesc, 1u);
    umma_tf32_ts(d, tmem_base + ATC_COL_AHI + kk * 8, b_lo0 + bo, idesc, 1u);
  }
  umma_commit(bar);
}
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
}

// LayerNorm(128) + ReLU on a row whose 128 channels are spread over the 4 slice-warps of quadrant q
__device__ __forceinline__ void atc_ln_relu(float (&z)[32], float* statA, float* statB, int r, int s, int q,
                                            const float* sGamma, const float* sBeta) {
  float p = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) p += z[i];
  statA[r * 4 + s] = p;
  quad_barrier(q);
  float4 t = ld4(statA + r * 4);
  const float mu = ((t.x + t.y) + (t.z + t.w)) * (1.0f / H);
  p = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) { z[i] -= mu; p = fmaf(z[i], z[i], p); }
  statB[r * 4 + s] = p;
  quad_barrier(q);
  t = ld4(statB + r * 4);
  const float rstd = 1.0f / sqrtf(((t.x + t.y) + (t.z + t.w)) * (1.0f / H) + LN_EPS);
#pragma unroll
  for (int i4 = 0; i4 < 8; ++i4) {
    const float4 g = ld4(sGamma + s * 32 + i4 * 4), b = ld4(sBeta + s * 32 + i4 * 4);
    z[i4 * 4 + 0] = fmaxf(fmaf(z[i4 * 4 + 0] * rstd, g.x, b.x), 0.f);
    z[i4 * 4 + 1] = fmaxf(fmaf(z[i4 * 4 + 1] * rstd, g.y, b.y), 0.f);
    z[i4 * 4 + 2] = fmaxf(fmaf(z[i4 * 4 + 2] * rstd, g.z, b.z), 0.f);
    z[i4 * 4 + 3] = fmaxf(fmaf(z[i4 * 4 + 3] * rstd, g.w, b.w), 0.f);
  }
}

// hidden activations -> TMEM (A_hi / A_lo), then the CTA-wide hand-over to the MMA-issuing thread
__device__ __forceinline__ void atc_store_and_mma(const float (&z)[32], bool keep, uint32_t tmem_base, int q, int s,
                                                  uint32_t w2_smem, uint32_t bar) {
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) tf32_split(keep ? z[i] : 0.f, hi[i], lo[i]);
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  tmem_st32(lane_addr + ATC_COL_AHI + s * 32, hi);
  tmem_st32(lane_addr + ATC_COL_ALO + s * 32, lo);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) { tc_fence_after(); atc_issue_mma(tmem_base, w2_smem, bar); }
}

// k-pass epilogue: 4 head logits of this thread's row from its 32 D columns, fused softmax over the warp's 32 rows
__device__ __forceinline__ float4 atc_logits_softmax(uint32_t tmem_base, int q, int s, const float* qrow, bool rowok) {
  uint32_t v[32];
  tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  float lg[4];
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    const float4 q0 = ld4(qrow + s * 32 + hh * 8), q1 = ld4(qrow + s * 32 + hh * 8 + 4);
    float a = q0.x * __uint_as_float(v[hh * 8]);
    a = fmaf(q0.y, __uint_as_float(v[hh * 8 + 1]), a); a = fmaf(q0.z, __uint_as_float(v[hh * 8 + 2]), a);
    a = fmaf(q0.w, __uint_as_float(v[hh * 8 + 3]), a); a = fmaf(q1.x, __uint_as_float(v[hh * 8 + 4]), a);
    a = fmaf(q1.y, __uint_as_float(v[hh * 8 + 5]), a); a = fmaf(q1.z, __uint_as_float(v[hh * 8 + 6]), a);
    a = fmaf(q1.w, __uint_as_float(v[hh * 8 + 7]), a);
    lg[hh] = rowok ? a : -INFINITY;
  }
  float w[4];
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    const float m = warp_max(lg[hh]);
    const float ex = rowok ? expf(lg[hh] - m) : 0.f;
    const float ssum = warp_sum(ex);
    w[hh] = ssum > 0.f ? ex / ssum : 0.f;
  }
  return make_float4(w[0], w[1], w[2], w[3]);
}

// v-pass epilogue: weighted column sums over the warp's 32 rows; returns the total of channel 32s + lane
__device__ __forceinline__ float atc_weighted_colsum(uint32_t tmem_base, int q, int s, int lane, float4 w4) {
  uint32_t v[32];
  tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  float val[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float wh = (i < 8) ? w4.x : (i < 16) ? w4.y : (i < 24) ? w4.z : w4.w;
    val[i] = wh * __uint_as_float(v[i]);
  }
  warp_reduce_scatter<32>(val, lane);
  return val[0];
}

__device__ __forceinline__ float sel4(float4 v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// common prologue: barrier, TMEM, W2 image -> smem (one bulk copy), returns the TMEM base
__device__ __forceinline__ uint32_t atc_setup(uint8_t* sW2, const float* W2tc, uint64_t* bars, uint32_t* tmem_slot) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if ((smem_u32(sW2) & 1023u) != 0u) __trap();      // SWIZZLE_128B operands need a 1024-byte aligned base
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);     // W2 landed
    mbar_init(smem_u32(&bars[1]), 1);     // MMAs of a tile retired
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), ATC_W2_BYTES);
    bulk_g2s(smem_u32(sW2), W2tc, ATC_W2_BYTES / 2, smem_u32(&bars[0]));
    bulk_g2s(smem_u32(sW2) + ATC_W2_BYTES / 2, W2tc + ATC_W2_BYTES / 8, ATC_W2_BYTES / 2, smem_u32(&bars[0]));
  }
  return *tmem_slot;
}


static inline int atc_grid(int tiles, int num_sms) { return tiles < num_sms ? (tiles > 0 ? tiles : 1) : num_sms; }

}  // namespace ddb
