// Tensor-core attention kernels (sm_100a): the same math as attn_knn.cu / attn_bond.cu, restructured so that the
// second Linear of the key / value MLPs runs on tcgen05 as a true GEMM with shared weights.
//
//   rows   : one attention candidate each (a kNN edge of a destination node / a triplet k->j->i of a bond edge j->i);
//            32 rows form one softmax group = one TMEM lane quadrant, 4 groups form a 128-row tile
//   threads: 16 warps; warp w = 4*s + q owns rows 32q..32q+31 (thread = row) and hidden channels 32s..32s+31
//   hidden : a = ReLU(LN(first-Linear pieces))  computed thread-per-row: no cross-lane reductions, row statistics are
//            exchanged between the 4 slice-warps of a quadrant through shared memory
//   GEMM   : a (hi/lo TF32 split) is written to TMEM with tcgen05.st and used as the A operand of 48 tcgen05.mma
//            (M128 N128 K8, 3xTF32) against W2 (hi/lo, K-major, 128B swizzle) resident in shared memory for the whole
//            persistent kernel; D (128 x 128 fp32) lives in TMEM
//   k pass : logits[row, head] = <q_group[head], D[row, head]>  (thread-local dot products), fused per-group softmax with
//            warp max / sum over the 32 rows, times e_w -> wbuf
//   v pass : out[group, c] = sum_rows w[row, head(c)] D[row, c] + b2[c] sum_rows w[row, head(c)]  via one 32-value
//            butterfly reduce-scatter across the 32 rows
//   The MMA of tile t overlaps the hidden computation of tile t+1; the epilogue of tile t runs right after it.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace ddb {

constexpr int ATC_THREADS = 512;
constexpr int ATC_W2_BYTES = 2 * 128 * 128 * 4;                       // hi | lo image of W2
constexpr int ANG_LD = 20;              // padded row of the per-tile angular features (bank-conflict-free float4 reads)
constexpr int ATC_COL_AHI = 0, ATC_COL_ALO = 128, ATC_COL_D = 256;    // TMEM column map (512 allocated)

__device__ __forceinline__ void quad_barrier(int q) { asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(128) : "memory"); }

// D (+)= A[tmem hi/lo] * W2[smem hi/lo]  : 16 k-steps x 3 MMAs, N = 128
__device__ __forceinline__ void atc_issue_mma(uint32_t tmem_base, uint32_t w2_smem, uint32_t bar) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t d = tmem_base + ATC_COL_D;
  const uint32_t b_hi = w2_smem, b_lo = w2_smem + ATC_W2_BYTES / 2;
#pragma unroll
  for (int kk = 0; kk < 16; ++kk) {
    const uint32_t bo = (kk >> 2) * (128 * 128) + (kk & 3) * 32;
    umma_tf32_ts(d, tmem_base + ATC_COL_AHI + kk * 8, umma_desc_sw128(b_hi + bo), idesc, kk ? 1u : 0u);
    umma_tf32_ts(d, tmem_base + ATC_COL_ALO + kk * 8, umma_desc_sw128(b_hi + bo), idesc, 1u);
    umma_tf32_ts(d, tmem_base + ATC_COL_AHI + kk * 8, umma_desc_sw128(b_lo + bo), idesc, 1u);
  }
  umma_commit(bar);
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
  for (int i = 0; i < 32; ++i) {
    float a = keep ? z[i] : 0.f;
    float h = tf32_rna(a);
    hi[i] = __float_as_uint(h);
    lo[i] = __float_as_uint(tf32_rna(a - h));
  }
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

// ================================================================================================ triplets
struct TripTcSmem {
  uint8_t* W2; float *Wa, *Wc, *gamma, *beta, *b2, *ang, *Q, *qry, *statA, *statB; uint64_t* bars; uint32_t* tmem_slot;
  __device__ explicit TripTcSmem(uint8_t* raw) {
    uint8_t* p = raw;      // dynamic smem is declared __align__(1024); keeping the pointer arithmetic purely additive
                           // lets the compiler keep these in the shared address space (LDS/STS instead of generic LD/ST)
    W2 = p; p += ATC_W2_BYTES;
    Wa = reinterpret_cast<float*>(p); p += 16 * H * 4;
    Wc = reinterpret_cast<float*>(p); p += NG * H * 4;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    ang = reinterpret_cast<float*>(p); p += 128 * ANG_LD * 4;
    Q = reinterpret_cast<float*>(p); p += 4 * H * 4;
    qry = reinterpret_cast<float*>(p); p += 2 * 4 * H * 4;
    statA = reinterpret_cast<float*>(p); p += 128 * 4 * 4;
    statB = reinterpret_cast<float*>(p); p += 128 * 4 * 4;
    bars = reinterpret_cast<uint64_t*>(p); p += 16;
    tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  static constexpr int bytes() {
    return ATC_W2_BYTES + (16 * H + NG * H + 3 * H + 128 * ANG_LD + 4 * H + 8 * H + 2 * 128 * 4) * 4 + 64;
  }
};

template <bool VPASS>
__global__ void __launch_bounds__(ATC_THREADS, 1) trip_tc_kernel(const TripArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TripTcSmem sm(smem_raw);
  const TripSide& side = VPASS ? a.v : a.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = warp >> 2, r = q * 32 + lane;
  const uint32_t tmem_base = atc_setup(sm.W2, side.W2tc, sm.bars, sm.tmem_slot);
  cta_copy_f4(sm.Wa, side.Wa, NANG * H);
  cta_copy_f4(sm.Wc, side.Wc, NG * H);
  cta_copy_f4(sm.gamma, side.w.gamma, H);
  cta_copy_f4(sm.beta, side.w.beta, H);
  cta_copy_f4(sm.b2, side.w.b2, H);
  __syncthreads();
  mbar_wait(smem_u32(&sm.bars[0]), 0);
  const uint32_t bar_mma = smem_u32(&sm.bars[1]), w2_smem = smem_u32(sm.W2);

  const int n_tiles = (a.n_bonds + 3) / 4;
  int it = 0;
  // state of the previous tile (its MMA is in flight while this tile's hidden activations are computed)
  int prev_e = -1; bool prev_ok = false; int prev_nvalid = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    // ---- P0: group (= bond edge j->i) and row (= edge k->j entering j) geometry
    const int e = tile * 4 + q;
    const bool gvalid = e < a.n_bonds;
    int j = 0, i = 0, s_begin = 0, cnt = 0;
    if (gvalid) { j = a.bsrc[e]; i = a.bdst[e]; s_begin = a.in_ptr[j]; cnt = a.in_ptr[j + 1] - s_begin; }
    const bool rvalid = gvalid && lane < cnt;
    int k = j, eid = 0;
    if (rvalid) { k = __ldg(a.in_src + s_begin + lane); eid = __ldg(a.in_eid + s_begin + lane); }
    const bool rowok = rvalid && k != i;                       // i == k triplets are removed (:117-118)
    const float4 xi = ldg4(a.x4 + (size_t)a.lig_idx[i] * 4), xj = ldg4(a.x4 + (size_t)a.lig_idx[j] * 4);
    {
      // AngularEncoding of the angle at i between (j - i) and (k - i); the 5 distinct sin/cos pairs are split over slices
      float th = 0.f;
      if (rowok) {
        const float4 xk = ldg4(a.x4 + (size_t)a.lig_idx[k] * 4);
        float ax = xj.x - xi.x, ay = xj.y - xi.y, az = xj.z - xi.z, bx = xk.x - xi.x, by = xk.y - xi.y, bz = xk.z - xi.z;
        float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
        th = atan2f(sqrtf(cx * cx + cy * cy + cz * cz), ax * bx + ay * by + az * bz);
      }
      float* o = sm.ang + r * ANG_LD;
      float sv, cv;
      if (s == 0) { sincosf(th, &sv, &cv); o[0] = th; o[1] = sv; o[4] = sv; o[7] = cv; o[10] = cv; }
      else if (s == 1) { sincosf(th * 2.f, &sv, &cv); o[2] = sv; o[8] = cv; }
      else if (s == 2) { sincosf(th * 3.f, &sv, &cv); o[3] = sv; o[9] = cv; }
      else {
        sincosf(th * 0.5f, &sv, &cv); o[5] = sv; o[11] = cv;
        sincosf(th * (float)(1.0 / 3.0), &sv, &cv); o[6] = sv; o[12] = cv;
      }
      // Q = Wc . gauss(d_ji) for channel 32s + lane of this group; the query row of the k pass
      float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      float d = sqrtf(dx * dx + dy * dy + dz * dz);
      float gl = lane < NG ? gauss_feat(d, lane) : 0.f;
      float qv = 0.f;
#pragma unroll
      for (int g = 0; g < NG; ++g) qv = fmaf(__shfl_sync(FULL, gl, g), sm.Wc[g * H + s * 32 + lane], qv);
      sm.Q[q * H + s * 32 + lane] = qv;
      if (!VPASS) sm.qry[((it & 1) * 4 + q) * H + s * 32 + lane] = gvalid ? __ldg(a.q + (size_t)e * a.ldq + s * 32 + lane) : 0.f;
    }
    quad_barrier(q);
    // ---- P1: hidden pre-activation of channels [32s, 32s+32) of this row, LayerNorm, ReLU
    float z[32];
    {
      const float* prow = side.P + (size_t)eid * H + s * 32;
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        float4 p = rvalid ? ldg4(prow + i4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 qq = ld4(sm.Q + q * H + s * 32 + i4 * 4);
        z[i4 * 4] = p.x + qq.x; z[i4 * 4 + 1] = p.y + qq.y; z[i4 * 4 + 2] = p.z + qq.z; z[i4 * 4 + 3] = p.w + qq.w;
      }
      float an[16];
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) { float4 t = ld4(sm.ang + r * ANG_LD + i4 * 4); an[i4 * 4] = t.x; an[i4 * 4 + 1] = t.y; an[i4 * 4 + 2] = t.z; an[i4 * 4 + 3] = t.w; }
#pragma unroll
      for (int aa = 0; aa < NANG; ++aa) {
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 w = ld4(sm.Wa + aa * H + s * 32 + i4 * 4);
          z[i4 * 4] = fmaf(an[aa], w.x, z[i4 * 4]); z[i4 * 4 + 1] = fmaf(an[aa], w.y, z[i4 * 4 + 1]);
          z[i4 * 4 + 2] = fmaf(an[aa], w.z, z[i4 * 4 + 2]); z[i4 * 4 + 3] = fmaf(an[aa], w.w, z[i4 * 4 + 3]);
        }
      }
    }
    atc_ln_relu(z, sm.statA, sm.statB, r, s, q, sm.gamma, sm.beta);
    // ---- epilogue of the previous tile (its MMA has had the whole P0/P1 to finish)
    if (it > 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      if (!VPASS) {
        float4 w4 = atc_logits_softmax(tmem_base, q, s, sm.qry + (((it - 1) & 1) * 4 + q) * H, prev_ok);
        if (prev_e >= 0) st4(a.wbuf + ((size_t)a.trip_base[prev_e] + lane) * NH + s * 4, w4);
      } else {
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prev_ok) w4 = ld4(a.wbuf + ((size_t)a.trip_base[prev_e] + lane) * NH + s * 4);
        float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
        if (prev_e >= 0) {
          const int c = s * 32 + lane;
          float upd = prev_nvalid > 0 ? tot + sm.b2[c] : 0.f;
          a.h_bond_out[(size_t)prev_e * H + c] = a.h_bond_in[(size_t)prev_e * H + c] + upd;
        }
      }
    }
    // ---- P2: A -> TMEM, issue the MMAs of this tile
    atc_store_and_mma(z, rowok, tmem_base, q, s, w2_smem, bar_mma);
    prev_e = gvalid ? e : -1; prev_ok = rowok; prev_nvalid = __popc(__ballot_sync(FULL, rowok));
  }
  if (it > 0) {
    mbar_wait(bar_mma, (it - 1) & 1);
    tc_fence_after();
    if (!VPASS) {
      float4 w4 = atc_logits_softmax(tmem_base, q, s, sm.qry + (((it - 1) & 1) * 4 + q) * H, prev_ok);
      if (prev_e >= 0) st4(a.wbuf + ((size_t)a.trip_base[prev_e] + lane) * NH + s * 4, w4);
    } else {
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (prev_ok) w4 = ld4(a.wbuf + ((size_t)a.trip_base[prev_e] + lane) * NH + s * 4);
      float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
      if (prev_e >= 0) {
        const int c = s * 32 + lane;
        float upd = prev_nvalid > 0 ? tot + sm.b2[c] : 0.f;
        a.h_bond_out[(size_t)prev_e * H + c] = a.h_bond_in[(size_t)prev_e * H + c] + upd;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static int atc_grid(int tiles, int num_sms) { return tiles < num_sms ? (tiles > 0 ? tiles : 1) : num_sms; }

void launch_trip_tc(const TripArgs& a, bool vpass, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0) return;
  static bool once = false;
  const int bytes = TripTcSmem::bytes();
  if (!once) {
    cudaFuncSetAttribute(trip_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(trip_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    once = true;
  }
  const int grid = atc_grid((a.n_bonds + 3) / 4, num_sms);
  if (vpass) trip_tc_kernel<true><<<grid, ATC_THREADS, bytes, stream>>>(a);
  else trip_tc_kernel<false><<<grid, ATC_THREADS, bytes, stream>>>(a);
}

// ================================================================================================ kNN edges
struct KnnTcSmem {
  uint8_t* W2; float *Wg, *Wt, *gamma, *beta, *b2, *G, *Hi, *qry, *statA, *statB; uint64_t* bars; uint32_t* tmem_slot;
  __device__ explicit KnnTcSmem(uint8_t* raw) {
    uint8_t* p = raw;
    W2 = p; p += ATC_W2_BYTES;
    Wg = reinterpret_cast<float*>(p); p += 4 * NG * H * 4;
    Wt = reinterpret_cast<float*>(p); p += 4 * H * 4;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    G = reinterpret_cast<float*>(p); p += 128 * NG * 4;
    Hi = reinterpret_cast<float*>(p); p += 4 * H * 4;
    qry = reinterpret_cast<float*>(p); p += 2 * 4 * H * 4;
    statA = reinterpret_cast<float*>(p); p += 128 * 4 * 4;
    statB = reinterpret_cast<float*>(p); p += 128 * 4 * 4;
    bars = reinterpret_cast<uint64_t*>(p); p += 16;
    tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  static constexpr int bytes() {
    return ATC_W2_BYTES + (4 * NG * H + 4 * H + 3 * H + 128 * NG + 4 * H + 8 * H + 2 * 128 * 4) * 4 + 64;
  }
};
static_assert(KnnTcSmem::bytes() <= 232448 && TripTcSmem::bytes() <= 232448, "shared memory budget");

template <bool VPASS>
__global__ void __launch_bounds__(ATC_THREADS, 1) knn_tc_kernel(const KnnAttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  KnnTcSmem sm(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = warp >> 2, r = q * 32 + lane;
  const uint32_t tmem_base = atc_setup(sm.W2, a.W2tc, sm.bars, sm.tmem_slot);
  cta_copy_f4(sm.Wg, a.w.Wg, 4 * NG * H);
  cta_copy_f4(sm.Wt, a.w.Wt, 4 * H);
  cta_copy_f4(sm.gamma, a.w.gamma, H);
  cta_copy_f4(sm.beta, a.w.beta, H);
  if (VPASS) cta_copy_f4(sm.b2, a.w.b2, H);
  __syncthreads();
  mbar_wait(smem_u32(&sm.bars[0]), 0);
  const uint32_t bar_mma = smem_u32(&sm.bars[1]), w2_smem = smem_u32(sm.W2);

  const int n_tiles = (a.n_dst + 3) / 4;
  int it = 0;
  int prev_node = -1; bool prev_ok = false; float prev_ew = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    // ---- P0: group = destination node, row = one of its (<= 32) incoming kNN edges
    const int slot = tile * 4 + q;
    const bool gvalid = slot < a.n_dst;
    int node = 0, deg = 0, nlig = 0; bool lig_dst = false;
    if (gvalid) { node = a.dst_list ? a.dst_list[slot] : slot; deg = a.deg[node]; nlig = a.nlig[node]; lig_dst = a.is_lig[node]; }
    const bool rowok = gvalid && lane < deg;
    const int j = rowok ? __ldg(a.nbr + (size_t)node * KNN + lane) : node;
    const int type = lig_dst ? (lane < nlig ? 0 : 2) : (lane < nlig ? 1 : 3);       // uni_transformer_edge.py:371-377
    {
      const float4 xi = ldg4(a.x4 + (size_t)node * 4), xj = ldg4(a.x4 + (size_t)j * 4);
      float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
      float d = sqrtf(dx * dx + dy * dy + dz * dz);
#pragma unroll
      for (int g = 0; g < NG / 4; ++g) sm.G[r * NG + s * (NG / 4) + g] = gauss_feat(d, s * (NG / 4) + g);
      sm.Hi[q * H + s * 32 + lane] = gvalid ? __ldg(a.Hi + (size_t)(a.hi_by_slot ? slot : node) * a.ldhi + s * 32 + lane) : 0.f;
      if (!VPASS) sm.qry[((it & 1) * 4 + q) * H + s * 32 + lane] =
          gvalid ? __ldg(a.q + (size_t)(a.q_by_slot ? slot : node) * a.ldq + s * 32 + lane) : 0.f;
    }
    quad_barrier(q);
    // ---- P1
    float z[32];
    {
      const float* hj = a.Hj + (size_t)j * a.ldhj + s * 32;
      const float* wt = sm.Wt + type * H + s * 32;
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        const float4 p = ldg4(hj + i4 * 4), hh = ld4(sm.Hi + q * H + s * 32 + i4 * 4), t = ld4(wt + i4 * 4);
        z[i4 * 4] = (p.x + hh.x) + t.x; z[i4 * 4 + 1] = (p.y + hh.y) + t.y; z[i4 * 4 + 2] = (p.z + hh.z) + t.z; z[i4 * 4 + 3] = (p.w + hh.w) + t.w;
      }
      const float* wg = sm.Wg + (size_t)type * NG * H + s * 32;
#pragma unroll
      for (int gb = 0; gb < NG / 4; ++gb) {
        const float4 g4 = ld4(sm.G + r * NG + gb * 4);
#pragma unroll
        for (int gc = 0; gc < 4; ++gc) {
          const float gv = sel4(g4, gc);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 w = ld4(wg + (gb * 4 + gc) * H + i4 * 4);
            z[i4 * 4] = fmaf(gv, w.x, z[i4 * 4]); z[i4 * 4 + 1] = fmaf(gv, w.y, z[i4 * 4 + 1]);
            z[i4 * 4 + 2] = fmaf(gv, w.z, z[i4 * 4 + 2]); z[i4 * 4 + 3] = fmaf(gv, w.w, z[i4 * 4 + 3]);
          }
        }
      }
    }
    atc_ln_relu(z, sm.statA, sm.statB, r, s, q, sm.gamma, sm.beta);
    // ---- epilogue of the previous tile
    if (it > 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      if (!VPASS) {
        float4 w4 = atc_logits_softmax(tmem_base, q, s, sm.qry + (((it - 1) & 1) * 4 + q) * H, prev_ok);
        if (prev_ok) st4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + s * 4,
                         make_float4(w4.x * prev_ew, w4.y * prev_ew, w4.z * prev_ew, w4.w * prev_ew));
      } else {
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prev_ok) w4 = ld4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + s * 4);
        float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
        float4 ws = make_float4(warp_sum(w4.x), warp_sum(w4.y), warp_sum(w4.z), warp_sum(w4.w));
        if (prev_node >= 0) {
          const int c = s * 32 + lane;
          a.out_h[(size_t)prev_node * a.ldo + c] = tot + sm.b2[c] * sel4(ws, lane >> 3);
        }
      }
    }
    atc_store_and_mma(z, rowok, tmem_base, q, s, w2_smem, bar_mma);
    prev_node = gvalid ? node : -1; prev_ok = rowok;
    prev_ew = (!VPASS && rowok) ? __ldg(a.e_w + (size_t)node * KNN + lane) : 0.f;
  }
  if (it > 0) {
    mbar_wait(bar_mma, (it - 1) & 1);
    tc_fence_after();
    if (!VPASS) {
      float4 w4 = atc_logits_softmax(tmem_base, q, s, sm.qry + (((it - 1) & 1) * 4 + q) * H, prev_ok);
      if (prev_ok) st4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + s * 4,
                       make_float4(w4.x * prev_ew, w4.y * prev_ew, w4.z * prev_ew, w4.w * prev_ew));
    } else {
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (prev_ok) w4 = ld4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + s * 4);
      float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
      float4 ws = make_float4(warp_sum(w4.x), warp_sum(w4.y), warp_sum(w4.z), warp_sum(w4.w));
      if (prev_node >= 0) {
        const int c = s * 32 + lane;
        a.out_h[(size_t)prev_node * a.ldo + c] = tot + sm.b2[c] * sel4(ws, lane >> 3);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

void launch_knn_tc(const KnnAttnArgs& a, bool vpass, int num_sms, cudaStream_t stream) {
  if (a.n_dst <= 0) return;
  static bool once = false;
  const int bytes = KnnTcSmem::bytes();
  if (!once) {
    cudaFuncSetAttribute(knn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(knn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    once = true;
  }
  const int grid = atc_grid((a.n_dst + 3) / 4, num_sms);
  if (vpass) knn_tc_kernel<true><<<grid, ATC_THREADS, bytes, stream>>>(a);
  else knn_tc_kernel<false><<<grid, ATC_THREADS, bytes, stream>>>(a);
}

// host-side packing of a second-Linear weight W2[128 out][128 in] (already scaled) into the hi | lo swizzled image
void pack_w2_tc(const float* W2, float* out /* 2*128*128 floats */) {
  float* hi = out;
  float* lo = out + 128 * 128;
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 128; ++k) {
      float w = W2[n * 128 + k];
      float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
      int off = sw128_offset_bytes(n, k, 128) / 4;
      hi[off] = h; lo[off] = l;
    }
}

}  // namespace ddb
