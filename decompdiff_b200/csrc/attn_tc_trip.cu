// Tensor-core bond update over triplets k->j->i (BondUpdateLayer, uni_transformer_edge.py:125-167); structure in attn_tc.cuh.
// Two GEMMs per 128-row tile run on tcgen05:
//   D2 = Ang[128 x 16] * Wa^T      the angular-feature term of the first Linear (A2 / B2 in 128B-swizzled smem, SS form)
//   D  = a[128 x 128]  * W2^T      the second Linear on the hidden activations (A in TMEM, TS form)
// so the SIMT side of a row is: gather P[kj] + Q[ji] + D2, LayerNorm, ReLU, TF32 split, and the thread-local epilogue.
#include "attn_tc.cuh"

namespace ddb {

constexpr int TT_A2_BYTES = 2 * 128 * 128;        // hi | lo, [128 rows][128 B] each (16 tf32 used per row)
constexpr int TT_COL_D2 = 384;

struct TripTcSmem {
  uint8_t *W2, *B2, *A2; float *gamma, *beta, *b2, *qry; float2* stat; uint64_t* bars; uint32_t* tmem_slot;
  __device__ explicit TripTcSmem(uint8_t* raw) {
    uint8_t* p = raw;      // purely additive carving keeps everything in the shared address space (LDS / STS)
    W2 = p; p += ATC_W2_BYTES;
    B2 = p; p += TT_A2_BYTES;
    A2 = p; p += TT_A2_BYTES;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    qry = reinterpret_cast<float*>(p); p += 4 * 4 * H * 4;
    stat = reinterpret_cast<float2*>(p); p += 2 * 128 * 4 * 8;      // [parity][row][slice] {sum, sum of squares}
    bars = reinterpret_cast<uint64_t*>(p); p += 32;
    tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  static constexpr int bytes() { return ATC_W2_BYTES + 2 * TT_A2_BYTES + (3 * H + 16 * H + 2 * 2 * 128 * 4) * 4 + 64; }
};
static_assert(TripTcSmem::bytes() <= 232448, "shared memory budget");

// byte offset of feature k (< 32) of row r inside a [128 rows][128 B] K-major SWIZZLE_128B tile
__device__ __forceinline__ int a2_off(int r, int k) { return r * 128 + ((((k >> 2) ^ (r & 7))) << 4) + (k & 3) * 4; }

__device__ __forceinline__ void a2_put(uint8_t* A2, int r, int k, float v) {
  uint32_t hi, lo;
  tf32_split(v, hi, lo);
  *reinterpret_cast<uint32_t*>(A2 + a2_off(r, k)) = hi;
  *reinterpret_cast<uint32_t*>(A2 + TT_A2_BYTES / 2 + a2_off(r, k)) = lo;
}

template <bool VPASS>
__device__ __forceinline__ void trip_epilogue(const TripArgs& a, const TripTcSmem& sm, uint32_t tmem_base, int q, int s, int lane,
                                              int buf, int prev_e, bool prev_ok, int prev_nvalid) {
  if (!VPASS) {
    float4 w4 = atc_logits_softmax(tmem_base, q, s, sm.qry + (buf * 4 + q) * H, prev_ok);
    if (prev_e >= 0) st4(a.wbuf + ((size_t)a.trip_base[prev_e] + lane) * NH + s * 4, w4);
  } else {
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (prev_ok) w4 = ld4(a.wbuf + ((size_t)a.trip_base[prev_e] + lane) * NH + s * 4);
    float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
    if (prev_e >= 0) {
      const int c = s * 32 + lane;
      float upd = prev_nvalid > 0 ? tot + sm.b2[c] : 0.f;
      a.h_bond_out[(size_t)prev_e * H + c] = a.h_bond_in[(size_t)prev_e * H + c] + upd;      // :274
    }
  }
}

constexpr int TT_THREADS = ATC_THREADS + 32;     // 16 worker warps + 1 MMA-issuing warp
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int BAR_A2_READY = 5, BAR_A_READY = 6;  // named barriers: workers arrive, the issuer warp syncs

template <bool VPASS>
__global__ void __launch_bounds__(TT_THREADS, 1) trip_tc_kernel(const TripArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TripTcSmem sm(smem_raw);
  const TripSide& side = VPASS ? a.v : a.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = (warp >> 2) & 3, r = q * 32 + lane;
  // barriers: [0] weights landed, [1] main MMA retired, [2] angular MMA retired
  if ((smem_u32(sm.W2) & 1023u) != 0u) __trap();
  if (tid == 0) {
    mbar_init(smem_u32(&sm.bars[0]), 1); mbar_init(smem_u32(&sm.bars[1]), 1); mbar_init(smem_u32(&sm.bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(sm.tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t bar = smem_u32(&sm.bars[0]);
    mbar_expect_tx(bar, ATC_W2_BYTES + TT_A2_BYTES);
    bulk_g2s(smem_u32(sm.W2), side.W2tc, ATC_W2_BYTES / 2, bar);
    bulk_g2s(smem_u32(sm.W2) + ATC_W2_BYTES / 2, side.W2tc + ATC_W2_BYTES / 8, ATC_W2_BYTES / 2, bar);
    bulk_g2s(smem_u32(sm.B2), side.Watc, TT_A2_BYTES, bar);
  }
  const uint32_t tmem_base = *sm.tmem_slot;
  cta_copy_f4(sm.gamma, side.w.gamma, H);
  cta_copy_f4(sm.beta, side.w.beta, H);
  cta_copy_f4(sm.b2, side.w.b2, H);
  // rows of A2 are 128 bytes but only 16 features are used: clear both images once (features 13..31 stay zero)
  for (int i = tid * 16; i < TT_A2_BYTES; i += TT_THREADS * 16) *reinterpret_cast<float4*>(sm.A2 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  mbar_wait(smem_u32(&sm.bars[0]), 0);
  const uint32_t bar_mma = smem_u32(&sm.bars[1]), bar_ang = smem_u32(&sm.bars[2]);
  const uint32_t w2_smem = smem_u32(sm.W2), a2_smem = smem_u32(sm.A2), b2_smem = smem_u32(sm.B2);
  const int n_tiles = (a.n_bonds + 3) / 4;

  if (warp == 16) {
    // ---------------------------------------------------------------- MMA issuer warp (one elected lane issues)
    // tensor-pipe order: ang(t0), [ang(t1), main(t0)], [ang(t2), main(t1)], ...  - the angular MMA runs one tile ahead
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto issue_ang = [&]() {
      named_sync(BAR_A2_READY, TT_THREADS);          // every worker has written its A2 features and is done with D2
      if (lane == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {             // K = 16 features: two k-steps of 8 inside the first 64 bytes of the rows
          umma_tf32_ss(tmem_base + TT_COL_D2, umma_desc_sw128(a2_smem + ks * 32), umma_desc_sw128(b2_smem + ks * 32), idesc, ks ? 1u : 0u);
          umma_tf32_ss(tmem_base + TT_COL_D2, umma_desc_sw128(a2_smem + TT_A2_BYTES / 2 + ks * 32), umma_desc_sw128(b2_smem + ks * 32), idesc, 1u);
          umma_tf32_ss(tmem_base + TT_COL_D2, umma_desc_sw128(a2_smem + ks * 32), umma_desc_sw128(b2_smem + TT_A2_BYTES / 2 + ks * 32), idesc, 1u);
        }
        umma_commit(bar_ang);
      }
      __syncwarp();
    };
    if (blockIdx.x < n_tiles) issue_ang();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (tile + gridDim.x < n_tiles) issue_ang();
      named_sync(BAR_A_READY, TT_THREADS);           // hidden activations are in TMEM, D of the previous tile is drained
      if (lane == 0) { tc_fence_after(); atc_issue_mma(tmem_base, w2_smem, bar_mma); }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- 16 worker warps: thread = (row r, channel slice s)
    // geometry of a row -> angular features -> A2 (the 13 features are split over the 4 slice-warps), then hand A2 over
    auto features = [&](int2 gm, int2 rm) {
      const bool rowok = rm.y >= 0;
      const float4 xi = ldg4(a.x4 + (size_t)gm.x * 4), xj = ldg4(a.x4 + (size_t)gm.y * 4);
      float dot = 1.f, cn = 0.f;          // invalid / excluded rows: theta = 0
      if (rowok) {
        const float4 xk = ldg4(a.x4 + (size_t)rm.y * 4);
        float ax = xj.x - xi.x, ay = xj.y - xi.y, az = xj.z - xi.z, bx = xk.x - xi.x, by = xk.y - xi.y, bz = xk.z - xi.z;
        float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
        cn = sqrtf(cx * cx + cy * cy + cz * cz);           // |(j-i) x (k-i)|          (:134-137)
        dot = ax * bx + ay * by + az * bz;
      }
      // AngularEncoding [theta, sin(f theta), cos(f theta)], f = [1,2,3,1,1/2,1/3] (common.py:46-54).  sin / cos of theta
      // follow from (cn, dot) directly (cn^2 + dot^2 = |a|^2 |b|^2), multiples and the half angle from the usual identities;
      // only theta itself and theta/3 need atan2f / sincosf
      if (s == 0) {
        a2_put(sm.A2, r, 0, atan2f(cn, dot));
      } else if (s == 3) {
        float sv, cv;
        sincosf(atan2f(cn, dot) * (float)(1.0 / 3.0), &sv, &cv);
        a2_put(sm.A2, r, 6, sv); a2_put(sm.A2, r, 12, cv);
      } else {
        const float n2 = cn * cn + dot * dot;
        const float inv = n2 > 0.f ? 1.0f / sqrtf(n2) : 0.f;
        const float sn = cn * inv, cs = n2 > 0.f ? dot * inv : 1.f;
        if (s == 1) {
          a2_put(sm.A2, r, 1, sn); a2_put(sm.A2, r, 4, sn); a2_put(sm.A2, r, 7, cs); a2_put(sm.A2, r, 10, cs);
          a2_put(sm.A2, r, 2, 2.f * sn * cs); a2_put(sm.A2, r, 8, cs * cs - sn * sn);
          a2_put(sm.A2, r, 3, sn * (3.f - 4.f * sn * sn)); a2_put(sm.A2, r, 9, cs * (4.f * cs * cs - 3.f));
        } else {    // half angle, theta/2 in [0, pi/2]: take the root that does not cancel, derive the other from sin(theta)
          float sh, ch;
          if (cs >= 0.f) { ch = sqrtf(0.5f * (1.f + cs)); sh = sn / (2.f * ch); }
          else { sh = sqrtf(0.5f * (1.f - cs)); ch = sn / (2.f * sh); }
          a2_put(sm.A2, r, 5, sh); a2_put(sm.A2, r, 11, ch);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // A2 was written through the generic proxy
      tc_fence_before();
      named_arrive(BAR_A2_READY, TT_THREADS);
    };
    auto load_meta = [&](int tile, int2& gm, int2& rm) {
      gm = make_int2(0, 0); rm = make_int2(-1, -1);
      const int en = tile * 4 + q;
      if (tile < n_tiles && en < a.n_bonds) { gm = __ldg(a.grp_meta + en); rm = __ldg(a.row_meta + (size_t)en * 32 + lane); }
    };

    int it = 0;
    int prev_e = -1; bool prev_ok = false; int prev_nvalid = 0;
    int2 gm, rm, gm_n, rm_n;
    load_meta(blockIdx.x, gm, rm);
    load_meta(blockIdx.x + gridDim.x, gm_n, rm_n);
    if (blockIdx.x < n_tiles) features(gm, rm);          // prologue: the angular MMA of the first tile
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int e = tile * 4 + q;
      const bool gvalid = e < a.n_bonds;
      const bool rvalid = rm.x >= 0, rowok = rm.y >= 0;          // rowok: valid and k != i (:117-118)
      // all global loads of this row are requested up front (P[kj] slice, Q[ji] slice, the LayerNorm shift, the query)
      float z[32];
      {
        const float* prow = side.P + (size_t)(rvalid ? rm.x : 0) * H + s * 32;
        const float* qrow = side.Q + (size_t)(gvalid ? e : 0) * H + s * 32;
        float4 pv[8], qv[8];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) pv[i4] = rvalid ? ldg4(prow + i4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) qv[i4] = gvalid ? ldg4(qrow + i4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float shift = (gvalid ? __ldg(side.Qm + e) : 0.f) + (rvalid ? __ldg(side.Pm + rm.x) : 0.f);
        float qry_v = 0.f;
        if (!VPASS && gvalid) qry_v = __ldg(a.q + (size_t)e * a.ldq + s * 32 + lane);
        // metadata two tiles ahead
        int2 gm_nn, rm_nn;
        load_meta(tile + 2 * gridDim.x, gm_nn, rm_nn);
        // ---- D2 of this tile was issued one iteration ago
        mbar_wait(bar_ang, it & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + TT_COL_D2 + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!VPASS) sm.qry[((it & 3) * 4 + q) * H + s * 32 + lane] = qry_v;
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          z[i4 * 4 + 0] = ((pv[i4].x + qv[i4].x) - shift) + __uint_as_float(v[i4 * 4 + 0]);
          z[i4 * 4 + 1] = ((pv[i4].y + qv[i4].y) - shift) + __uint_as_float(v[i4 * 4 + 1]);
          z[i4 * 4 + 2] = ((pv[i4].z + qv[i4].z) - shift) + __uint_as_float(v[i4 * 4 + 2]);
          z[i4 * 4 + 3] = ((pv[i4].w + qv[i4].w) - shift) + __uint_as_float(v[i4 * 4 + 3]);
        }
        // ---- features of the NEXT tile -> A2 (D2 and A2 are free again: every worker got here through the wait above)
        if (tile + gridDim.x < n_tiles) features(gm_n, rm_n);
        gm = gm_n; rm = rm_n; gm_n = gm_nn; rm_n = rm_nn;
      }
      // ---- LayerNorm with ONE exchange: the row statistics are taken about the shift Pm[kj] + Qm[ji] (known to every
      // slice without communication), then ReLU
      {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { s1 += z[i]; s2 = fmaf(z[i], z[i], s2); }
        float2* st = sm.stat + ((it & 1) * 128 + r) * 4;
        st[s] = make_float2(s1, s2);
        quad_barrier(q);
        const float4 t01 = *reinterpret_cast<const float4*>(st), t23 = *reinterpret_cast<const float4*>(st + 2);
        const float mu = ((t01.x + t01.z) + (t23.x + t23.z)) * (1.0f / H);
        const float var = fmaxf(((t01.y + t01.w) + (t23.y + t23.w)) * (1.0f / H) - mu * mu, 0.f);
        const float rstd = 1.0f / sqrtf(var + LN_EPS);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 g = ld4(sm.gamma + s * 32 + i4 * 4), b = ld4(sm.beta + s * 32 + i4 * 4);
          z[i4 * 4 + 0] = fmaxf(fmaf((z[i4 * 4 + 0] - mu) * rstd, g.x, b.x), 0.f);
          z[i4 * 4 + 1] = fmaxf(fmaf((z[i4 * 4 + 1] - mu) * rstd, g.y, b.y), 0.f);
          z[i4 * 4 + 2] = fmaxf(fmaf((z[i4 * 4 + 2] - mu) * rstd, g.z, b.z), 0.f);
          z[i4 * 4 + 3] = fmaxf(fmaf((z[i4 * 4 + 3] - mu) * rstd, g.w, b.w), 0.f);
        }
      }
      // ---- epilogue of the previous tile (its main MMA has had this tile's loads / features / LayerNorm to finish)
      if (it > 0) {
        mbar_wait(bar_mma, (it - 1) & 1);
        tc_fence_after();
        trip_epilogue<VPASS>(a, sm, tmem_base, q, s, lane, (it - 1) & 3, prev_e, prev_ok, prev_nvalid);
      }
      // ---- hidden activations -> TMEM; the issuer warp starts the main MMA once every worker has arrived
      {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) tf32_split(rowok ? z[i] : 0.f, hi[i], lo[i]);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        tmem_st32(lane_addr + ATC_COL_AHI + s * 32, hi);
        tmem_st32(lane_addr + ATC_COL_ALO + s * 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        named_arrive(BAR_A_READY, TT_THREADS);
      }
      prev_e = gvalid ? e : -1; prev_ok = rowok; prev_nvalid = __popc(__ballot_sync(FULL, rowok));
    }
    if (it > 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      trip_epilogue<VPASS>(a, sm, tmem_base, q, s, lane, (it - 1) & 3, prev_e, prev_ok, prev_nvalid);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

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
  if (vpass) trip_tc_kernel<true><<<grid, TT_THREADS, bytes, stream>>>(a);
  else trip_tc_kernel<false><<<grid, TT_THREADS, bytes, stream>>>(a);
}

// host-side packing of Wa[13][128] (first-Linear columns of the angular encoding, transposed) into the B operand of the
// angular MMA: rows n = output channel, K = 16 features (13 used), hi | lo, 128-byte rows with the 128B swizzle
void pack_wa_tc(const float* Wa, float* out /* 2*128*32 floats */) {
  for (int i = 0; i < 2 * 128 * 32; ++i) out[i] = 0.f;
  float* hi = out;
  float* lo = out + 128 * 32;
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < NANG; ++k) {
      float w = Wa[k * H + n];
      float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
      int off = (n * 128 + ((((k >> 2) ^ (n & 7))) << 4) + (k & 3) * 4) / 4;
      hi[off] = h; lo[off] = l;
    }
}

}  // namespace ddb
