// Tensor-core bond update over triplets k->j->i (BondUpdateLayer, uni_transformer_edge.py:125-167); structure in attn_tc.cuh.
// Two GEMMs per 128-row tile run on tcgen05:
//   D2 = Ang[128 x 16] * Wa^T      the angular-feature term of the first Linear (A2 / B2 in 128B-swizzled smem, SS form)
//   D  = a[128 x 128]  * W2^T      the second Linear on the hidden activations (A in TMEM, TS form)
// so the SIMT side of a row is: P'[kj] (kept in registers while the source atom j is unchanged - groups are visited source-major)
// + Q'[ji] (staged per warp) + D2, LayerNorm, ReLU, TF32 split, and the thread-local epilogue.
#include "attn_tc.cuh"

namespace ddb {

constexpr int TT_A2_BYTES = 2 * 128 * 128;        // hi | lo, [128 rows][128 B] each (16 tf32 used per row)
constexpr int TT_COL_D2 = 384;

struct TripTcSmem {
  uint8_t *W2, *B2, *A2; float *gamma, *beta, *b2, *qry, *qrow, *xyz; float2* stat; uint64_t* bars; uint32_t* tmem_slot;
  __device__ explicit TripTcSmem(uint8_t* raw) {
    uint8_t* p = raw;      // purely additive carving keeps everything in the shared address space (LDS / STS)
    W2 = p; p += ATC_W2_BYTES;
    B2 = p; p += TT_A2_BYTES;
    A2 = p; p += TT_A2_BYTES;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    qry = reinterpret_cast<float*>(p); p += 16 * 128 * 4;       // per warp: 4-deep ring of 32-float query slices
    qrow = reinterpret_cast<float*>(p); p += 16 * 64 * 4;       // per warp: 2 x 32-float slices of the centred Q row
    xyz = reinterpret_cast<float*>(p); p += 16 * 34 * 16;       // per warp: positions x_k of its 32 rows, x_i, x_j (cp.async staging)
    stat = reinterpret_cast<float2*>(p); p += 2 * 128 * 4 * 8;      // [parity][slice][row] {sum, sum of squares}
    bars = reinterpret_cast<uint64_t*>(p); p += 128;       // two sets of 8: the second phase of a paired launch uses its own
    tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  static constexpr int bytes() { return ATC_W2_BYTES + 2 * TT_A2_BYTES + (3 * H + 16 * H + 16 * 64 + 16 * 34 * 4 + 2 * 2 * 128 * 4) * 4 + 160; }
};
static_assert(TripTcSmem::bytes() <= 232448, "shared memory budget");

// byte offset of feature k (< 32) of row r inside a [128 rows][128 B] K-major SWIZZLE_128B tile
__device__ __forceinline__ int a2_off(int r, int k) { return r * 128 + ((((k >> 2) ^ (r & 7))) << 4) + (k & 3) * 4; }

// four consecutive features (chunk c = k / 4) of row r -> both images with one 16-byte store each (conflict-free: the swizzle
// spreads 8 consecutive rows over the 8 chunks of a 128-byte line; single 4-byte stores were 4-way bank conflicts)
__device__ __forceinline__ void a2_put4(uint8_t* A2, int r, int c, float v0, float v1, float v2, float v3) {
  uint4 hi, lo;
  tf32_split(v0, hi.x, lo.x); tf32_split(v1, hi.y, lo.y); tf32_split(v2, hi.z, lo.z); tf32_split(v3, hi.w, lo.w);
  const int off = r * 128 + ((c ^ (r & 7)) << 4);
  *reinterpret_cast<uint4*>(A2 + off) = hi;
  *reinterpret_cast<uint4*>(A2 + TT_A2_BYTES / 2 + off) = lo;
}


// ---- packed fp32 (FADD2 / FMUL2 / FFMA2 on sm_100): the element-wise phases work on channel pairs
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 u2f(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// 16 worker warps + one warpgroup whose first warp issues the MMAs (tcgen05.mma blocks its issuing thread while the tensor
// queue is full, so the issuer must not be a worker).  setmaxnreg moves the registers of the idle warps to the workers.
constexpr int TT_THREADS = ATC_THREADS + 128;
constexpr int TT_ISSUER = 16;
#ifndef TT_EARLY_PREFETCH
#define TT_EARLY_PREFETCH 0
#endif
constexpr int TT_SYNC = ATC_THREADS + 32;       // participants of the hand-over barriers: 512 workers + the issuing warp
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
constexpr int BAR_A_READY = 6;      // named barrier: every worker warp arrives, the issuing warp waits on it

// `first` / `last`: key pass and value pass may run back to back inside one launch (trip_tc_pair_kernel, see attn_tc_knn.cu)
template <bool VPASS>
__device__ __forceinline__ void trip_tc_body(const TripArgs& a, const bool first, const bool last) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  TripTcSmem sm(smem_raw);
  uint64_t* const bars = sm.bars + (first ? 0 : 8);      // a fresh barrier set per phase (no re-initialisation of used barriers)
  const TripSide& side = VPASS ? a.v : a.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = (warp >> 2) & 3, r = q * 32 + lane;
  // barriers: [0] weights landed, [1] main MMA retired, [2] angular MMA retired
  if ((smem_u32(sm.W2) & 1023u) != 0u) __trap();
  if (tid == 0) {
    // [0] weights landed, [1] main MMA retired, [2] angular MMA retired, [3] angular features of a tile written (3 producer warps),
    // [4] D2 of a tile read by every worker warp
    for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), i == 3 ? 3 : i == 4 ? 16 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (first && warp == 0) { __syncwarp(); tmem_alloc(smem_u32(sm.tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
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
  if (first) pdl_wait();      // set-up on static data above; the previous kernels' results are visible below
  __syncthreads();
  mbar_wait(smem_u32(&bars[0]), 0);
  const uint32_t bar_mma = smem_u32(&bars[1]), bar_ang = smem_u32(&bars[2]), bar_a2f = smem_u32(&bars[3]), bar_d2c = smem_u32(&bars[4]);
  const uint32_t w2_smem = smem_u32(sm.W2), a2_smem = smem_u32(sm.A2), b2_smem = smem_u32(sm.B2);
  // Work distribution: the groups (bond edges j->i) are visited in SOURCE-major order (a.grp_order) and every (CTA, quadrant)
  // pair walks one contiguous chunk of that order.  All groups with the same source j read the same rows P'[k->j], so a thread
  // keeps its row slice in registers and gathers again only when j changes (once per ~n_lig groups).
  const int per = (a.n_groups + 4 * (int)gridDim.x - 1) / (4 * (int)gridDim.x);      // iterations of every CTA

  if (warp >= 16) {
    // ---------------------------------------------------------------- warp 16 issues the MMAs, warps 17..19 produce the angular features
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
#endif
    if (warp == TT_ISSUER) {
      // tensor-pipe order: ang(t0), [ang(t1), main(t0)], [ang(t2), main(t1)], ...  - the angular MMA runs one tile ahead
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      auto issue_ang = [&](int t) {
        if (lane == 0) {
          mbar_wait(bar_a2f, t & 1);                   // the producers have written the features of tile t
          if (t > 0) mbar_wait(bar_d2c, (t - 1) & 1);   // every worker warp has read D2 of tile t-1
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
      if (per > 0) issue_ang(0);
      for (int it = 0; it < per; ++it) {
        if (it + 1 < per) issue_ang(it + 1);
        named_sync(BAR_A_READY, TT_SYNC);           // hidden activations are in TMEM, D of the previous tile is in registers
        if (lane == 0) { tc_fence_after(); atc_issue_mma(tmem_base, w2_smem, bar_mma); }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- producers: geometry of a row -> its 13 angular features -> A2.
      // Thread = row; warp 17 serves quadrants 0 and 3, warps 18 / 19 quadrants 1 / 2.  Everything here used to sit on the worker
      // warps' critical path (~125 instructions per thread and tile, the slice with atan2f + sincosf the slowest of its quadrant).
      const int pw = warp - 17;
      auto row_meta_of = [&](int qq, int t, int2& gm, int2& rm) {
        gm = make_int2(0, 0); rm = make_int2(-1, -1);
        const int gb = ((int)blockIdx.x * 4 + qq) * per, pos = gb + t;
        if (t < per && pos < min(a.n_groups, gb + per)) { gm = __ldg(a.grp_meta + pos); rm = __ldg(a.row_meta + (size_t)pos * 32 + lane); }
      };
      auto put_features = [&](int qq, int2 rm, float4 xi, float4 xj, float4 xk) {
        const int rr = qq * 32 + lane;
        const bool rowok = rm.y >= 0;
        const float ax = xj.x - xi.x, ay = xj.y - xi.y, az = xj.z - xi.z, bx = xk.x - xi.x, by = xk.y - xi.y, bz = xk.z - xi.z;
        const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
        float cn = sqrtf(cx * cx + cy * cy + cz * cz);             // |(j-i) x (k-i)|          (:134-137)
        float dot = ax * bx + ay * by + az * bz;
        if (!rowok) { cn = 0.f; dot = 1.f; }
        // AngularEncoding [theta, sin(f theta), cos(f theta)], f = [1,2,3,1,1/2,1/3] (common.py:46-54): sin / cos of theta follow
        // from (cn, dot), multiples and the half angle from the usual identities; only theta and theta / 3 need atan2f / sincosf
        const float theta = atan2f(cn, dot);
        const float n2 = cn * cn + dot * dot;
        const float inv = n2 > 0.f ? rsqrtf(n2) : 0.f;
        const float sn = cn * inv, cs = n2 > 0.f ? dot * inv : 1.f;
        float sh, ch, s3, c3;
        if (cs >= 0.f) { ch = sqrtf(0.5f * (1.f + cs)); sh = sn / (2.f * ch); }
        else { sh = sqrtf(0.5f * (1.f - cs)); ch = sn / (2.f * sh); }
        sincosf(theta * (float)(1.0 / 3.0), &s3, &c3);
        a2_put4(sm.A2, rr, 0, theta, sn, 2.f * sn * cs, sn * (3.f - 4.f * sn * sn));
        a2_put4(sm.A2, rr, 1, sn, sh, s3, cs);
        a2_put4(sm.A2, rr, 2, cs * cs - sn * sn, cs * (4.f * cs * cs - 3.f), cs, ch);
        a2_put4(sm.A2, rr, 3, c3, 0.f, 0.f, 0.f);
      };
      const int q0 = pw == 0 ? 0 : pw, q1 = 3;          // warp 17 also serves quadrant 3
      int2 gm0, rm0, gm1, rm1;
      row_meta_of(q0, 0, gm0, rm0);
      if (pw == 0) row_meta_of(q1, 0, gm1, rm1);
      for (int t = 0; t < per; ++t) {
        // positions of this tile's rows (requested first), metadata of the next tile
        const float4 xi0 = ldg4(a.x4 + (size_t)gm0.x * 4), xj0 = ldg4(a.x4 + (size_t)gm0.y * 4),
                     xk0 = ldg4(a.x4 + (size_t)(rm0.y >= 0 ? rm0.y : gm0.y) * 4);
        float4 xi1 = xi0, xj1 = xj0, xk1 = xk0;
        if (pw == 0) {
          xi1 = ldg4(a.x4 + (size_t)gm1.x * 4); xj1 = ldg4(a.x4 + (size_t)gm1.y * 4);
          xk1 = ldg4(a.x4 + (size_t)(rm1.y >= 0 ? rm1.y : gm1.y) * 4);
        }
        const int2 rm0c = rm0, rm1c = rm1;
        row_meta_of(q0, t + 1, gm0, rm0);
        if (pw == 0) row_meta_of(q1, t + 1, gm1, rm1);
        if (t > 0) mbar_wait(bar_ang, (t - 1) & 1);      // the angular MMA of the previous tile has read A2
#ifndef DDB_EXP_NOFEAT      // timing experiment only
        put_features(q0, rm0c, xi0, xj0, xk0);
        if (pw == 0) put_features(q1, rm1c, xi1, xj1, xk1);
#endif
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // A2 was written through the generic proxy
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a2f) : "memory");
      }
    }
  } else {
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
#endif
    auto hand_over_a = [&]() { named_arrive(BAR_A_READY, TT_SYNC); };
    // ---------------------------------------------------------------- 16 warps: thread = (row r, channel slice s)
    // metadata of a tile: clamped so that every load below is unconditional (padding rows read edge 0 / node 0; their
    // results are never stored and they get zero attention weight)
    const int g_begin = ((int)blockIdx.x * 4 + q) * per, g_end = min(a.n_groups, g_begin + per);
    auto load_meta = [&](int i, int& e, int2& gm, int2& rm) {      // metadata is stored in visiting order: three independent loads
      e = -1; gm = make_int2(0, 0); rm = make_int2(-1, -1);
      const int pos = g_begin + i;
      if (pos < g_end) { e = __ldg(a.grp_order + pos); gm = __ldg(a.grp_meta + pos); rm = __ldg(a.row_meta + (size_t)pos * 32 + lane); }
    };
    float* const wq = sm.qrow + warp * 64;          // this warp's private staging: [parity][32] slice of the Q row
    float* const wqry = sm.qry + warp * 128;        // k pass: [4-deep ring][32] slice of the query row
    const float* __restrict__ Pc = side.P;          // centred rows (trip_prep): LayerNorm is shift invariant
    const float* __restrict__ Qc = side.Q;

    // softmax over the 32 rows of the previous group for heads 4s..4s+3 -> wbuf; chunked groups also record {max, sum of exp}
    auto finish_k = [&](const float (&lg)[4], bool ok, int pe, int tb, int pair) {
      float ex[4], mx[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mx[hh]) : "f"(lg[hh]));
        ex[hh] = ok ? __expf(lg[hh] - mx[hh]) : 0.f;
      }
      float sum[4] = {ex[0], ex[1], ex[2], ex[3]};
      warp_allreduce4(sum, lane);
      float w[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) w[hh] = sum[hh] > 0.f ? __fdividef(ex[hh], sum[hh]) : 0.f;
      if (pe >= 0) st4(a.wbuf + ((size_t)tb + lane) * NH + s * 4, make_float4(w[0], w[1], w[2], w[3]));
      if (pe >= 0 && pair >= 0 && lane < 4)
        a.stats[(size_t)(tb >> 5) * NH + s * 4 + lane] = make_float2(lane == 0 ? mx[0] : lane == 1 ? mx[1] : lane == 2 ? mx[2] : mx[3],
                                                                     lane == 0 ? sum[0] : lane == 1 ? sum[1] : lane == 2 ? sum[2] : sum[3]);
    };
    int it = 0;
    int prev_e = -1, prev_tb = 0, prev_pair = -1; bool prev_ok = false; int prev_nvalid = 0;
    int2 gm, rm, gm_n, rm_n;
    int e, e_n;
    load_meta(0, e, gm, rm);
    load_meta(1, e_n, gm_n, rm_n);
    float4 pv[8];
    if (per > 0) {
      const float* prow = Pc + (size_t)max(rm.x, 0) * H + s * 32;
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) ldg8(prow + i8 * 8, pv[2 * i8], pv[2 * i8 + 1]);
      wq[lane] = __ldg(Qc + (size_t)max(e, 0) * H + s * 32 + lane);
    }
    TL_DECL
    for (; it < per; ++it) {
      TL_MARK(0);
      const bool rowok = rm.y >= 0;          // valid and k != i (:117-118)
      // requests for later: the Q slice of the next group, the query slice of this one, metadata two groups ahead
      const float q_next = __ldg(Qc + (size_t)max(e_n, 0) * H + s * 32 + lane);
      float qry_v = 0.f;
      if (!VPASS) qry_v = __ldg(a.q + (size_t)max(e, 0) * a.ldq + s * 32 + lane);
      int2 gm_nn, rm_nn;
      int e_nn;
      load_meta(it + 2, e_nn, gm_nn, rm_nn);
      const int tb = (g_begin + it) * 32;      // wbuf rows of a (group, chunk) are its 32 slots in visiting order
      const int pair = e >= 0 ? __ldg(a.vg_pair + g_begin + it) : -1;
      // ---- first Linear: z = P'[kj] (in registers) + Q'[ji] (staged) + D2 (angular MMA, issued one iteration ago)
      float2 z[16];
      {
        mbar_wait(bar_ang, it & 1);
        TL_MARK(1);
        tc_fence_after();
        __syncwarp();
        const float* qs = wq + (it & 1) * 32;
        uint32_t v[32];      // D2 first: z[i] is born as v[i] dies (z = P' + Q' before the load kept 96 registers live)
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + TT_COL_D2 + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 qv = ld4(qs + i4 * 4);
          z[i4 * 2] = __fadd2_rn(__fadd2_rn(f2(pv[i4].x, pv[i4].y), f2(qv.x, qv.y)), u2f(v[4 * i4], v[4 * i4 + 1]));
          z[i4 * 2 + 1] = __fadd2_rn(__fadd2_rn(f2(pv[i4].z, pv[i4].w), f2(qv.z, qv.w)), u2f(v[4 * i4 + 2], v[4 * i4 + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_d2c) : "memory");      // D2 may be overwritten
      }
      // ---- prefetch the P' rows of the next tile (consumed one iteration from now), stage its Q slice and this tile's query
      // ---- the next group reads other rows only when its source atom differs: gather them now, consumed one iteration later
      if (__any_sync(FULL, rm_n.x != rm.x)) {
        const float* prow_next = Pc + (size_t)max(rm_n.x, 0) * H + s * 32;
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) ldg8(prow_next + i8 * 8, pv[2 * i8], pv[2 * i8 + 1]);
      }
      {
        wq[((it + 1) & 1) * 32 + lane] = q_next;
        if (!VPASS) wqry[(it & 3) * 32 + lane] = qry_v;
      }
      TL_MARK(2);
      TL_MARK(3);
      // ---- LayerNorm with ONE exchange (single-pass statistics: the rows are centred up to the small angular term), ReLU
      {
        float2 s1 = f2(0.f, 0.f), s2 = f2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) { s1 = __fadd2_rn(s1, z[i]); s2 = __ffma2_rn(z[i], z[i], s2); }
        float2* st = sm.stat + (it & 1) * 512 + r;      // [parity][slice][row]: every access below is a contiguous 256 bytes per warp
        st[s * 128] = make_float2(s1.x + s1.y, s2.x + s2.y);
        TL_MARK(4);
        quad_barrier(q);
        TL_MARK(5);
        const float2 t0 = st[0], t1 = st[128], t2 = st[256], t3 = st[384];
        const float mu = ((t0.x + t1.x) + (t2.x + t3.x)) * (1.0f / H);
        const float var = fmaxf(((t0.y + t1.y) + (t2.y + t3.y)) * (1.0f / H) - mu * mu, 0.f);
        const float rstd = rsqrtf(var + LN_EPS);
        const float2 rs2 = f2(rstd, rstd), nm2 = f2(-mu * rstd, -mu * rstd);
#ifdef DDB_EXP_NONORM      // timing experiment only: rstd is never negative
        if (rstd < 0.f)
#endif
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 g = ld4(sm.gamma + s * 32 + i4 * 4), b = ld4(sm.beta + s * 32 + i4 * 4);
          float2 u0 = __ffma2_rn(z[i4 * 2], rs2, nm2), u1 = __ffma2_rn(z[i4 * 2 + 1], rs2, nm2);      // (z - mu) * rstd
          u0 = __ffma2_rn(u0, f2(g.x, g.y), f2(b.x, b.y));
          u1 = __ffma2_rn(u1, f2(g.z, g.w), f2(b.z, b.w));
          z[i4 * 2] = f2(fmaxf(u0.x, 0.f), fmaxf(u0.y, 0.f));
          z[i4 * 2 + 1] = f2(fmaxf(u1.x, 0.f), fmaxf(u1.y, 0.f));
        }
      }
      // ---- drain D of the previous tile into registers (its main MMA had this tile's first Linear / LayerNorm to finish)
      float lg[4] = {0.f, 0.f, 0.f, 0.f};
      float val[32];
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), f4 = make_float4(1.f, 1.f, 1.f, 1.f);
      float hb_in = 0.f;
      if (VPASS && it > 0) {      // requested before the wait on the tensor core: attention weights of this thread's row, residual input
        if (prev_ok) {           // (the chunk -> whole-group factor is applied after the wait: a multiply here would stall on the loads)
          w4 = ld4(a.wbuf + ((size_t)prev_tb + lane) * NH + s * 4);
          if (prev_pair >= 0) f4 = ld4(a.factor + (size_t)(prev_tb >> 5) * NH + s * 4);
        }
        if (prev_e >= 0) hb_in = __ldg(a.h_bond_in + (size_t)prev_e * H + s * 32 + lane);
      }
      TL_MARK(6);
      if (it > 0) {
        mbar_wait(bar_mma, (it - 1) & 1);
        TL_MARK(7);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!VPASS) {
          const float* qr = wqry + ((it - 1) & 3) * 32;
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            const float4 q0 = ld4(qr + hh * 8), q1 = ld4(qr + hh * 8 + 4);
            float2 acc = __fmul2_rn(f2(q0.x, q0.y), u2f(v[hh * 8], v[hh * 8 + 1]));
            acc = __ffma2_rn(f2(q0.z, q0.w), u2f(v[hh * 8 + 2], v[hh * 8 + 3]), acc);
            acc = __ffma2_rn(f2(q1.x, q1.y), u2f(v[hh * 8 + 4], v[hh * 8 + 5]), acc);
            acc = __ffma2_rn(f2(q1.z, q1.w), u2f(v[hh * 8 + 6], v[hh * 8 + 7]), acc);
            lg[hh] = prev_ok ? acc.x + acc.y : -INFINITY;
          }
        } else {
          w4 = mul4(w4, f4);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float wh = (i < 8) ? w4.x : (i < 16) ? w4.y : (i < 24) ? w4.z : w4.w;
            val[i] = wh * __uint_as_float(v[i]);
          }
          // reduce over the group's rows right away: 32 live values become one before the TF32 split below needs its registers
          // (holding them across the split cost ~50 spill stores per thread and tile; the tensor pipe has the slack)
          warp_reduce_scatter<32>(val, lane);
        }
      }
      // ---- hidden activations -> TMEM (D is in registers, so the issuer may start the main MMA right away)
      {
        // hi = z truncated to TF32, lo = z - hi (exact); 16 columns at a time keeps the register peak low
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 zz = z[half * 8 + i];
            hi[2 * i] = __float_as_uint(zz.x) & 0xffffe000u;
            hi[2 * i + 1] = __float_as_uint(zz.y) & 0xffffe000u;
            const float2 l = __fadd2_rn(zz, f2(-__uint_as_float(hi[2 * i]), -__uint_as_float(hi[2 * i + 1])));
            lo[2 * i] = __float_as_uint(l.x); lo[2 * i + 1] = __float_as_uint(l.y);
          }
          tmem_st16(lane_addr + ATC_COL_AHI + s * 32 + half * 16, hi);
          tmem_st16(lane_addr + ATC_COL_ALO + s * 32 + half * 16, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        TL_MARK(8);
        hand_over_a();
        TL_MARK(9);
      }
      // ---- finish the epilogue of the previous tile from registers while the tensor core works
      if (it > 0) {
        if (!VPASS) {
          // softmax over the 32 rows of the group, 4 heads at once: max by one REDUX each, sums by a transposed all-reduce
          finish_k(lg, prev_ok, prev_e, prev_tb, prev_pair);
        } else {
          if (prev_e >= 0) {
            const int c = s * 32 + lane;
            if (prev_pair >= 0) {
              a.part[(size_t)(prev_tb >> 5) * H + c] = val[0];      // chunked group: launch_trip_combine finishes the edge
            } else {
              const float upd = prev_nvalid > 0 ? val[0] + sm.b2[c] : 0.f;
              a.h_bond_out[(size_t)prev_e * H + c] = hb_in + upd;      // :274
            }
          }
        }
      }
      prev_e = e; prev_ok = rowok; prev_nvalid = __popc(__ballot_sync(FULL, rowok)); prev_tb = tb; prev_pair = pair;
      e = e_n; gm = gm_n; rm = rm_n; e_n = e_nn; gm_n = gm_nn; rm_n = rm_nn;
      TL_MARK(10);
    }
    TL_FLUSH(VPASS ? 1 : 0);
    // ---- epilogue of the last tile
    if (it > 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      if (!VPASS) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float* qr = wqry + ((it - 1) & 3) * 32;
        float lg[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < 8; ++d) acc = fmaf(qr[hh * 8 + d], __uint_as_float(v[hh * 8 + d]), acc);
          lg[hh] = prev_ok ? acc : -INFINITY;
        }
        finish_k(lg, prev_ok, prev_e, prev_tb, prev_pair);
      } else {
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prev_ok) {
          w4 = ld4(a.wbuf + ((size_t)prev_tb + lane) * NH + s * 4);
          if (prev_pair >= 0) w4 = mul4(w4, ld4(a.factor + (size_t)(prev_tb >> 5) * NH + s * 4));
        }
        float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
        if (prev_e >= 0) {
          const int c = s * 32 + lane;
          if (prev_pair >= 0) {
            a.part[(size_t)(prev_tb >> 5) * H + c] = tot;
          } else {
            float upd = prev_nvalid > 0 ? tot + sm.b2[c] : 0.f;
            a.h_bond_out[(size_t)prev_e * H + c] = a.h_bond_in[(size_t)prev_e * H + c] + upd;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();      // also orders this phase's attention weights before the next phase's reads within the CTA
  if (last && warp == 0) tmem_dealloc(tmem_base, 512);
}

template <bool VPASS>
__global__ void __launch_bounds__(TT_THREADS, 1) trip_tc_kernel(const TripArgs a) { trip_tc_body<VPASS>(a, true, true); }
__global__ void __launch_bounds__(TT_THREADS, 1) trip_tc_pair_kernel(const TripArgs a) {
  trip_tc_body<false>(a, true, false);
  trip_tc_body<true>(a, false, true);
}

// key + value pass in one launch: every (CTA, quadrant) walks the same groups in both phases (not for chunked groups, whose
// rescale factors need every CTA's key pass)
void launch_trip_tc_pair(const TripArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0 || a.n_groups <= 0) return;
  static DeviceOnce once;
  const int bytes = TripTcSmem::bytes();
  if (!once.done()) { cudaFuncSetAttribute(trip_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  const int grid = atc_grid((a.n_groups + 3) / 4, num_sms);
  launch_pdl(trip_tc_pair_kernel, dim3(grid), dim3(TT_THREADS), bytes, stream, a);
}

void launch_trip_tc(const TripArgs& a, bool vpass, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0 || a.n_groups <= 0) return;
  static DeviceOnce once;
  const int bytes = TripTcSmem::bytes();
  if (!once.done()) {
    cudaFuncSetAttribute(trip_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(trip_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    once.mark();
  }
  const int grid = atc_grid((a.n_groups + 3) / 4, num_sms);
  if (vpass) launch_pdl(trip_tc_kernel<true>, dim3(grid), dim3(TT_THREADS), bytes, stream, a);
  else launch_pdl(trip_tc_kernel<false>, dim3(grid), dim3(TT_THREADS), bytes, stream, a);
}

#ifdef DDB_TIMELINE
extern "C" int ddb_debug_timeline_trip(unsigned long long* out /* 2*2*16: {k, v} x {warp 0, warp 13} x phase */) {
  return (int)cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * 64);
}
#endif

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
