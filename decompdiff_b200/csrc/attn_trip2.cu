// Bond update over triplets k->j->i (BondUpdateLayer, uni_transformer_edge.py:125-167) with the second Linear COMMUTED through
// the attention (re-association 2'):
//
//   keys    logit[row, h] = <q_g[h], W2k[8h:8h+8, :] a_k[row]> = <U_g[h, :], a_k[row, :]>,   U_g[h, m] = sum_d q_g[8h+d] W2k[8h+d, m]
//   values  out_g[c]      = sum_rows w[row, h(c)] (W2v a_v[row] + b2)[c] = <W2v[c, :], S_g[h(c), :]> + b2[c],
//                                                                           S_g[h, m] = sum_rows w[row, h] a_v[row, m]
//
// i.e. 2 048 (+512 amortised) MACs per row and MLP instead of 16 384, all in fp32 FMA (packed FFMA2) - no TF32 split of the hidden
// activations, no 128 KB W2 image, no main accumulator in TMEM.  What stays on the tensor core is the one true shared-weight GEMM
// of the layer input: the angular term of the first Linear, D2 = Ang[128 x 16] * Wa^T (tcgen05, 3xTF32, one tile ahead).
//
//   rows    : 32 row slots (source atoms k of the bond edges entering j) x 4 groups (bond edges j->i) = one 128-row tile;
//             thread = (row, 32-channel slice); the 4 slice-warps of a quadrant sit on one SM sub-partition (TMEM lane rule), so
//             every exchange between them is a SPLIT-PHASE mbarrier (arrive ... independent work ... wait), never a blocking bar
//   P'      : the per-edge part of the first Linear (trip_prep) is stored in CSR order (rows of one source atom j contiguous) and
//             kept in TENSOR MEMORY: the thread's own row slice sits in its TMEM lane for as long as the quadrant stays on the same
//             source atom (~n_lig consecutive groups), double buffered; no shared memory and no re-gather per tile
//   k pass  : U_g per quadrant (warp-private slice), partial logits over the thread's 32 channels, one exchange between the slice
//             warps, softmax over the warp's 32 rows -> wbuf
//   v pass  : hidden activations and weights of a group go through shared memory once (swizzled), every thread accumulates a
//             4 head x 4 channel block of S_g over the rows, applies its part of W2v and the partial outputs are reduced with
//             shuffles + one exchange
#include "attn_tc.cuh"

namespace ddb {

namespace {

constexpr int T2_THREADS = 640;                 // 16 worker warps + one warpgroup whose first warp issues the MMAs
constexpr int T2_IMG = 2 * 128 * 32;            // one TF32 image of a [128 rows][16 cols] SWIZZLE_32B K-major operand (2 k-blocks)
constexpr int T2_COL_D2 = 0, T2_COL_P = 256;    // TMEM: D2 x2 (angular MMA, double buffered) | P' x2 (row slices of two source atoms)

// barrier slots
constexpr int B_W = 0, B_D2 = 1 /* +buf */, B_A2F = 3, B_LN = 4 /* +q */, B_X = 8 /* +q */, B_RED = 12 /* +q */, B_COUNT = 16;

__device__ __forceinline__ uint64_t t2_desc_sw32(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// byte offset of feature column k of row r inside one image
__device__ __forceinline__ int t2_off(int r, int k) { return (k >> 3) * 4096 + r * 32 + (((((k >> 2) & 1) ^ ((r >> 2) & 1))) << 4) + (k & 3) * 4; }

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// one arrival per warp: the warp's shared-memory writes are ordered before it by the warp barrier
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

template <bool VPASS>
struct T2Smem {
  uint8_t *A2, *B2; float *W2, *gamma, *beta, *b2, *qrow; float2* stat; uint64_t* bars; uint32_t* tmem_slot;
  float *ux, *qs;          // k pass: [16 warps][512] U slice / partial-logit exchange, [2][4][128] staged query rows (pair order)
  float *av, *wsm, *red;   // v pass: [4][32][128] hidden activations, [4][32][16] attention weights, [4][4][128] partial outputs
  __device__ explicit T2Smem(uint8_t* p) {
    A2 = p; p += 2 * T2_IMG;
    B2 = p; p += 2 * T2_IMG;
    W2 = reinterpret_cast<float*>(p); p += H * H * 4;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    qrow = reinterpret_cast<float*>(p); p += 16 * 64 * 4;
    stat = reinterpret_cast<float2*>(p); p += 2 * 128 * 4 * 8;
    bars = reinterpret_cast<uint64_t*>(p); p += B_COUNT * 8;
    tmem_slot = reinterpret_cast<uint32_t*>(p); p += 16;
    ux = qs = av = wsm = red = nullptr;
    if (!VPASS) {
      ux = reinterpret_cast<float*>(p); p += 16 * 512 * 4;
      qs = reinterpret_cast<float*>(p); p += 2 * 4 * 128 * 4;
    } else {
      av = reinterpret_cast<float*>(p); p += 4 * 32 * 128 * 4;
      wsm = reinterpret_cast<float*>(p); p += 4 * 32 * 16 * 4;
      red = reinterpret_cast<float*>(p); p += 4 * 4 * 128 * 4;
    }
  }
  static constexpr int bytes() {
    return 4 * T2_IMG + H * H * 4 + 3 * H * 4 + 16 * 64 * 4 + 2 * 128 * 4 * 8 + B_COUNT * 8 + 16 +
           (VPASS ? (4 * 32 * 128 + 4 * 32 * 16 + 4 * 4 * 128) * 4 : (16 * 512 + 2 * 4 * 128) * 4);
  }
};
static_assert(T2Smem<true>::bytes() <= 232448 && T2Smem<false>::bytes() <= 232448, "shared memory budget");

struct Grp {                    // one group = bond edge j->i in visiting order
  int e, node_i, node_j, base, pk;      // edge id (-1: none), merged node ids, first CSR row of source atom j, deg | excl << 8
  __device__ int deg() const { return pk & 0xff; }
  __device__ int excl() const { return (pk >> 8) & 0xff; }
};

template <bool VPASS>
__global__ void __launch_bounds__(T2_THREADS, 1) trip2_kernel(const TripArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  T2Smem<VPASS> sm(smem_raw);
  const TripSide& side = VPASS ? a.v : a.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = (warp >> 2) & 3, r = q * 32 + lane;
  auto bar = [&](int i) { return smem_u32(&sm.bars[i]); };
  if ((smem_u32(sm.A2) & 1023u) != 0u) __trap();
  if (tid == 0) {
    mbar_init(bar(B_W), 1); mbar_init(bar(B_D2), 1); mbar_init(bar(B_D2 + 1), 1); mbar_init(bar(B_A2F), 16);
    for (int i = 0; i < 4; ++i) { mbar_init(bar(B_LN + i), 4); mbar_init(bar(B_X + i), 4); mbar_init(bar(B_RED + i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(sm.tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_expect_tx(bar(B_W), H * H * 4 + 2 * T2_IMG);
    bulk_g2s(smem_u32(sm.W2), side.W2c, H * H * 4, bar(B_W));
    bulk_g2s(smem_u32(sm.B2), side.Wa32, 2 * T2_IMG, bar(B_W));
  }
  const uint32_t tmem_base = *sm.tmem_slot;
  cta_copy_f4(sm.gamma, side.w.gamma, H);
  cta_copy_f4(sm.beta, side.w.beta, H);
  cta_copy_f4(sm.b2, side.w.b2, H);
  // only 13 of the 16 feature columns are written per tile: clear both images once
  for (int i = tid * 16; i < 2 * T2_IMG; i += T2_THREADS * 16) *reinterpret_cast<float4*>(sm.A2 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  mbar_wait(bar(B_W), 0);

  // contiguous range of tiles (4 consecutive groups of the source-major order each) per CTA
  const int n_tiles = (a.n_bonds + 3) / 4;
  const int per = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = (int)blockIdx.x * per, nt = max(min(n_tiles, t0 + per) - t0, 0);

  if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 16 && lane == 0) {
      // ------------------------------------------------------------ MMA issuer: D2[t & 1] = A2(t) * Wa^T once every worker warp arrived
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a2 = smem_u32(sm.A2), b2 = smem_u32(sm.B2);
      for (int t = 0; t < nt; ++t) {
        mbar_wait(bar(B_A2F), t & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + T2_COL_D2 + (t & 1) * 128;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t o = (uint64_t)((kb * 4096) >> 4);
          umma_tf32_ss(d, t2_desc_sw32(a2) + o, t2_desc_sw32(b2) + o, idesc, kb ? 1u : 0u);
          umma_tf32_ss(d, t2_desc_sw32(a2 + T2_IMG) + o, t2_desc_sw32(b2) + o, idesc, 1u);
          umma_tf32_ss(d, t2_desc_sw32(a2) + o, t2_desc_sw32(b2 + T2_IMG) + o, idesc, 1u);
        }
        umma_commit(bar(B_D2 + (t & 1)));
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float* const wq = sm.qrow + warp * 64;                       // [parity][32] slice of the Q' row (warp-private)
    const float* __restrict__ Pc = side.Pcsr;
    const float* __restrict__ Qc = side.Q;

    auto load_grp = [&](int t) {
      Grp g; g.e = -1; g.node_i = 0; g.node_j = 0; g.base = -1; g.pk = 32 << 8;
      const int pos = (t0 + t) * 4 + q;
      if (t < nt && pos < a.n_bonds) {
        const int4 m = __ldg(a.grp4 + pos);
        g.e = m.x; g.node_i = m.y; g.node_j = m.z; g.base = m.w; g.pk = __ldg(a.grp_pk + pos);
      }
      return g;
    };
    // row slice of the P' rows of source atom `base` -> this thread's TMEM lane, buffer `buf`
    auto load_p = [&](int base, int buf) {
      const float* prow = Pc + (size_t)(base + lane) * H + s * 32;      // rows past the atom's degree are other atoms' rows (finite, masked)
      float4 v[8];
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) ldg8(prow + i8 * 8, v[2 * i8], v[2 * i8 + 1]);
      uint32_t u[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        u[4 * i] = __float_as_uint(v[i].x); u[4 * i + 1] = __float_as_uint(v[i].y);
        u[4 * i + 2] = __float_as_uint(v[i].z); u[4 * i + 3] = __float_as_uint(v[i].w);
      }
      tmem_st32(lane_addr + T2_COL_P + buf * 128 + s * 32, u);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    // angular features of this thread's row -> A2 (the 13 features are split over the 4 slice-warps), then one arrival per warp
    auto features = [&](const Grp& g, float4 xi, float4 xj, float4 xk) {
      const bool rowok = lane < g.deg() && lane != g.excl();
      const float ax = xj.x - xi.x, ay = xj.y - xi.y, az = xj.z - xi.z, bx = xk.x - xi.x, by = xk.y - xi.y, bz = xk.z - xi.z;
      const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
      float cn = sqrtf(cx * cx + cy * cy + cz * cz);             // |(j-i) x (k-i)|          (:134-137)
      float dot = ax * bx + ay * by + az * bz;
      if (!rowok) { cn = 0.f; dot = 1.f; }
      auto put = [&](int k, float v) {
        uint32_t hi, lo;
        tf32_split(v, hi, lo);
        *reinterpret_cast<uint32_t*>(sm.A2 + t2_off(r, k)) = hi;
        *reinterpret_cast<uint32_t*>(sm.A2 + T2_IMG + t2_off(r, k)) = lo;
      };
      // AngularEncoding [theta, sin(f theta), cos(f theta)], f = [1,2,3,1,1/2,1/3] (common.py:46-54): sin / cos of theta follow from
      // (cn, dot), multiples and the half angle from the usual identities; only theta and theta/3 need atan2f / sincosf
      if (s == 0) {
        put(0, atan2f(cn, dot));
      } else if (s == 3) {
        float sv, cv;
        sincosf(atan2f(cn, dot) * (float)(1.0 / 3.0), &sv, &cv);
        put(6, sv); put(12, cv);
      } else {
        const float n2 = cn * cn + dot * dot;
        const float inv = n2 > 0.f ? rsqrtf(n2) : 0.f;
        const float sn = cn * inv, cs = n2 > 0.f ? dot * inv : 1.f;
        if (s == 1) {
          put(1, sn); put(4, sn); put(7, cs); put(10, cs);
          put(2, 2.f * sn * cs); put(8, cs * cs - sn * sn);
          put(3, sn * (3.f - 4.f * sn * sn)); put(9, cs * (4.f * cs * cs - 3.f));
        } else {
          float sh, ch;
          if (cs >= 0.f) { ch = sqrtf(0.5f * (1.f + cs)); sh = sn / (2.f * ch); }
          else { sh = sqrtf(0.5f * (1.f - cs)); ch = sn / (2.f * sh); }
          put(5, sh); put(11, ch);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // A2 was written through the generic proxy
      tc_fence_before();
      warp_arrive(bar(B_A2F), lane);
    };

    // ---- pipeline state.  advance(T) computes z(T) + LayerNorm statistics of tile T, issues the angular features of tile T+1 and
    // loads the P' rows of tile T+1's source atom; it runs in the shadow of the exchanges of tile T-1 (see the loop).
    Grp gA = load_grp(0), gB = load_grp(1), gC = load_grp(2);      // tiles T, T+1, T+2 when advance(T) starts
    float2 z[16];
    int pbuf = 0;                                                   // TMEM buffer holding tile T's P' rows
    float4 xk = make_float4(0.f, 0.f, 0.f, 0.f), xj = xk, xi_n = xk;       // geometry of tile T+1 (features run one tile ahead)
    float q_nx = 0.f, qry_nx = 0.f;                                 // Q' / query slices of tile T+1 (requested one advance earlier)
    float4 w4_nx = make_float4(0.f, 0.f, 0.f, 0.f);                // v pass: attention weights of this thread's row, tile T
    auto src_geometry = [&](const Grp& g) {                         // positions of the row's atom k and of j (per source atom)
      xk = ldg4(a.xcsr + (size_t)(max(g.base, 0) + lane) * 4);
      xj = ldg4(a.x4 + (size_t)g.node_j * 4);
    };
    auto advance = [&](int T) {
      // staged per-group rows of tile T (requested one advance ago)
      wq[(T & 1) * 32 + lane] = q_nx;
      if (!VPASS) {
        const int c = s * 32 + lane, hh = c >> 3, d = c & 7;
        sm.qs[((T & 1) * 4 + q) * 128 + (hh >> 1) * 16 + d * 2 + (hh & 1)] = qry_nx;      // pair order: (head 2p, head 2p+1) adjacent
      }
      // requests for tile T+1
      q_nx = __ldg(Qc + (size_t)max(gB.e, 0) * H + s * 32 + lane);
      if (!VPASS) qry_nx = __ldg(a.q + (size_t)max(gB.e, 0) * a.ldq + s * 32 + lane);
      if (VPASS) w4_nx = ld4(a.wbuf + ((size_t)((t0 + T) * 4 + q) * 32 + lane) * NH + s * 4);
      const Grp gD = load_grp(T + 3);
      __syncwarp();
      // ---- z(T) = D2 (angular MMA) + P'[kj] (TMEM) + Q'[ji] (staged), 16 channels at a time
      mbar_wait(bar(B_D2 + (T & 1)), (T >> 1) & 1);
      tc_fence_after();
      const float* qs_ = wq + (T & 1) * 32;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t d2[16], pp[16];
        tmem_ld16(lane_addr + T2_COL_D2 + (T & 1) * 128 + s * 32 + half * 16, d2);
        tmem_ld16(lane_addr + T2_COL_P + pbuf * 128 + s * 32 + half * 16, pp);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const float4 qv = ld4(qs_ + half * 16 + i4 * 4);
          float2 u0 = __fadd2_rn(f2(__uint_as_float(d2[4 * i4]), __uint_as_float(d2[4 * i4 + 1])), f2(__uint_as_float(pp[4 * i4]), __uint_as_float(pp[4 * i4 + 1])));
          float2 u1 = __fadd2_rn(f2(__uint_as_float(d2[4 * i4 + 2]), __uint_as_float(d2[4 * i4 + 3])), f2(__uint_as_float(pp[4 * i4 + 2]), __uint_as_float(pp[4 * i4 + 3])));
          z[half * 8 + i4 * 2] = __fadd2_rn(u0, f2(qv.x, qv.y));
          z[half * 8 + i4 * 2 + 1] = __fadd2_rn(u1, f2(qv.z, qv.w));
        }
      }
      // ---- LayerNorm statistics: {sum, sum of squares about the slice mean}, combined after the exchange
      {
        float2 s1 = f2(0.f, 0.f), s2 = f2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) s1 = __fadd2_rn(s1, z[i]);
        const float part = s1.x + s1.y, mu_s = part * (1.0f / 32.0f);
        const float2 nm = f2(-mu_s, -mu_s);
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float2 dz = __fadd2_rn(z[i], nm); s2 = __ffma2_rn(dz, dz, s2); }
        sm.stat[((T & 1) * 128 + r) * 4 + s] = make_float2(part, s2.x + s2.y);
      }
      // ---- tile T+1: angular features (its D2 is ready when advance(T+1) runs), P' rows of a new source atom
      if (T + 1 < nt) {
        tc_fence_before();                    // D2 / P' reads above are complete (wait::ld); order them before the arrival
        features(gB, xi_n, xj, xk);
      }
      // the quadrant moves to another source atom with tile T+1: its rows go to the other TMEM buffer (z(T) above is done with
      // this one only after the flip below); then the geometry for tile T+2
      const bool flip = (T + 1 < nt) && gB.base != gA.base && gB.e >= 0;
      if (flip) load_p(gB.base, pbuf ^ 1);
      if (T + 2 < nt && gC.e >= 0) {
        if (gC.base != gB.base || gB.e < 0) src_geometry(gC);
        xi_n = ldg4(a.x4 + (size_t)gC.node_i * 4);
      }
      if (flip) pbuf ^= 1;
      gA = gB; gB = gC; gC = gD;
    };
    // LayerNorm + ReLU of z (tile T) in place
    auto normalise = [&](int T) {
      const float2* st = sm.stat + ((T & 1) * 128 + r) * 4;
      const float4 t01 = *reinterpret_cast<const float4*>(st), t23 = *reinterpret_cast<const float4*>(st + 2);
      const float mu = ((t01.x + t01.z) + (t23.x + t23.z)) * (1.0f / H);
      const float d0 = t01.x * (1.0f / 32.0f) - mu, d1 = t01.z * (1.0f / 32.0f) - mu, d2 = t23.x * (1.0f / 32.0f) - mu, d3 = t23.z * (1.0f / 32.0f) - mu;
      const float m2 = ((t01.y + t01.w) + (t23.y + t23.w)) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
      const float rstd = rsqrtf(m2 * (1.0f / H) + LN_EPS);
      const float2 rs2 = f2(rstd, rstd), nm2 = f2(-mu * rstd, -mu * rstd);
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        const float4 gm = ld4(sm.gamma + s * 32 + i4 * 4), bt = ld4(sm.beta + s * 32 + i4 * 4);
        float2 u0 = __ffma2_rn(z[i4 * 2], rs2, nm2), u1 = __ffma2_rn(z[i4 * 2 + 1], rs2, nm2);
        u0 = __ffma2_rn(u0, f2(gm.x, gm.y), f2(bt.x, bt.y));
        u1 = __ffma2_rn(u1, f2(gm.z, gm.w), f2(bt.z, bt.w));
        z[i4 * 2] = f2(fmaxf(u0.x, 0.f), fmaxf(u0.y, 0.f));
        z[i4 * 2 + 1] = f2(fmaxf(u1.x, 0.f), fmaxf(u1.y, 0.f));
      }
    };

    if (nt > 0) {
      // ---- prologue: rows and geometry of tile 0, its features, then advance(0) (which also prepares tile 1)
      if (gA.e >= 0) {
        load_p(gA.base, 0);
        src_geometry(gA);
        xi_n = ldg4(a.x4 + (size_t)gA.node_i * 4);
        q_nx = __ldg(Qc + (size_t)gA.e * H + s * 32 + lane);
        if (!VPASS) qry_nx = __ldg(a.q + (size_t)gA.e * a.ldq + s * 32 + lane);
      }
      features(gA, xi_n, xj, xk);
      if (nt > 1 && gB.e >= 0) {               // geometry of tile 1 for the features issued inside advance(0)
        if (gB.base != gA.base || gA.e < 0) src_geometry(gB);
        xi_n = ldg4(a.x4 + (size_t)gB.node_i * 4);
      }
      float4 w4_cur = make_float4(0.f, 0.f, 0.f, 0.f);
      // advance() shifts the group registers; keep what the epilogue of a tile needs
      int cur_e = gA.e, cur_pk = gA.pk;
      advance(0);
      if (VPASS) w4_cur = w4_nx;
      warp_arrive(bar(B_LN + q), lane);
      int prev_e = -1, prev_nvalid = 0; float hb_prev = 0.f;

      for (int it = 0; it < nt; ++it) {
        const bool rowok = cur_e >= 0 && lane < (cur_pk & 0xff) && lane != ((cur_pk >> 8) & 0xff);
        const int nvalid = __popc(__ballot_sync(FULL, rowok));
        const int pos = (t0 + it) * 4 + q;
        // ---- 1. hidden activations of tile `it`
        mbar_wait(bar(B_LN + q), it & 1);
        normalise(it);
        if (!VPASS) {
          // ---- k pass.  U slice of this warp: U[h][m], m = 32 s + lane, from the staged query row and W2k (pair layout)
          float* const uw = sm.ux + warp * 512;
          {
            const float* qp = sm.qs + ((it & 1) * 4 + q) * 128;
            const float2* wp = reinterpret_cast<const float2*>(sm.W2) + s * 32 + lane;      // [(hp*8 + d)][128 m] float2
#pragma unroll
            for (int hp = 0; hp < 8; ++hp) {
              float2 acc = f2(0.f, 0.f);
#pragma unroll
              for (int d2 = 0; d2 < 4; ++d2) {
                const float4 qq = ld4(qp + hp * 16 + d2 * 4);
                acc = __ffma2_rn(f2(qq.x, qq.y), wp[(hp * 8 + d2 * 2) * 128], acc);
                acc = __ffma2_rn(f2(qq.z, qq.w), wp[(hp * 8 + d2 * 2 + 1) * 128], acc);
              }
              uw[(2 * hp) * 32 + lane] = acc.x;
              uw[(2 * hp + 1) * 32 + lane] = acc.y;
            }
          }
          __syncwarp();
          // partial logits over this thread's 32 channels, all 16 heads
          float part[16];
#pragma unroll
          for (int h = 0; h < 16; ++h) {
            float2 acc = f2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 u = ld4(uw + h * 32 + i * 4);
              acc = __ffma2_rn(z[2 * i], f2(u.x, u.y), acc);
              acc = __ffma2_rn(z[2 * i + 1], f2(u.z, u.w), acc);
            }
            part[h] = acc.x + acc.y;
          }
          __syncwarp();                        // every lane is done with U: the region becomes this warp's exchange rows
          {
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int j = 0; j < 4; ++j) st4(uw + lane * 16 + ((j ^ sw) << 2), make_float4(part[4 * j], part[4 * j + 1], part[4 * j + 2], part[4 * j + 3]));
          }
          warp_arrive(bar(B_X + q), lane);
          // ---- 2. in the shadow of the exchange: z / statistics of the next tile, features of the one after
          cur_e = gA.e; cur_pk = gA.pk;
          if (it + 1 < nt) advance(it + 1);
          // ---- 3. logits of heads 4s..4s+3, softmax over the 32 rows of the group
          mbar_wait(bar(B_X + q), it & 1);
          float lg[4] = {0.f, 0.f, 0.f, 0.f};
          {
            const int sw = (lane >> 1) & 3;
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
              const float4 p4 = ld4(sm.ux + (s2 * 4 + q) * 512 + lane * 16 + ((s ^ sw) << 2));
              lg[0] += p4.x; lg[1] += p4.y; lg[2] += p4.z; lg[3] += p4.w;
            }
          }
          if (it + 1 < nt) warp_arrive(bar(B_LN + q), lane);      // after the exchange rows were read: the next U may overwrite them
          float ex[4];
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            const float v = rowok ? lg[hh] : -INFINITY;
            float m;
            asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
            ex[hh] = rowok ? __expf(v - m) : 0.f;
          }
          float sum[4] = {ex[0], ex[1], ex[2], ex[3]};
          warp_allreduce4(sum, lane);
          float w[4];
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) w[hh] = sum[hh] > 0.f ? __fdividef(ex[hh], sum[hh]) : 0.f;
          if (pos < a.n_bonds) st4(a.wbuf + ((size_t)pos * 32 + lane) * NH + s * 4, make_float4(w[0], w[1], w[2], w[3]));
          (void)nvalid; (void)prev_e; (void)prev_nvalid; (void)hb_prev; (void)w4_cur;
        } else {
          // ---- v pass.  hidden activations and attention weights of the group -> shared memory (swizzled 16-byte chunks)
          {
            float* const arow = sm.av + (q * 32 + lane) * 128;
            const int sw = lane & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) st4(arow + ((s * 8 + (i ^ sw)) << 2), make_float4(z[2 * i].x, z[2 * i].y, z[2 * i + 1].x, z[2 * i + 1].y));
            st4(sm.wsm + (q * 32 + lane) * 16 + ((s ^ ((lane >> 1) & 3)) << 2), rowok ? w4_cur : make_float4(0.f, 0.f, 0.f, 0.f));
          }
          // finish the previous tile: sum the partial outputs of the 4 slice-warps, bias, residual (:274)
          if (it > 0) {
            mbar_wait(bar(B_RED + q), (it - 1) & 1);
            const int c = s * 32 + lane;
            const float* rp = sm.red + q * 512 + c;
            const float o = (rp[0] + rp[128]) + (rp[256] + rp[384]);
            if (prev_e >= 0) a.h_bond_out[(size_t)prev_e * H + c] = hb_prev + (prev_nvalid > 0 ? o + sm.b2[c] : 0.f);
          }
          warp_arrive(bar(B_X + q), lane);
          const int deg_cur = cur_e >= 0 ? (cur_pk & 0xff) : 0;
          prev_e = cur_e; prev_nvalid = nvalid;
          if (cur_e >= 0) hb_prev = __ldg(a.h_bond_in + (size_t)cur_e * H + s * 32 + lane);
          cur_e = gA.e; cur_pk = gA.pk;
          // ---- 2. next tile's z / statistics, features of the one after
          if (it + 1 < nt) { advance(it + 1); w4_cur = w4_nx; }
          // ---- 3. S[4 heads][4 channels] of this thread over the rows of the group
          mbar_wait(bar(B_X + q), it & 1);
          const int t128 = s * 32 + lane, hq = t128 & 3, mq = t128 >> 2;
          float2 S[4][2];
#pragma unroll
          for (int i = 0; i < 4; ++i) { S[i][0] = f2(0.f, 0.f); S[i][1] = f2(0.f, 0.f); }
          {
            const float* ab = sm.av + q * 32 * 128;
            const float* wb = sm.wsm + q * 32 * 16;
#pragma unroll 4
            for (int k = 0; k < deg_cur; ++k) {
              const float4 wv = ld4(wb + k * 16 + ((hq ^ ((k >> 1) & 3)) << 2));
              const float4 av4 = ld4(ab + k * 128 + (((mq & ~7) | ((mq ^ k) & 7)) << 2));
              const float2 a01 = f2(av4.x, av4.y), a23 = f2(av4.z, av4.w);
              S[0][0] = __ffma2_rn(f2(wv.x, wv.x), a01, S[0][0]); S[0][1] = __ffma2_rn(f2(wv.x, wv.x), a23, S[0][1]);
              S[1][0] = __ffma2_rn(f2(wv.y, wv.y), a01, S[1][0]); S[1][1] = __ffma2_rn(f2(wv.y, wv.y), a23, S[1][1]);
              S[2][0] = __ffma2_rn(f2(wv.z, wv.z), a01, S[2][0]); S[2][1] = __ffma2_rn(f2(wv.z, wv.z), a23, S[2][1]);
              S[3][0] = __ffma2_rn(f2(wv.w, wv.w), a01, S[3][0]); S[3][1] = __ffma2_rn(f2(wv.w, wv.w), a23, S[3][1]);
            }
          }
          if (it + 1 < nt) warp_arrive(bar(B_LN + q), lane);      // the group's rows in shared memory are free again
          // ---- 4. this thread's part of W2v: channels of heads 4hq..4hq+3 (32 outputs) over its 4 hidden channels; the 8 lanes
          // that share hq are reduced with a butterfly reduce-scatter, the 4 slice-warps through shared memory
          float val[32];
          {
            const float* wrow = sm.W2 + (size_t)(hq * 32) * H + mq * 4;
#pragma unroll
            for (int cc = 0; cc < 32; ++cc) {
              const float4 w4 = ld4(wrow + cc * H);
              const float2 p = __ffma2_rn(f2(w4.z, w4.w), S[cc >> 3][1], __fmul2_rn(f2(w4.x, w4.y), S[cc >> 3][0]));
              val[cc] = p.x + p.y;
            }
          }
#pragma unroll
          for (int m = 16, n = 32; m >= 4; m >>= 1, n >>= 1) {
            const bool upper = lane & m;
            const int half = n >> 1;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i < half) {
                const float keep = upper ? val[i + half] : val[i];
                const float send = upper ? val[i] : val[i + half];
                val[i] = keep + __shfl_xor_sync(FULL, send, m);
              }
            }
          }
          {
            const int c0 = hq * 32 + ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4;
            st4(sm.red + q * 512 + s * 128 + c0, make_float4(val[0], val[1], val[2], val[3]));
          }
          warp_arrive(bar(B_RED + q), lane);
        }
      }
      if (VPASS) {      // epilogue of the last tile
        mbar_wait(bar(B_RED + q), (nt - 1) & 1);
        const int c = s * 32 + lane;
        const float* rp = sm.red + q * 512 + c;
        const float o = (rp[0] + rp[128]) + (rp[256] + rp[384]);
        if (prev_e >= 0) a.h_bond_out[(size_t)prev_e * H + c] = hb_prev + (prev_nvalid > 0 ? o + sm.b2[c] : 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace

void launch_trip2(const TripArgs& a, bool vpass, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0) return;
  static DeviceOnce once;
  if (!once.done()) {
    cudaFuncSetAttribute(trip2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Smem<false>::bytes());
    cudaFuncSetAttribute(trip2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2Smem<true>::bytes());
    once.mark();
  }
  const int grid = atc_grid((a.n_bonds + 3) / 4, num_sms);
  if (vpass) trip2_kernel<true><<<grid, T2_THREADS, T2Smem<true>::bytes(), stream>>>(a);
  else trip2_kernel<false><<<grid, T2_THREADS, T2Smem<false>::bytes(), stream>>>(a);
}

// host-side packing of Wa[13][128] (first-Linear columns of the angular encoding, transposed) into the B operand of the angular
// MMA: rows n = output channel, K = 16 features (13 used), hi | lo images in the SWIZZLE_32B K-major layout of the kernel
void pack_wa_sw32(const float* Wa, float* out /* 2 * T2_IMG / 4 floats */) {
  for (int i = 0; i < 2 * T2_IMG / 4; ++i) out[i] = 0.f;
  float* hi = out;
  float* lo = out + T2_IMG / 4;
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < NANG; ++k) {
      const float w = Wa[k * H + n];
      const float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
      const int off = ((k >> 3) * 4096 + n * 32 + ((((k >> 2) & 1) ^ ((n >> 2) & 1)) << 4) + (k & 3) * 4) / 4;
      hi[off] = h; lo[off] = l;
    }
}

// key second Linear W2k[128 out][128 in] (already scaled by 1/sqrt(8)) in the pair layout of the U computation:
// out[((hp * 8 + d) * 128 + m) * 2 + b] = W2k[8 (2 hp + b) + d][m]
void pack_w2k_pairs(const float* W2, float* out /* 128*128 floats */) {
  for (int hp = 0; hp < 8; ++hp)
    for (int d = 0; d < 8; ++d)
      for (int m = 0; m < 128; ++m)
        for (int b = 0; b < 2; ++b) out[((size_t)(hp * 8 + d) * 128 + m) * 2 + b] = W2[(size_t)(8 * (2 * hp + b) + d) * 128 + m];
}

}  // namespace ddb
