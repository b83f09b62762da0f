// Tensor-core bond update over triplets k->j->i (BondUpdateLayer, uni_transformer_edge.py:125-167), third generation: the same
// two tcgen05 GEMMs per 128-row tile as attn_tc_trip.cu (angular term of the first Linear, second Linear on the hidden
// activations), but NO worker thread keeps gathered rows or prefetched scalars in registers:
//   * the rows P'[k->j] of a source atom j (one "unit" = up to 32 rows, shared by every group j->i of that source) are staged in
//     shared memory by the bulk-copy engine (cp.async.bulk, one 512-byte copy per row into 528-byte padded rows, completion
//     counted on an mbarrier); two unit buffers alternate, a tile touches at most two consecutive units (the host pads the
//     visiting order so that this holds for any ligand size)
//   * per-tile scalars {edge id, partner chunk, valid-row mask, unit ordinal, first CSR row} and the per-group operands (Q' slice,
//     query slice; attention weights, chunk factor and residual row in the value pass) reach per-warp rings with cp.async
//     (LDGSTS), issued one to two tiles ahead
// The round-2 profile of attn_tc_trip.cu showed a third of the worker time in long-scoreboard stalls on spill stores of
// just-loaded prefetch values and on spill reloads (L1 is ~28 KB next to 225 KB of shared memory); here the register file holds
// only the row slice being transformed.  A tile = 4 consecutive positions of the visiting order (quadrant q <-> position 4t + q),
// a CTA walks a contiguous range of tiles.
#include "attn_tc.cuh"

namespace ddb {

constexpr int T3_IMG = 128 * 64;                 // one TF32 image of a [128 rows][16 features] operand: 64-byte rows, SWIZZLE_64B
constexpr int T3_COL_D2 = 384;
constexpr int T3_PROW = 132;                     // floats per staged P' row (528 B): thread = row reads are bank-conflict free
constexpr int T3_PBUF = 32 * T3_PROW;            // floats per unit buffer
constexpr int T3_THREADS = ATC_THREADS + 128;
constexpr int T3_ISSUER = 16;
constexpr int T3_SYNC = ATC_THREADS + 32;
constexpr int T3_BAR_A_READY = 6, T3_BAR_WORKERS = 7;
#ifdef T3_SPIN
#define T3_WAIT mbar_wait_spin
#else
#define T3_WAIT mbar_wait
#endif
// per-warp ring (bytes): scalars 4 stages x 128 | Q' slice 2 x 128 | k: query slice 2 x 128 / v: weights 512 + factor 64 + residual 128
constexpr int T3_RING_SCAL = 0, T3_RING_Q = 512, T3_RING_X = 768;
constexpr int T3_RING_K = 1024, T3_RING_V = 1472;

template <bool VPASS>
struct Trip3Smem {
  uint8_t *W2, *B2, *A2, *ring; float *P, *gamma, *beta, *b2; float2* stat; uint64_t* bars; uint32_t* tmem_slot;
  static constexpr int RING = VPASS ? T3_RING_V : T3_RING_K;
  __device__ explicit Trip3Smem(uint8_t* raw) {
    uint8_t* p = raw;
    W2 = p; p += ATC_W2_BYTES;
    B2 = p; p += 2 * T3_IMG;
    A2 = p; p += 2 * T3_IMG;
    P = reinterpret_cast<float*>(p); p += 2 * T3_PBUF * 4;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    stat = reinterpret_cast<float2*>(p); p += 2 * 128 * 4 * 8;      // [parity][slice][row] {sum, sum of squares}
    ring = p; p += 16 * T3_RING_V;                                  // both passes carve the larger ring (one layout for the pair kernel)
    bars = reinterpret_cast<uint64_t*>(p); p += 128;       // two sets of 8: the second phase of a paired launch uses its own
    tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  static constexpr int bytes() { return ATC_W2_BYTES + 4 * T3_IMG + 2 * T3_PBUF * 4 + 3 * H * 4 + 2 * 128 * 4 * 8 + 16 * T3_RING_V + 128 + 32; }
};
static_assert(Trip3Smem<true>::bytes() <= 232448, "shared memory budget");

// K-major SWIZZLE_64B shared-memory matrix descriptor: rows of 64 bytes (16 tf32), 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// four consecutive features (chunk c = k / 4) of row r -> both images (16-byte chunk index XOR bits 1..2 of the row: Swizzle<2,4,3>)
__device__ __forceinline__ void t3_put4(uint8_t* A2, int r, int c, float v0, float v1, float v2, float v3) {
  uint4 hi, lo;
  tf32_split(v0, hi.x, lo.x); tf32_split(v1, hi.y, lo.y); tf32_split(v2, hi.z, lo.z); tf32_split(v3, hi.w, lo.w);
  const int off = r * 64 + ((c ^ ((r >> 1) & 3)) << 4);
  *reinterpret_cast<uint4*>(A2 + off) = hi;
  *reinterpret_cast<uint4*>(A2 + T3_IMG + off) = lo;
}

__device__ __forceinline__ float2 t3f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 t3u2f(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }
__device__ __forceinline__ void t3_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void t3_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void cpa4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <bool VPASS>
__device__ __forceinline__ void trip3_body(const TripArgs& a, const bool first, const bool last) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Trip3Smem<VPASS> sm(smem_raw);
  uint64_t* const bars = sm.bars + (first ? 0 : 8);      // a fresh barrier set per phase (no re-initialisation of used barriers)
  const TripSide& side = VPASS ? a.v : a.k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = (warp >> 2) & 3, r = q * 32 + lane;
  if ((smem_u32(sm.W2) & 1023u) != 0u) __trap();
  if (tid == 0) {
    // [0] weights landed, [1] main MMA retired, [2] angular MMA retired, [3] angular features of a tile written (3 producer warps),
    // [4] D2 of a tile read by every worker warp, [5] / [6] unit buffer 0 / 1 landed
    for (int i = 0; i < 7; ++i) mbar_init(smem_u32(&bars[i]), i == 3 ? 3 : i == 4 ? 16 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (first && warp == 0) { __syncwarp(); tmem_alloc(smem_u32(sm.tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, ATC_W2_BYTES + 2 * T3_IMG);
    bulk_g2s(smem_u32(sm.W2), side.W2tc, ATC_W2_BYTES / 2, bar);
    bulk_g2s(smem_u32(sm.W2) + ATC_W2_BYTES / 2, side.W2tc + ATC_W2_BYTES / 8, ATC_W2_BYTES / 2, bar);
    bulk_g2s(smem_u32(sm.B2), side.Wa64, 2 * T3_IMG, bar);
  }
  const uint32_t tmem_base = *sm.tmem_slot;
  cta_copy_f4(sm.gamma, side.w.gamma, H);
  cta_copy_f4(sm.beta, side.w.beta, H);
  cta_copy_f4(sm.b2, side.w.b2, H);
  // features 13..15 of every row stay zero; the unit buffers and rings start from defined values (rows of padding groups are
  // computed and dropped - they must not hold NaN patterns that would only cost denormal / exception paths)
  for (int i = tid * 16; i < 2 * T3_IMG; i += T3_THREADS * 16) *reinterpret_cast<float4*>(sm.A2 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 2 * T3_PBUF; i += T3_THREADS) sm.P[i] = 0.f;
  for (int i = tid * 4; i < 16 * T3_RING_V; i += T3_THREADS * 4) *reinterpret_cast<uint32_t*>(sm.ring + i) = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the unit buffers are written by the bulk-copy engine later
  if (first) pdl_wait();      // set-up on static data above; the previous kernels' results are visible below
  __syncthreads();
  mbar_wait(smem_u32(&bars[0]), 0);
  const uint32_t bar_mma = smem_u32(&bars[1]), bar_ang = smem_u32(&bars[2]), bar_a2f = smem_u32(&bars[3]), bar_d2c = smem_u32(&bars[4]);
  const uint32_t w2_smem = smem_u32(sm.W2), a2_smem = smem_u32(sm.A2), b2_smem = smem_u32(sm.B2);
  // a CTA walks the contiguous tile range [t0, t0 + cnt); every role derives the same count
  const int per = (a.n_tiles3 + (int)gridDim.x - 1) / (int)gridDim.x;
  const int t0 = (int)blockIdx.x * per, cnt = max(0, min(a.n_tiles3, t0 + per) - t0);

  if (warp >= 16) {
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
#endif
    if (warp == T3_ISSUER) {
      // tensor-pipe order: ang(t0), [ang(t1), main(t0)], [ang(t2), main(t1)], ...  - the angular MMA runs one tile ahead
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      auto issue_ang = [&](int t) {
        if (lane == 0) {
          mbar_wait(bar_a2f, t & 1);                   // the producers have written the features of tile t
          if (t > 0) mbar_wait(bar_d2c, (t - 1) & 1);   // every worker warp has read D2 of tile t-1
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {             // K = 16 features: two k-steps of 8 inside the 64-byte rows
            umma_tf32_ss(tmem_base + T3_COL_D2, umma_desc_sw64(a2_smem + ks * 32), umma_desc_sw64(b2_smem + ks * 32), idesc, ks ? 1u : 0u);
            umma_tf32_ss(tmem_base + T3_COL_D2, umma_desc_sw64(a2_smem + T3_IMG + ks * 32), umma_desc_sw64(b2_smem + ks * 32), idesc, 1u);
            umma_tf32_ss(tmem_base + T3_COL_D2, umma_desc_sw64(a2_smem + ks * 32), umma_desc_sw64(b2_smem + T3_IMG + ks * 32), idesc, 1u);
          }
          umma_commit(bar_ang);
        }
        __syncwarp();
      };
      if (cnt > 0) issue_ang(0);
      for (int it = 0; it < cnt; ++it) {
        if (it + 1 < cnt) issue_ang(it + 1);
        t3_sync(T3_BAR_A_READY, T3_SYNC);           // hidden activations are in TMEM, D of the previous tile is in registers
        if (lane == 0) { tc_fence_after(); atc_issue_mma(tmem_base, w2_smem, bar_mma); }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- producers: geometry of a row -> its 13 angular features -> A2.
      // Thread = row; warp 17 serves quadrants 0 and 3, warps 18 / 19 quadrants 1 / 2.
      const int pw = warp - 17;
      auto row_meta_of = [&](int qq, int t, int2& gm, int2& rm) {
        gm = make_int2(0, 0); rm = make_int2(-1, -1);
        if (t < cnt) {
          const int pos = (t0 + t) * 4 + qq;
          gm = __ldg(a.grp_meta + pos); rm = __ldg(a.row_meta + (size_t)pos * 32 + lane);
        }
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
        t3_put4(sm.A2, rr, 0, theta, sn, 2.f * sn * cs, sn * (3.f - 4.f * sn * sn));
        t3_put4(sm.A2, rr, 1, sn, sh, s3, cs);
        t3_put4(sm.A2, rr, 2, cs * cs - sn * sn, cs * (4.f * cs * cs - 3.f), cs, ch);
        t3_put4(sm.A2, rr, 3, c3, 0.f, 0.f, 0.f);
      };
      const int q0 = pw == 0 ? 0 : pw, q1 = 3;          // warp 17 also serves quadrant 3
      int2 gm0, rm0, gm1, rm1;
      row_meta_of(q0, 0, gm0, rm0);
      if (pw == 0) row_meta_of(q1, 0, gm1, rm1);
      for (int t = 0; t < cnt; ++t) {
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
        put_features(q0, rm0c, xi0, xj0, xk0);
        if (pw == 0) put_features(q1, rm1c, xi1, xj1, xk1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // A2 was written through the generic proxy
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a2f) : "memory");
      }
    }
  } else {
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
#endif
    // ---------------------------------------------------------------- 16 warps: thread = (row r, channel slice s)
    uint8_t* const ring = sm.ring + warp * T3_RING_V;
    const uint32_t ring_u = smem_u32(ring);
    const int4* __restrict__ recs = a.tile_rec;      // 8 int4 per tile: position p of the tile at [2p] {e, pair, valid mask, unit ordinal}, [2p+1] {first CSR row, ...}
    const float* __restrict__ Pcsr = side.Pcsr;     // centred rows in CSR order (trip_prep): the rows of one unit are contiguous
    const float* __restrict__ Qc = side.Q;
    auto rec0 = [&](int it, int p) { return *reinterpret_cast<const int4*>(ring + T3_RING_SCAL + (it & 3) * 128 + p * 32); };
    auto rec1x = [&](int it, int p) { return *reinterpret_cast<const int*>(ring + T3_RING_SCAL + (it & 3) * 128 + p * 32 + 16); };
    auto fetch_rec = [&](int it) {      // tile record of iteration `it` -> ring stage it & 3 (8 lanes x 16 bytes)
      if (lane < 8 && it < cnt) cpa16(ring_u + T3_RING_SCAL + (it & 3) * 128 + lane * 16, recs + (size_t)(t0 + it) * 8 + lane);
    };
    // bulk copies of the units first used by tile `it` (ordinals above `ul_before`): issued by worker warp 0, one row per lane
    auto load_units = [&](int it, int ul_before) {
      const int uf = rec0(it, 0).w, ul = rec0(it, 3).w;
      for (int u = max(uf, ul_before + 1); u <= ul; ++u) {
        const int row0 = u == uf ? rec1x(it, 0) : rec1x(it, 3);
        const uint32_t bar = smem_u32(&bars[5 + (u & 1)]);
        if (lane == 0) mbar_expect_tx(bar, 32 * H * 4);
        __syncwarp();
        bulk_g2s(smem_u32(sm.P + (u & 1) * T3_PBUF + lane * T3_PROW), Pcsr + ((size_t)row0 + lane) * H, H * 4, bar);
      }
    };
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
    // ---- prologue: records of the first two tiles, the first units, the Q' slice of the first group
    int ul_prev = -1;
    uint32_t unit_phase = 0;      // bit b: parity of the next completion of unit buffer b
    if (cnt > 0) {
      fetch_rec(0); fetch_rec(1);
      cpa_commit(); cpa_wait_all();
      __syncwarp();
      ul_prev = rec0(0, 0).w - 1;
      if (warp == 0) load_units(0, ul_prev);
      const int e0 = rec0(0, q).x;
      if (e0 >= 0) cpa4(ring_u + T3_RING_Q + lane * 4, Qc + (size_t)e0 * H + s * 32 + lane);
      cpa_commit();
    }
    int it = 0;
    for (; it < cnt; ++it) {
      cpa_wait_all();      // everything requested one iteration ago has landed
      __syncwarp();
      const int4 rc = rec0(it, q);
      const int e = rc.x;
      const bool rowok = (rc.z >> lane) & 1;          // valid and k != i (:117-118)
      const int tb = ((t0 + it) * 4 + q) * 32;        // wbuf rows of a (group, chunk) are its 32 slots in visiting order
      const int uf = rec0(it, 0).w, ul = rec0(it, 3).w;
      // ---- requests for later: the record two tiles ahead, the Q' slice of the next group, this group's query slice (k) /
      // the previous group's attention weights, chunk factor and residual row (v)
      fetch_rec(it + 2);
      if (it + 1 < cnt) {
        const int e_n = rec0(it + 1, q).x;
        if (e_n >= 0) cpa4(ring_u + T3_RING_Q + ((it + 1) & 1) * 128 + lane * 4, Qc + (size_t)e_n * H + s * 32 + lane);
      }
      if (!VPASS) {
        if (e >= 0) cpa4(ring_u + T3_RING_X + (it & 1) * 128 + lane * 4, a.q + (size_t)e * a.ldq + s * 32 + lane);
      } else if (it > 0) {
        const int4 rp = rec0(it - 1, q);
        const int ptb = tb - 4 * 32;      // same quadrant, previous tile
        if ((rp.z >> lane) & 1) cpa16(ring_u + T3_RING_X + lane * 16, a.wbuf + ((size_t)ptb + lane) * NH + s * 4);
        if (rp.y >= 0 && lane == 0) cpa16(ring_u + T3_RING_X + 512, a.factor + (size_t)(ptb >> 5) * NH + s * 4);
        if (rp.x >= 0) cpa4(ring_u + T3_RING_X + 576 + lane * 4, a.h_bond_in + (size_t)rp.x * H + s * 32 + lane);
      }
      cpa_commit();
      // ---- units first used by this tile were requested one tile ago (prologue for the first tile): wait for them
      for (int u = max(uf, ul_prev + 1); u <= ul; ++u) {
        mbar_wait(smem_u32(&bars[5 + (u & 1)]), (unit_phase >> (u & 1)) & 1u);
        unit_phase ^= 1u << (u & 1);
      }
      // ---- first Linear: z = P'[kj] (unit buffer) + Q'[ji] (ring) + D2 (angular MMA, issued one iteration ago)
      float2 z[16];
      {
        T3_WAIT(bar_ang, it & 1);
        tc_fence_after();
        __syncwarp();
        const float* qs = reinterpret_cast<const float*>(ring + T3_RING_Q + (it & 1) * 128);
        const float* ps = sm.P + (rc.w & 1) * T3_PBUF + lane * T3_PROW + s * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + T3_COL_D2 + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 qv = ld4(qs + i4 * 4), pv = ld4(ps + i4 * 4);
          z[i4 * 2] = __fadd2_rn(__fadd2_rn(t3f2(pv.x, pv.y), t3f2(qv.x, qv.y)), t3u2f(v[4 * i4], v[4 * i4 + 1]));
          z[i4 * 2 + 1] = __fadd2_rn(__fadd2_rn(t3f2(pv.z, pv.w), t3f2(qv.z, qv.w)), t3u2f(v[4 * i4 + 2], v[4 * i4 + 3]));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_d2c) : "memory");      // D2 may be overwritten
      }
      // ---- the next tile's new units replace buffers this tile may have read: every worker is past its reads first
      if (it + 1 < cnt) {
        const int ul_n = rec0(it + 1, 3).w;
        if (ul_n > ul) {
          t3_sync(T3_BAR_WORKERS, ATC_THREADS);
          if (warp == 0) load_units(it + 1, ul);
        }
      }
      ul_prev = ul;
      // ---- LayerNorm with ONE exchange (single-pass statistics: the rows are centred up to the small angular term), ReLU
      {
        float2 s1 = t3f2(0.f, 0.f), s2 = t3f2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) { s1 = __fadd2_rn(s1, z[i]); s2 = __ffma2_rn(z[i], z[i], s2); }
        float2* st = sm.stat + (it & 1) * 512 + r;      // [parity][slice][row]: every access below is a contiguous 256 bytes per warp
        st[s * 128] = make_float2(s1.x + s1.y, s2.x + s2.y);
        quad_barrier(q);
        const float2 u0 = st[0], u1 = st[128], u2 = st[256], u3 = st[384];
        const float mu = ((u0.x + u1.x) + (u2.x + u3.x)) * (1.0f / H);
        const float var = fmaxf(((u0.y + u1.y) + (u2.y + u3.y)) * (1.0f / H) - mu * mu, 0.f);
        const float rstd = rsqrtf(var + LN_EPS);
        const float2 rs2 = t3f2(rstd, rstd), nm2 = t3f2(-mu * rstd, -mu * rstd);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 g = ld4(sm.gamma + s * 32 + i4 * 4), b = ld4(sm.beta + s * 32 + i4 * 4);
          float2 w0 = __ffma2_rn(z[i4 * 2], rs2, nm2), w1 = __ffma2_rn(z[i4 * 2 + 1], rs2, nm2);      // (z - mu) * rstd
          w0 = __ffma2_rn(w0, t3f2(g.x, g.y), t3f2(b.x, b.y));
          w1 = __ffma2_rn(w1, t3f2(g.z, g.w), t3f2(b.z, b.w));
          z[i4 * 2] = t3f2(fmaxf(w0.x, 0.f), fmaxf(w0.y, 0.f));
          z[i4 * 2 + 1] = t3f2(fmaxf(w1.x, 0.f), fmaxf(w1.y, 0.f));
        }
      }
      // ---- TF32 split of the hidden activations BEFORE the wait on the tensor core: hi = z truncated to TF32, lo = z - hi (exact)
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        hi[2 * i] = __float_as_uint(z[i].x) & 0xffffe000u;
        hi[2 * i + 1] = __float_as_uint(z[i].y) & 0xffffe000u;
        const float2 l = __fadd2_rn(z[i], t3f2(-__uint_as_float(hi[2 * i]), -__uint_as_float(hi[2 * i + 1])));
        lo[2 * i] = __float_as_uint(l.x); lo[2 * i + 1] = __float_as_uint(l.y);
      }
      int4 rp = make_int4(-1, -1, 0, 0);
      bool prev_ok = false;
      if (it > 0) {
        rp = rec0(it - 1, q);
        prev_ok = (rp.z >> lane) & 1;
      }
      // ---- the only work between "main MMA of the previous tile retired" and "main MMA of this tile may start": store the new A
      // operand, pull the previous D into registers.  Everything else (logits / weighted sums, softmax) runs under the next MMA.
      uint32_t v[32];
      {
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (it > 0) { T3_WAIT(bar_mma, (it - 1) & 1); tc_fence_after(); }
        tmem_st32(lane_addr + ATC_COL_AHI + s * 32, hi);
        tmem_st32(lane_addr + ATC_COL_ALO + s * 32, lo);
        if (it > 0) tmem_ld32(lane_addr + ATC_COL_D + s * 32, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (it > 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        t3_arrive(T3_BAR_A_READY, T3_SYNC);
      }
      // ---- epilogue of the previous tile from registers while the tensor core works
      if (it > 0) {
        const int ptb = tb - 4 * 32;      // same quadrant, previous tile
        if (!VPASS) {
          float lg[4];
          const float* qr = reinterpret_cast<const float*>(ring + T3_RING_X + ((it - 1) & 1) * 128);
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            const float4 q0 = ld4(qr + hh * 8), q1 = ld4(qr + hh * 8 + 4);
            float2 acc = __fmul2_rn(t3f2(q0.x, q0.y), t3u2f(v[hh * 8], v[hh * 8 + 1]));
            acc = __ffma2_rn(t3f2(q0.z, q0.w), t3u2f(v[hh * 8 + 2], v[hh * 8 + 3]), acc);
            acc = __ffma2_rn(t3f2(q1.x, q1.y), t3u2f(v[hh * 8 + 4], v[hh * 8 + 5]), acc);
            acc = __ffma2_rn(t3f2(q1.z, q1.w), t3u2f(v[hh * 8 + 6], v[hh * 8 + 7]), acc);
            lg[hh] = prev_ok ? acc.x + acc.y : -INFINITY;
          }
          finish_k(lg, prev_ok, rp.x, ptb, rp.y);
        } else {
          float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
          cpa_wait_all();      // weights / factor / residual row requested at the top of this iteration
          __syncwarp();
          if (prev_ok) {
            w4 = ld4(reinterpret_cast<const float*>(ring + T3_RING_X) + lane * 4);
            if (rp.y >= 0) w4 = mul4(w4, ld4(reinterpret_cast<const float*>(ring + T3_RING_X + 512)));
          }
          float val[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float wh = (i < 8) ? w4.x : (i < 16) ? w4.y : (i < 24) ? w4.z : w4.w;
            val[i] = wh * __uint_as_float(v[i]);
          }
          warp_reduce_scatter<32>(val, lane);
          if (rp.x >= 0) {
            const int c = s * 32 + lane;
            if (rp.y >= 0) {
              a.part[(size_t)(ptb >> 5) * H + c] = val[0];      // chunked group: launch_trip_combine finishes the edge
            } else {
              const float hb_in = *reinterpret_cast<const float*>(ring + T3_RING_X + 576 + lane * 4);
              const float upd = rp.z != 0 ? val[0] + sm.b2[c] : 0.f;
              a.h_bond_out[(size_t)rp.x * H + c] = hb_in + upd;      // :274
            }
          }
        }
      }
    }
    // ---- epilogue of the last tile
    if (it > 0) {
      const int4 rp = rec0(it - 1, q);
      const bool prev_ok = (rp.z >> lane) & 1;
      const int ptb = ((t0 + it - 1) * 4 + q) * 32;
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      if (!VPASS) {
        cpa_wait_all();
        __syncwarp();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float* qr = reinterpret_cast<const float*>(ring + T3_RING_X + ((it - 1) & 1) * 128);
        float lg[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const float4 q0 = ld4(qr + hh * 8), q1 = ld4(qr + hh * 8 + 4);
          float2 acc = __fmul2_rn(t3f2(q0.x, q0.y), t3u2f(v[hh * 8], v[hh * 8 + 1]));
          acc = __ffma2_rn(t3f2(q0.z, q0.w), t3u2f(v[hh * 8 + 2], v[hh * 8 + 3]), acc);
          acc = __ffma2_rn(t3f2(q1.x, q1.y), t3u2f(v[hh * 8 + 4], v[hh * 8 + 5]), acc);
          acc = __ffma2_rn(t3f2(q1.z, q1.w), t3u2f(v[hh * 8 + 6], v[hh * 8 + 7]), acc);
          lg[hh] = prev_ok ? acc.x + acc.y : -INFINITY;
        }
        finish_k(lg, prev_ok, rp.x, ptb, rp.y);
      } else {
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prev_ok) {
          w4 = ld4(a.wbuf + ((size_t)ptb + lane) * NH + s * 4);
          if (rp.y >= 0) w4 = mul4(w4, ld4(a.factor + (size_t)(ptb >> 5) * NH + s * 4));
        }
        float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
        if (rp.x >= 0) {
          const int c = s * 32 + lane;
          if (rp.y >= 0) {
            a.part[(size_t)(ptb >> 5) * H + c] = tot;
          } else {
            float upd = rp.z != 0 ? tot + sm.b2[c] : 0.f;
            a.h_bond_out[(size_t)rp.x * H + c] = a.h_bond_in[(size_t)rp.x * H + c] + upd;
          }
        }
      }
    }
    cpa_wait_all();
  }
  tc_fence_before();
  __syncthreads();      // also orders this phase's attention weights before the next phase's reads within the CTA
  if (last && warp == 0) tmem_dealloc(tmem_base, 512);
}

template <bool VPASS>
__global__ void __launch_bounds__(T3_THREADS, 1) trip3_kernel(const TripArgs a) { trip3_body<VPASS>(a, true, true); }
__global__ void __launch_bounds__(T3_THREADS, 1) trip3_pair_kernel(const TripArgs a) {
  trip3_body<false>(a, true, false);
  trip3_body<true>(a, false, true);
}

// key + value pass in one launch: every CTA walks the same tiles in both phases (not for chunked groups, whose rescale factors
// need every CTA's key pass)
void launch_trip3_pair(const TripArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0 || a.n_tiles3 <= 0) return;
  static DeviceOnce once;
  const int bytes = Trip3Smem<true>::bytes();
  if (!once.done()) { cudaFuncSetAttribute(trip3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  launch_pdl(trip3_pair_kernel, dim3(atc_grid(a.n_tiles3, num_sms)), dim3(T3_THREADS), bytes, stream, a);
}

void launch_trip3(const TripArgs& a, bool vpass, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0 || a.n_tiles3 <= 0) return;
  static DeviceOnce once;
  const int bytes = Trip3Smem<true>::bytes();
  if (!once.done()) {
    cudaFuncSetAttribute(trip3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(trip3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    once.mark();
  }
  const int grid = atc_grid(a.n_tiles3, num_sms);
  if (vpass) launch_pdl(trip3_kernel<true>, dim3(grid), dim3(T3_THREADS), bytes, stream, a);
  else launch_pdl(trip3_kernel<false>, dim3(grid), dim3(T3_THREADS), bytes, stream, a);
}

// host-side packing of Wa[13][128] (first-Linear columns of the angular encoding, transposed) into the B operand of the
// angular MMA: rows n = output channel, K = 16 features (13 used), hi | lo, 64-byte rows with the 64B swizzle
void pack_wa_sw64(const float* Wa, float* out /* 2*128*16 floats */) {
  for (int i = 0; i < 2 * 128 * 16; ++i) out[i] = 0.f;
  float* hi = out;
  float* lo = out + 128 * 16;
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < NANG; ++k) {
      float w = Wa[k * H + n];
      float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
      int off = (n * 64 + ((((k >> 2) ^ ((n >> 1) & 3))) << 4) + (k & 3) * 4) / 4;
      hi[off] = h; lo[off] = l;
    }
}

}  // namespace ddb
