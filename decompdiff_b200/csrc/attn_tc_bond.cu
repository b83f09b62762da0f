// Tensor-core attention over the ligand bond edges: NodeUpdateLayer / PosUpdateLayer with edge_feat = h_bond
// (uni_transformer_edge.py:42-74 called at :273, :188-210 called at :283).  Same tile pipeline as the kNN-edge kernel
// (attn_tc.cuh, attn_tc_knn.cu) without a feature GEMM: the first Linear of a bond edge j->i is the sum of three projected rows,
//   z = P_dst[i] (staged per warp) + P_src[j] (row gather) + P_edge[e] (row gather),
// 32 rows = the edges entering one ligand atom = one softmax group = one TMEM lane quadrant, 4 atoms per 128-row tile.
//   K pass      : logits = <q_i[head], D[row, head]>, softmax over the group -> wbuf          (node and position layers)
//   V_NODE pass : out_h[i] += sum_rows w[row, head(c)] D[row, c] + b2[c] sum_rows w[row, head(c)]
//   V_POS pass  : second Linear has 16 outputs (one scalar per head, xv_func): D = A W2^T with N = 16;
//                 dx_i = mean_heads sum_rows w[row, h] (D[row, h] + b2[h]) (x_i - x_j);  x_i += (dx_edge + dx) * mask   (:284-285)
// The launch is small (n_lig / 4 tiles, ~3 per SM), so rows are gathered at the top of each iteration without cross-tile
// prefetch; the fixed cost is the 128 KB W2 image per CTA.
#include <cstdlib>

#include "attn_tc.cuh"

namespace ddb {

constexpr int BT_THREADS = ATC_THREADS + 128;     // 16 worker warps + the issuing warpgroup
constexpr int BT_SYNC = ATC_THREADS + 32;
constexpr int BBAR_A_READY = 6;
enum { BT_K = 0, BT_V_NODE = 1, BT_V_POS = 2 };

struct BondTcSmem {
  uint8_t* W2; float *gamma, *beta, *b2, *hit, *qry; float2* stat; uint64_t* bars; uint32_t* tmem_slot;
  __device__ explicit BondTcSmem(uint8_t* raw) {
    uint8_t* p = raw;
    W2 = p; p += ATC_W2_BYTES;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    hit = reinterpret_cast<float*>(p); p += 16 * 32 * 4;        // per warp: dst-side row slice
    qry = reinterpret_cast<float*>(p); p += 16 * 64 * 4;        // per warp: 2-deep ring of 32-float query slices
    stat = reinterpret_cast<float2*>(p); p += 2 * 128 * 4 * 8;  // [parity][row][slice] {sum, centred sum of squares}
    bars = reinterpret_cast<uint64_t*>(p); p += 64;        // two sets of 4: the second phase of a paired launch uses its own
    tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  static constexpr int bytes() { return ATC_W2_BYTES + (3 * H + 16 * 32 + 16 * 64 + 2 * 128 * 4 * 2) * 4 + 96; }
};

__device__ __forceinline__ float2 bf2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bu2f(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// main MMA with N output columns (128: node / key MLPs, 16: the position value MLP)
template <int N>
__device__ __forceinline__ void bt_issue_mma(uint32_t tmem_base, uint32_t w2_smem, uint32_t bar) {
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t d = tmem_base + ATC_COL_D;
  // B image: hi | lo, each 4 K-blocks of [N rows][128 B]
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

// `first` / `last`: key pass and value pass may run back to back inside one launch (bond_tc_pair_kernel, see attn_tc_knn.cu)
template <int PASS>
__device__ __forceinline__ void bond_tc_body(const BondAttnArgs& a, const bool first, const bool last) {
  constexpr int NOUT = PASS == BT_V_POS ? 16 : 128;
  constexpr int W2_BYTES = 2 * NOUT * 128 * 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  BondTcSmem sm(smem_raw);
  uint64_t* const bars = sm.bars + (first ? 0 : 4);      // a fresh barrier set per phase (no re-initialisation of used barriers)
  const BondSide& side = PASS == BT_K ? a.k : a.v;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = (warp >> 2) & 3, r = q * 32 + lane;
  if ((smem_u32(sm.W2) & 1023u) != 0u) __trap();
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (first && warp == 0) { __syncwarp(); tmem_alloc(smem_u32(sm.tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, W2_BYTES);
    bulk_g2s(smem_u32(sm.W2), side.W2tc, W2_BYTES / 2, bar);
    bulk_g2s(smem_u32(sm.W2) + W2_BYTES / 2, side.W2tc + W2_BYTES / 8, W2_BYTES / 2, bar);
  }
  const uint32_t tmem_base = *sm.tmem_slot;
  cta_copy_f4(sm.gamma, side.w.gamma, H);
  cta_copy_f4(sm.beta, side.w.beta, H);
  if (PASS == BT_V_NODE) cta_copy_f4(sm.b2, side.w.b2, H);
  if (PASS == BT_V_POS && tid < 16) sm.b2[tid] = side.w.b2[tid];
  if (first) pdl_wait();      // set-up on static data above; the previous kernels' results are visible below
  __syncthreads();
  mbar_wait(smem_u32(&bars[0]), 0);
  const uint32_t bar_mma = smem_u32(&bars[1]), w2_smem = smem_u32(sm.W2);
  const int n_tiles = (a.n_vg + 3) / 4;

  if (warp >= 16) {
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
#endif
    if (warp == 16) {
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        asm volatile("bar.sync %0, %1;" ::"r"(BBAR_A_READY), "r"(BT_SYNC) : "memory");
        if (lane == 0) { tc_fence_after(); bt_issue_mma<NOUT>(tmem_base, w2_smem, bar_mma); }
        __syncwarp();
      }
    }
  } else {
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
#endif
    float* const whit = sm.hit + warp * 32;
    float* const wqry = sm.qry + warp * 64;
    const int step = gridDim.x;
    int it = 0;
    // group = (chunk of) the edges entering a ligand atom: {atom (-1: none), first CSR slot, rows, partner chunk or -1};
    // row = {source atom, edge id}
    auto load_group = [&](int tile) {
      const int v = tile * 4 + q;
      int4 g = make_int4(-1, 0, 0, -1);
      if (tile < n_tiles && v < a.n_vg) g = __ldg(a.vg + v);
      return g;
    };
    auto load_row = [&](int4 g) {
      int2 rw = make_int2(0, 0);
      if (lane < g.z) rw = make_int2(__ldg(a.in_src + g.y + lane), __ldg(a.in_eid + g.y + lane));
      return rw;
    };
    // softmax over the rows of the previous group for heads 4s..4s+3 -> wbuf; chunked groups also record {max, sum of exp}
    auto finish_k = [&](const float (&lg)[4], bool ok, int slot0, int vgi, int pair) {
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
      if (ok) st4(a.wbuf + ((size_t)slot0 + lane) * NH + s * 4, make_float4(w[0], w[1], w[2], w[3]));
      if (pair >= 0 && lane < 4)
        a.stats[(size_t)vgi * NH + s * 4 + lane] = make_float2(lane == 0 ? mx[0] : lane == 1 ? mx[1] : lane == 2 ? mx[2] : mx[3],
                                                               lane == 0 ? sum[0] : lane == 1 ? sum[1] : lane == 2 ? sum[2] : sum[3]);
    };
    int4 g = load_group(blockIdx.x);
    int2 rw = load_row(g);
    int4 g_n = load_group(blockIdx.x + step);
    int prev_at = -1, prev_slot0 = 0, prev_vg = 0, prev_pair = -1; bool prev_ok = false;
    float4 prev_rel = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tile = blockIdx.x; tile < n_tiles; tile += step, ++it) {
      const int at = g.x;
      const bool gvalid = at >= 0;
      const bool rowok = lane < g.z;
      const int atc = gvalid ? at : 0;
      // ---- gathers of this tile (small launch: no cross-tile prefetch of rows), metadata of the next one
      float4 pj[8], pe[8];
      {
        const float* prow = side.Hj + (size_t)rw.x * a.ldh + s * 32;
        const float* erow = side.Pe + (size_t)rw.y * a.ldpe + s * 32;
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) ldg8(prow + i8 * 8, pj[2 * i8], pj[2 * i8 + 1]);
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) ldg8(erow + i8 * 8, pe[2 * i8], pe[2 * i8 + 1]);
      }
      const float hi = __ldg(side.Hi + (size_t)atc * a.ldh + s * 32 + lane);
      float qry_v = 0.f;
      if (PASS == BT_K) qry_v = __ldg(a.q + (size_t)atc * a.ldq + s * 32 + lane);
      float4 rel = make_float4(0.f, 0.f, 0.f, 0.f);
      if (PASS == BT_V_POS && s == 0) {      // x_i - x_j of this thread's edge (rel_x, :198-199)
        const float4 xi = ldg4(a.x4 + (size_t)__ldg(a.lig_idx + atc) * 4), xj = ldg4(a.x4 + (size_t)__ldg(a.lig_idx + rw.x) * 4);
        rel = make_float4(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z, 0.f);
      }
      const int2 rw_n = load_row(g_n);
      const int4 g_nn = load_group(tile + 2 * step);
      whit[lane] = hi;
      if (PASS == BT_K) wqry[(it & 1) * 32 + lane] = qry_v;
      __syncwarp();
      // ---- first Linear
      float2 z[16];
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        const float4 hv = ld4(whit + i4 * 4);
        z[i4 * 2] = __fadd2_rn(__fadd2_rn(bf2(pj[i4].x, pj[i4].y), bf2(pe[i4].x, pe[i4].y)), bf2(hv.x, hv.y));
        z[i4 * 2 + 1] = __fadd2_rn(__fadd2_rn(bf2(pj[i4].z, pj[i4].w), bf2(pe[i4].z, pe[i4].w)), bf2(hv.z, hv.w));
      }
      // ---- LayerNorm (one exchange, slice-centred second moments), ReLU
      {
        float2 s1 = bf2(0.f, 0.f), s2 = bf2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) s1 = __fadd2_rn(s1, z[i]);
        const float part = s1.x + s1.y;
        const float mu_s = part * (1.0f / 32.0f);
        const float2 nm = bf2(-mu_s, -mu_s);
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float2 dz = __fadd2_rn(z[i], nm); s2 = __ffma2_rn(dz, dz, s2); }
        float2* st = sm.stat + ((it & 1) * 128 + r) * 4;
        st[s] = make_float2(part, s2.x + s2.y);
        quad_barrier(q);
        const float4 t01 = *reinterpret_cast<const float4*>(st), t23 = *reinterpret_cast<const float4*>(st + 2);
        const float mu = ((t01.x + t01.z) + (t23.x + t23.z)) * (1.0f / H);
        const float d0 = t01.x * (1.0f / 32.0f) - mu, d1 = t01.z * (1.0f / 32.0f) - mu, d2 = t23.x * (1.0f / 32.0f) - mu, d3 = t23.z * (1.0f / 32.0f) - mu;
        const float m2 = ((t01.y + t01.w) + (t23.y + t23.w)) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
        const float rstd = rsqrtf(m2 * (1.0f / H) + LN_EPS);
        const float2 rs2 = bf2(rstd, rstd), nm2 = bf2(-mu * rstd, -mu * rstd);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 gm = ld4(sm.gamma + s * 32 + i4 * 4), bt = ld4(sm.beta + s * 32 + i4 * 4);
          float2 u0 = __ffma2_rn(z[i4 * 2], rs2, nm2), u1 = __ffma2_rn(z[i4 * 2 + 1], rs2, nm2);
          u0 = __ffma2_rn(u0, bf2(gm.x, gm.y), bf2(bt.x, bt.y));
          u1 = __ffma2_rn(u1, bf2(gm.z, gm.w), bf2(bt.z, bt.w));
          z[i4 * 2] = bf2(fmaxf(u0.x, 0.f), fmaxf(u0.y, 0.f));
          z[i4 * 2 + 1] = bf2(fmaxf(u1.x, 0.f), fmaxf(u1.y, 0.f));
        }
      }
      // ---- drain D of the previous tile into registers
      float lg[4] = {0.f, 0.f, 0.f, 0.f};
      float val[32];
      float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 wpos[4];
      if (PASS == BT_V_NODE && it > 0 && prev_ok) {
        w4 = ld4(a.wbuf + ((size_t)prev_slot0 + lane) * NH + s * 4);
        if (prev_pair >= 0) w4 = mul4(w4, ld4(a.factor + (size_t)prev_vg * NH + s * 4));      // chunk -> whole-group softmax
      }
      if (PASS == BT_V_POS && it > 0 && s == 0) {
#pragma unroll
        for (int h4 = 0; h4 < 4; ++h4) {
          wpos[h4] = prev_ok ? ld4(a.wbuf + ((size_t)prev_slot0 + lane) * NH + h4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (prev_pair >= 0) wpos[h4] = mul4(wpos[h4], ld4(a.factor + (size_t)prev_vg * NH + h4 * 4));
        }
      }
      float cpos = 0.f;
      if (it > 0) {
        mbar_wait(bar_mma, (it - 1) & 1);
        tc_fence_after();
        if (PASS == BT_V_POS) {
          if (s == 0) {           // 16 head outputs of this row; c = sum_h w[h] (D[h] + b2[h])
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4) {
              cpos = fmaf(wpos[h4].x, __uint_as_float(v[h4 * 4 + 0]) + sm.b2[h4 * 4 + 0], cpos);
              cpos = fmaf(wpos[h4].y, __uint_as_float(v[h4 * 4 + 1]) + sm.b2[h4 * 4 + 1], cpos);
              cpos = fmaf(wpos[h4].z, __uint_as_float(v[h4 * 4 + 2]) + sm.b2[h4 * 4 + 2], cpos);
              cpos = fmaf(wpos[h4].w, __uint_as_float(v[h4 * 4 + 3]) + sm.b2[h4 * 4 + 3], cpos);
            }
          }
        } else {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (PASS == BT_K) {
            const float* qr = wqry + ((it - 1) & 1) * 32;
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              const float4 q0 = ld4(qr + hh * 8), q1 = ld4(qr + hh * 8 + 4);
              float2 acc = __fmul2_rn(bf2(q0.x, q0.y), bu2f(v[hh * 8], v[hh * 8 + 1]));
              acc = __ffma2_rn(bf2(q0.z, q0.w), bu2f(v[hh * 8 + 2], v[hh * 8 + 3]), acc);
              acc = __ffma2_rn(bf2(q1.x, q1.y), bu2f(v[hh * 8 + 4], v[hh * 8 + 5]), acc);
              acc = __ffma2_rn(bf2(q1.z, q1.w), bu2f(v[hh * 8 + 6], v[hh * 8 + 7]), acc);
              lg[hh] = prev_ok ? acc.x + acc.y : -INFINITY;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float wh = (i < 8) ? w4.x : (i < 16) ? w4.y : (i < 24) ? w4.z : w4.w;
              val[i] = wh * __uint_as_float(v[i]);
            }
          }
        }
      }
      // ---- hidden activations -> TMEM
      {
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi16[16], lo16[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 zz = z[half * 8 + i];
            hi16[2 * i] = __float_as_uint(zz.x) & 0xffffe000u;
            hi16[2 * i + 1] = __float_as_uint(zz.y) & 0xffffe000u;
            const float2 l = __fadd2_rn(zz, bf2(-__uint_as_float(hi16[2 * i]), -__uint_as_float(hi16[2 * i + 1])));
            lo16[2 * i] = __float_as_uint(l.x); lo16[2 * i + 1] = __float_as_uint(l.y);
          }
          tmem_st16(lane_addr + ATC_COL_AHI + s * 32 + half * 16, hi16);
          tmem_st16(lane_addr + ATC_COL_ALO + s * 32 + half * 16, lo16);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        asm volatile("bar.arrive %0, %1;" ::"r"(BBAR_A_READY), "r"(BT_SYNC) : "memory");
      }
      // ---- finish the epilogue of the previous tile
      if (it > 0) {
        if (PASS == BT_K) {
          finish_k(lg, prev_ok, prev_slot0, prev_vg, prev_pair);
        } else if (PASS == BT_V_NODE) {
          float ws[4] = {w4.x, w4.y, w4.z, w4.w};
          warp_allreduce4(ws, lane);
          warp_reduce_scatter<32>(val, lane);
          if (prev_at >= 0) {
            const int c = s * 32 + lane;
            const float res = val[0] + sm.b2[c] * ws[lane >> 3];
            if (prev_pair >= 0) {
              a.part_h[(size_t)prev_vg * H + c] = res;           // chunked atom: launch_bond_combine adds the two chunks
            } else {
              float* dst = a.out_h + (size_t)__ldg(a.lig_idx + prev_at) * a.ldo + c;
              *dst = *dst + res;
            }
          }
        } else if (s == 0) {
          float ax = warp_sum(cpos * prev_rel.x), ay = warp_sum(cpos * prev_rel.y), az = warp_sum(cpos * prev_rel.z);
          if (lane == 0 && prev_at >= 0 && prev_pair >= 0) {
            st4(a.part_dx + (size_t)prev_vg * 4, make_float4(ax * (1.f / NH), ay * (1.f / NH), az * (1.f / NH), 0.f));
          } else if (lane == 0 && prev_at >= 0) {
            const int node = __ldg(a.lig_idx + prev_at);
            float4 xi = ldg4(a.x4 + (size_t)node * 4);
            const float4 de = a.dx_edge ? ld4(a.dx_edge + (size_t)prev_at * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float mk = (a.upd_mask == nullptr || a.upd_mask[prev_at]) ? 1.f : 0.f;
            xi.x += (de.x + ax * (1.f / NH)) * mk;       // x + (dx_edge + dx_bond) * mask   (:284-285)
            xi.y += (de.y + ay * (1.f / NH)) * mk;
            xi.z += (de.z + az * (1.f / NH)) * mk;
            st4(a.x4_out + (size_t)node * 4, xi);
          }
        }
      }
      prev_at = gvalid ? at : -1; prev_slot0 = g.y; prev_ok = rowok; prev_rel = rel; prev_vg = tile * 4 + q; prev_pair = g.w;
      g = g_n; g_n = g_nn; rw = rw_n;
    }
    // ---- epilogue of the last tile (same code path: one more drain without a new A)
    if (it > 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      if (PASS == BT_K) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const float* qr = wqry + ((it - 1) & 1) * 32;
        float lg[4];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < 8; ++d) acc = fmaf(qr[hh * 8 + d], __uint_as_float(v[hh * 8 + d]), acc);
          lg[hh] = prev_ok ? acc : -INFINITY;
        }
        finish_k(lg, prev_ok, prev_slot0, prev_vg, prev_pair);
      } else if (PASS == BT_V_NODE) {
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (prev_ok) {
          w4 = ld4(a.wbuf + ((size_t)prev_slot0 + lane) * NH + s * 4);
          if (prev_pair >= 0) w4 = mul4(w4, ld4(a.factor + (size_t)prev_vg * NH + s * 4));
        }
        float tot = atc_weighted_colsum(tmem_base, q, s, lane, w4);
        float4 ws = make_float4(warp_sum(w4.x), warp_sum(w4.y), warp_sum(w4.z), warp_sum(w4.w));
        if (prev_at >= 0) {
          const int c = s * 32 + lane;
          const float res = tot + sm.b2[c] * sel4(ws, lane >> 3);
          if (prev_pair >= 0) {
            a.part_h[(size_t)prev_vg * H + c] = res;
          } else {
            float* dst = a.out_h + (size_t)__ldg(a.lig_idx + prev_at) * a.ldo + c;
            *dst = *dst + res;
          }
        }
      } else if (s == 0) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float cpos = 0.f;
        if (prev_ok) {
#pragma unroll
          for (int h = 0; h < 16; ++h) {
            float wv = a.wbuf[((size_t)prev_slot0 + lane) * NH + h];
            if (prev_pair >= 0) wv *= __ldg(a.factor + (size_t)prev_vg * NH + h);
            cpos = fmaf(wv, __uint_as_float(v[h]) + sm.b2[h], cpos);
          }
        }
        float ax = warp_sum(cpos * prev_rel.x), ay = warp_sum(cpos * prev_rel.y), az = warp_sum(cpos * prev_rel.z);
        if (lane == 0 && prev_at >= 0 && prev_pair >= 0) {
          st4(a.part_dx + (size_t)prev_vg * 4, make_float4(ax * (1.f / NH), ay * (1.f / NH), az * (1.f / NH), 0.f));
        } else if (lane == 0 && prev_at >= 0) {
          const int node = __ldg(a.lig_idx + prev_at);
          float4 xi = ldg4(a.x4 + (size_t)node * 4);
          const float4 de = a.dx_edge ? ld4(a.dx_edge + (size_t)prev_at * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float mk = (a.upd_mask == nullptr || a.upd_mask[prev_at]) ? 1.f : 0.f;
          xi.x += (de.x + ax * (1.f / NH)) * mk;
          xi.y += (de.y + ay * (1.f / NH)) * mk;
          xi.z += (de.z + az * (1.f / NH)) * mk;
          st4(a.x4_out + (size_t)node * 4, xi);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();      // also orders this phase's attention weights before the next phase's reads within the CTA
  if (last && warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int PASS>
__global__ void __launch_bounds__(BT_THREADS, 1) bond_tc_kernel(const BondAttnArgs a) { bond_tc_body<PASS>(a, true, true); }
template <int P2>
__global__ void __launch_bounds__(BT_THREADS, 1) bond_tc_pair_kernel(const BondAttnArgs a) {
  bond_tc_body<BT_K>(a, true, false);
  bond_tc_body<P2>(a, false, true);
}
template <int P2>
static void launch_bond_tc_pair(const BondAttnArgs& a, int num_sms, cudaStream_t stream) {
  static DeviceOnce once;
  const int bytes = BondTcSmem::bytes();
  if (!once.done()) { cudaFuncSetAttribute(bond_tc_pair_kernel<P2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  const int grid = atc_grid((a.n_vg + 3) / 4, num_sms);
  launch_pdl(bond_tc_pair_kernel<P2>, dim3(grid), dim3(BT_THREADS), bytes, stream, a);
}

template <int PASS>
static void launch_bond_tc_pass(const BondAttnArgs& a, int num_sms, cudaStream_t stream) {
  static DeviceOnce once;
  const int bytes = BondTcSmem::bytes();
  if (!once.done()) { cudaFuncSetAttribute(bond_tc_kernel<PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  const int grid = atc_grid((a.n_vg + 3) / 4, num_sms);
  launch_pdl(bond_tc_kernel<PASS>, dim3(grid), dim3(BT_THREADS), bytes, stream, a);
}

// node variant: key pass then value pass; position variant: key pass then the 16-output value pass + x update
int launch_bond_tc(const BondAttnArgs& a, bool pos, int num_sms, cudaStream_t stream) {      // returns the number of launches
  if (a.n_lig <= 0 || a.n_vg <= 0) return 0;
  const bool chunked = a.n_vg > a.n_lig;      // some atom has more than 32 incoming edges
  static const int pair_mode = getenv("DDB_PAIR") ? atoi(getenv("DDB_PAIR")) : -1;      // see api.cu: pairs pay on small grids only
  static const int pair_waves = getenv("DDB_PAIR_WAVES") ? atoi(getenv("DDB_PAIR_WAVES")) : 4;
  const bool pair = pair_mode >= 0 ? pair_mode != 0 : (a.n_vg + 3) / 4 < pair_waves * num_sms;
  if (!chunked && pair) {      // key + value phase in one launch (a chunked group needs the factors of all CTAs in between)
    if (pos) launch_bond_tc_pair<BT_V_POS>(a, num_sms, stream); else launch_bond_tc_pair<BT_V_NODE>(a, num_sms, stream);
    return 1;
  }
  launch_bond_tc_pass<BT_K>(a, num_sms, stream);
  if (chunked) launch_chunk_factors(a.stats, reinterpret_cast<const int*>(a.vg) + 3, 4, a.n_vg, const_cast<float*>(a.factor), stream);
  if (pos) launch_bond_tc_pass<BT_V_POS>(a, num_sms, stream);
  else launch_bond_tc_pass<BT_V_NODE>(a, num_sms, stream);
  if (chunked) launch_bond_combine(a, pos, stream);
  return chunked ? 4 : 2;
}

// host-side packing of the position value MLP's second Linear W2[16 out][128 in] into the hi | lo swizzled image (N = 16)
void pack_w2x_tc(const float* W2, float* out /* 2*16*128 floats */) {
  float* hi = out;
  float* lo = out + 16 * 128;
  for (int n = 0; n < 16; ++n)
    for (int k = 0; k < 128; ++k) {
      const float w = W2[n * 128 + k];
      const float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
      const int off = sw128_offset_bytes(n, k, 16) / 4;
      hi[off] = h; lo[off] = l;
    }
}

}  // namespace ddb
