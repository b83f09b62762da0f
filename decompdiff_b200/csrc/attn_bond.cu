// K3: attention over the ligand bond graph.
//   bond_attn_node : NodeUpdateLayer with edge_feat = h_bond      (uni_transformer_edge.py:42-74, called at :273)
//   bond_attn_pos  : PosUpdateLayer  with edge_feat = new h_bond  (:188-210, called at :283) + the x update (:284-285)
//   trip_prep/k/v  : BondUpdateLayer over triplets k->j->i        (:125-167, called at :274)
//
// Same decomposition as the kNN kernels: the first Linear is split into per-edge / per-atom projections
// produced by the GEMMs, the second Linear is contracted with the query (keys) or applied after the
// weighted sum (values).  The query MLP of the bond layer is evaluated per EDGE (the reference recomputes
// it for every triplet, :149,157).  One warp owns one softmax group (a ligand atom / a bond edge).
#include "kernels.cuh"

namespace ddb {

constexpr int BND_WARPS = 12;
constexpr int BND_THREADS = BND_WARPS * 32;

__device__ __forceinline__ float compb(const float4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }

// U[h] = sum_{c in head h} q[c] * W2k[c, lane*4..+3]   (q staged in the warp's shared scratch)
__device__ __forceinline__ void key_contract(const float* qs, const float* sW2, int lane, float4 (&U)[NH]) {
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    U[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4) {
      float4 qv = ld4(qs + h * DH + c4 * 4);
#pragma unroll
      for (int c = 0; c < 4; ++c) U[h] = fma4(compb(qv, c), ld4(sW2 + (size_t)(h * DH + c4 * 4 + c) * H + lane * 4), U[h]);
    }
  }
}

// logits of a 4-row block -> wbuf rows (row index base+s), invalid rows skipped
__device__ __forceinline__ void store_logits4(const float4 (&U)[NH], const float4 (&z)[4], int lane, float* wbuf,
                                              const int (&row)[4], const bool (&ok)[4]) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float part[32];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
      for (int h = 0; h < NH; ++h) part[s2 * NH + h] = dot4(U[h], z[p * 2 + s2]);
    warp_reduce_scatter<32>(part, lane);
    const bool hi = lane >> 4;
    const int r = hi ? row[p * 2 + 1] : row[p * 2];
    const bool v = hi ? ok[p * 2 + 1] : ok[p * 2];
    if (v) wbuf[(size_t)r * NH + (lane & 15)] = part[0];
  }
}

// out[c] = <W2[c,:], S[c/8]> + b2[c] * wsum ; lane returns c = chunk*32 + lane for chunk 0..3 in o[4]
__device__ __forceinline__ void value_contract(const float* sW2, const float4 (&S)[NH], int lane, float (&o)[4]) {
#pragma unroll
  for (int chunk = 0; chunk < 4; ++chunk) {
    float part[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      int c = chunk * 32 + i;
      part[i] = dot4(ld4(sW2 + (size_t)c * H + lane * 4), S[c / DH]);
    }
    warp_reduce_scatter<32>(part, lane);
    o[chunk] = part[0];
  }
}

// ---------------------------------------------------------------------------------- bond-edge attention
struct BondSmem {
  float* W2k; float* W2v; float* gk; float* bk; float* gv; float* bv; float* scratch;
  __device__ BondSmem(float* base, bool with_v) {
    W2k = base; W2v = W2k + H * H; gk = W2v + (with_v ? H * H : 0); bk = gk + H; gv = bk + H; bv = gv + H; scratch = bv + H;
  }
  static int bytes(bool with_v) { return (H * H + (with_v ? H * H : 0) + 4 * H + BND_WARPS * H) * 4; }
};

__device__ __forceinline__ void bond_hidden4(const BondAttnArgs& a, const BondSide& side, float4 hi, float4 gam, float4 bet,
                                             int lane, int s0, int s_end, float4 (&z)[4]) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    int slot = min(s0 + s, s_end - 1);
    int src = __ldg(a.in_src + slot), eid = __ldg(a.in_eid + slot);
    z[s] = add4(add4(hi, ldg4(side.Hj + (size_t)src * a.ldh + lane * 4)), ldg4(side.Pe + (size_t)eid * a.ldpe + lane * 4));
  }
  ln_relu_rows<4>(z, gam, bet, lane);
}

// softmax over wbuf rows [s_begin, s_end) per head (in place); rows flagged by skip_src are excluded
__device__ __forceinline__ void group_softmax(float* wbuf, int s_begin, int s_end, int lane, const int* in_src, int skip_src) {
  const int h = lane & 15, half = lane >> 4;
  float m = -INFINITY;
  for (int s = s_begin + half; s < s_end; s += 2)
    if (skip_src < 0 || __ldg(in_src + s) != skip_src) m = fmaxf(m, __ldcg(wbuf + (size_t)s * NH + h));
  m = fmaxf(m, __shfl_xor_sync(FULL, m, 16));
  float ssum = 0.f;
  for (int s = s_begin + half; s < s_end; s += 2)
    if (skip_src < 0 || __ldg(in_src + s) != skip_src) ssum += expf(__ldcg(wbuf + (size_t)s * NH + h) - m);
  ssum += __shfl_xor_sync(FULL, ssum, 16);
  for (int s = s_begin + half; s < s_end; s += 2) {
    bool skip = skip_src >= 0 && __ldg(in_src + s) == skip_src;
    wbuf[(size_t)s * NH + h] = skip ? 0.f : expf(__ldcg(wbuf + (size_t)s * NH + h) - m) / ssum;
  }
}

template <bool POS>
__global__ void __launch_bounds__(BND_THREADS, 1) bond_attn_kernel(const BondAttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  BondSmem sm(smem, !POS);
  cta_copy_f4(sm.W2k, a.k.w.W2, H * H);
  if (!POS) cta_copy_f4(sm.W2v, a.v.w.W2, H * H);
  cta_copy_f4(sm.gk, a.k.w.gamma, H); cta_copy_f4(sm.bk, a.k.w.beta, H);
  cta_copy_f4(sm.gv, a.v.w.gamma, H); cta_copy_f4(sm.bv, a.v.w.beta, H);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* qs = sm.scratch + warp * H;
  const float4 gk = ld4(sm.gk + lane * 4), bk = ld4(sm.bk + lane * 4), gv = ld4(sm.gv + lane * 4), bv = ld4(sm.bv + lane * 4);
  float4 Wv[POS ? NH : 1];
  float b2x = 0.f;
  if (POS) {
#pragma unroll
    for (int h = 0; h < NH; ++h) Wv[h] = ldg4(a.v.w.W2 + (size_t)h * H + lane * 4);
    b2x = __ldg(a.v.w.b2 + (lane & 15));
  }

  for (int at = blockIdx.x * BND_WARPS + warp; at < a.n_lig; at += gridDim.x * BND_WARPS) {
    const int s_begin = a.in_ptr[at], s_end = a.in_ptr[at + 1];
    const int node = a.lig_idx[at];
    float acc[3] = {0.f, 0.f, 0.f};
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (s_end > s_begin) {
      __syncwarp();
      st4(qs + lane * 4, ldg4(a.q + (size_t)at * a.ldq + lane * 4));
      __syncwarp();
      {
        float4 U[NH];
        key_contract(qs, sm.W2k, lane, U);
        const float4 hik = ldg4(a.k.Hi + (size_t)at * a.ldh + lane * 4);
#pragma unroll 1
        for (int s0 = s_begin; s0 < s_end; s0 += 4) {
          float4 z[4];
          bond_hidden4(a, a.k, hik, gk, bk, lane, s0, s_end, z);
          int row[4]; bool ok[4];
#pragma unroll
          for (int s = 0; s < 4; ++s) { row[s] = s0 + s; ok[s] = s0 + s < s_end; }
          store_logits4(U, z, lane, a.wbuf, row, ok);
        }
      }
      __syncwarp();
      group_softmax(a.wbuf, s_begin, s_end, lane, a.in_src, -1);
      __syncwarp();
      const float4 hiv = ldg4(a.v.Hi + (size_t)at * a.ldh + lane * 4);
      if (!POS) {
        float4 S[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) S[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int s0 = s_begin; s0 < s_end; s0 += 4) {
          float4 z[4];
          bond_hidden4(a, a.v, hiv, gv, bv, lane, s0, s_end, z);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            if (s0 + s < s_end) {
#pragma unroll
              for (int h4 = 0; h4 < NH / 4; ++h4) {
                float4 w = __ldcg(reinterpret_cast<const float4*>(a.wbuf + (size_t)(s0 + s) * NH + h4 * 4));
                S[h4 * 4 + 0] = fma4(w.x, z[s], S[h4 * 4 + 0]); S[h4 * 4 + 1] = fma4(w.y, z[s], S[h4 * 4 + 1]);
                S[h4 * 4 + 2] = fma4(w.z, z[s], S[h4 * 4 + 2]); S[h4 * 4 + 3] = fma4(w.w, z[s], S[h4 * 4 + 3]);
              }
            }
          }
        }
        value_contract(sm.W2v, S, lane, o);
      } else {
        const float4 xi = ldg4(a.x4 + (size_t)node * 4);
#pragma unroll 1
        for (int s0 = s_begin; s0 < s_end; s0 += 4) {
          float4 z[4];
          bond_hidden4(a, a.v, hiv, gv, bv, lane, s0, s_end, z);
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            float part[32];
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
              for (int h = 0; h < NH; ++h) part[s2 * NH + h] = dot4(Wv[h], z[p * 2 + s2]);
            warp_reduce_scatter<32>(part, lane);
            const int slot = s0 + p * 2 + (lane >> 4);
            if (slot < s_end) {
              float c = __ldcg(a.wbuf + (size_t)slot * NH + (lane & 15)) * (part[0] + b2x);
              float4 xj = ldg4(a.x4 + (size_t)a.lig_idx[__ldg(a.in_src + slot)] * 4);
              acc[0] = fmaf(c, xi.x - xj.x, acc[0]);
              acc[1] = fmaf(c, xi.y - xj.y, acc[1]);
              acc[2] = fmaf(c, xi.z - xj.z, acc[2]);
            }
          }
        }
      }
    }
    if (!POS) {
      if (s_end > s_begin) {
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
          int c = chunk * 32 + lane;
          float* dst = a.out_h + (size_t)node * a.ldo + c;
          *dst = *dst + (o[chunk] + __ldg(a.v.w.b2 + c));   // sum of softmax weights is 1
        }
      }
    } else {
      acc[0] = warp_sum(acc[0]); acc[1] = warp_sum(acc[1]); acc[2] = warp_sum(acc[2]);
      if (lane == 0) {
        float4 xi = ldg4(a.x4 + (size_t)node * 4);
        float4 de = a.dx_edge ? ld4(a.dx_edge + (size_t)at * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float mk = (a.upd_mask == nullptr || a.upd_mask[at]) ? 1.f : 0.f;
        xi.x += (de.x + acc[0] * (1.f / NH)) * mk;       // x + (dx_edge + dx_bond) * mask   (:284-285)
        xi.y += (de.y + acc[1] * (1.f / NH)) * mk;
        xi.z += (de.z + acc[2] * (1.f / NH)) * mk;
        st4(a.x4_out + (size_t)node * 4, xi);
      }
    }
  }
}

static int bnd_grid(int groups, int num_sms) {
  int need = (groups + BND_WARPS - 1) / BND_WARPS;
  return need < num_sms ? (need > 0 ? need : 1) : num_sms;
}

void launch_bond_attn_node(const BondAttnArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_lig <= 0) return;
  static DeviceOnce once;
  int bytes = BondSmem::bytes(true);
  if (!once.done()) { cudaFuncSetAttribute(bond_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  bond_attn_kernel<false><<<bnd_grid(a.n_lig, num_sms), BND_THREADS, bytes, stream>>>(a);
}
void launch_bond_attn_pos(const BondAttnArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_lig <= 0) return;
  static DeviceOnce once;
  int bytes = BondSmem::bytes(false);
  if (!once.done()) { cudaFuncSetAttribute(bond_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  bond_attn_kernel<true><<<bnd_grid(a.n_lig, num_sms), BND_THREADS, bytes, stream>>>(a);
}

// ---------------------------------------------------------------------------------------- triplets
// prep: P[e] = Pe[e] + Hk[src(e)] + Hj[dst(e)] + Wd . gauss(d_e)   for the k and the v MLP.  One warp handles PREP_EDGES edges at
// a time (lane = 4 channels) so that every weight row fetched from L1 feeds several edges - the kernel is bound by the L1
// wavefronts of the Wd / Wc reads otherwise.
constexpr int PREP_EDGES = 2;      // 2 edges x 64 registers -> 32 warps per SM: the kernel is latency / HBM bound, occupancy matters more than weight reuse
__device__ __forceinline__ float4 fma4p(float s, float4 w, float4 acc) {      // acc + s * w as two packed FFMA2
  const float2 ss = make_float2(s, s);
  const float2 lo = __ffma2_rn(make_float2(w.x, w.y), ss, make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(make_float2(w.z, w.w), ss, make_float2(acc.z, acc.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__global__ void __launch_bounds__(256, 4) trip_prep_kernel(const TripArgs a) {
  const int lane = threadIdx.x & 31;
  const int e0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PREP_EDGES;
  if (e0 >= a.n_bonds) return;
  int ks[PREP_EDGES], js[PREP_EDGES], es[PREP_EDGES];
  float gl[PREP_EDGES];
#pragma unroll
  for (int u = 0; u < PREP_EDGES; ++u) {
    es[u] = min(e0 + u, a.n_bonds - 1);
    int nk, nj;
    if (a.edge_meta) { const int4 em = __ldg(a.edge_meta + es[u]); ks[u] = em.x; js[u] = em.y; nk = em.z; nj = em.w; }
    else { ks[u] = a.bsrc[es[u]]; js[u] = a.bdst[es[u]]; nk = a.lig_idx[ks[u]]; nj = a.lig_idx[js[u]]; }
    float4 xk = ldg4(a.x4 + (size_t)nk * 4), xj = ldg4(a.x4 + (size_t)nj * 4);
    float dx = xj.x - xk.x, dy = xj.y - xk.y, dz = xj.z - xk.z;
    float d = sqrtf(dx * dx + dy * dy + dz * dz);       // (pos[i]-pos[j]).pow(2).sum(-1).sqrt()  (:130)
    gl[u] = lane < NG ? gauss_feat(d, lane) : 0.f;
    if (a.xcsr && lane == 0 && e0 + u < a.n_bonds) st4(a.xcsr + (size_t)a.csr_slot[es[u]] * 4, xk);
  }
  // both MLPs in one sweep over the 20 Gaussians: one broadcast per (edge, Gaussian) serves the four weight rows, and the
  // accumulations run as packed FFMA2 (same per-element fma as before, bit-identical results)
  float4 z[2][PREP_EDGES], qv[2][PREP_EDGES];
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const TripSide& t = side ? a.v : a.k;
#pragma unroll
    for (int u = 0; u < PREP_EDGES; ++u) {
      z[side][u] = add4(add4(ldg4(t.Pe + (size_t)es[u] * a.ldpe + lane * 4), ldg4(t.Hk + (size_t)ks[u] * a.ldh + lane * 4)),
                        ldg4(t.Hj + (size_t)js[u] * a.ldh + lane * 4));
      qv[side][u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const bool tc = a.k.Q != nullptr;      // the tensor-core kernels take both sides or none (ddb_batch_create)
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float gg[PREP_EDGES];
#pragma unroll
    for (int u = 0; u < PREP_EDGES; ++u) gg[u] = __shfl_sync(FULL, gl[u], g);
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      const TripSide& t = side ? a.v : a.k;
      const float4 wd = ldg4(t.Wd + g * H + lane * 4);
      float4 wc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tc) wc = ldg4(t.Wc + g * H + lane * 4);
#pragma unroll
      for (int u = 0; u < PREP_EDGES; ++u) {
        z[side][u] = fma4p(gg[u], wd, z[side][u]);
        if (tc) qv[side][u] = fma4p(gg[u], wc, qv[side][u]);
      }
    }
  }
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const TripSide& t = side ? a.v : a.k;
#pragma unroll
    for (int u = 0; u < PREP_EDGES; ++u) {
      if (t.Q) {      // tensor-core kernels: rows are stored CENTRED (LayerNorm is invariant to a per-row shift, and centred rows
        // keep its single-pass statistics well conditioned); Q = Wc . gauss(d) is the same edge seen as j->i
        const float4 zz = z[side][u], qq = qv[side][u];
        const float pm = warp_sum((zz.x + zz.y) + (zz.z + zz.w)) * (1.0f / H);
        const float qm = warp_sum((qq.x + qq.y) + (qq.z + qq.w)) * (1.0f / H);
        if (e0 + u < a.n_bonds) {
          float* prow = t.Pcsr ? t.Pcsr + (size_t)a.csr_slot[es[u]] * H : t.P + (size_t)es[u] * H;
          st4(prow + lane * 4, make_float4(zz.x - pm, zz.y - pm, zz.z - pm, zz.w - pm));
          st4(t.Q + (size_t)es[u] * H + lane * 4, make_float4(qq.x - qm, qq.y - qm, qq.z - qm, qq.w - qm));
        }
      } else if (e0 + u < a.n_bonds) {
        st4(t.P + (size_t)es[u] * H + lane * 4, z[side][u]);
      }
    }
  }
}
void launch_trip_prep(const TripArgs& a, cudaStream_t stream) {
  if (a.n_bonds <= 0) return;
  const int per_block = 8 * PREP_EDGES;
  trip_prep_kernel<<<(a.n_bonds + per_block - 1) / per_block, 256, 0, stream>>>(a);
}

struct TripSmem {
  float* W2; float* Wc; float* Wa; float* gamma; float* beta; float* scratch;
  static constexpr int kWarpFloats = H + 32 * 16;     // q row + angular features of up to 32 triplets
  __device__ TripSmem(float* base) {
    W2 = base; Wc = W2 + H * H; Wa = Wc + NG * H; gamma = Wa + 16 * H; beta = gamma + H; scratch = beta + H;
  }
  static int bytes() { return (H * H + NG * H + 16 * H + 2 * H + BND_WARPS * kWarpFloats) * 4; }
};

__device__ __forceinline__ void load_trip_weights(const TripSide& t, const TripSmem& sm) {
  cta_copy_f4(sm.W2, t.w.W2, H * H);
  cta_copy_f4(sm.Wc, t.Wc, NG * H);
  cta_copy_f4(sm.Wa, t.Wa, NANG * H);
  cta_copy_f4(sm.gamma, t.w.gamma, H);
  cta_copy_f4(sm.beta, t.w.beta, H);
  __syncthreads();
}

// Q = Wc . gauss(d_ji)   (the part of the first Linear that depends on the edge j->i only)
__device__ __forceinline__ float4 trip_q_term(const TripSmem& sm, float4 xi, float4 xj, int lane) {
  float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
  float d = sqrtf(dx * dx + dy * dy + dz * dz);
  float gl = lane < NG ? gauss_feat(d, lane) : 0.f;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int g = 0; g < NG; ++g) q = fma4(__shfl_sync(FULL, gl, g), ld4(sm.Wc + g * H + lane * 4), q);
  return q;
}

// AngularEncoding of the angle at i between (j - i) and (k - i)  (:133-140, common.py:46-54); lane = triplet
__device__ __forceinline__ void trip_angles(float* ang, float4 xi, float4 xj, float4 xk, int lane, bool valid) {
  if (valid) {
    float ax = xj.x - xi.x, ay = xj.y - xi.y, az = xj.z - xi.z;
    float bx = xk.x - xi.x, by = xk.y - xi.y, bz = xk.z - xi.z;
    float dot = ax * bx + ay * by + az * bz;
    float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    float th = atan2f(sqrtf(cx * cx + cy * cy + cz * cz), dot);
    const float f[6] = {1.f, 2.f, 3.f, 1.f, 0.5f, (float)(1.0 / 3.0)};
    float* o = ang + lane * 16;
    o[0] = th;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float sv, cv;
      sincosf(th * f[i], &sv, &cv);
      o[1 + i] = sv; o[7 + i] = cv;
    }
  }
}

// hidden activations of 4 triplets (chunk-local indices t0..t0+3, clamped to n_t-1)
__device__ __forceinline__ void trip_hidden4(const TripArgs& a, const float* P, const TripSmem& sm, const float* ang,
                                             float4 qterm, float4 gam, float4 bet, int lane, int slot0, int t0, int n_t,
                                             float4 (&z)[4]) {
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    int t = min(t0 + s, n_t - 1);
    int eid = __ldg(a.in_eid + slot0 + t);
    z[s] = add4(ldg4(P + (size_t)eid * H + lane * 4), qterm);
  }
#pragma unroll
  for (int ab = 0; ab < 4; ++ab) {
    float4 a4[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) a4[s] = ld4(ang + min(t0 + s, n_t - 1) * 16 + ab * 4);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (ab * 4 + c < NANG) {
        float4 w = ld4(sm.Wa + (ab * 4 + c) * H + lane * 4);
#pragma unroll
        for (int s = 0; s < 4; ++s) z[s] = fma4(compb(a4[s], c), w, z[s]);
      }
    }
  }
  ln_relu_rows<4>(z, gam, bet, lane);
}

template <bool VPASS>
__global__ void __launch_bounds__(BND_THREADS, 1) trip_kernel(const TripArgs a) {
  extern __shared__ __align__(16) float smem[];
  TripSmem sm(smem);
  const TripSide& side = VPASS ? a.v : a.k;
  load_trip_weights(side, sm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* qs = sm.scratch + warp * TripSmem::kWarpFloats;
  float* ang = qs + H;
  const float4 gam = ld4(sm.gamma + lane * 4), bet = ld4(sm.beta + lane * 4);

  for (int e = blockIdx.x * BND_WARPS + warp; e < a.n_bonds; e += gridDim.x * BND_WARPS) {
    const int j = a.bsrc[e], i = a.bdst[e];       // edge j -> i; triplets over edges k -> j, k != i
    const int s_begin = a.in_ptr[j], s_end = a.in_ptr[j + 1];
    const int base = a.trip_base[e];
    const float4 xi = ldg4(a.x4 + (size_t)a.lig_idx[i] * 4), xj = ldg4(a.x4 + (size_t)a.lig_idx[j] * 4);
    const float4 qterm = trip_q_term(sm, xi, xj, lane);
    float4 U[VPASS ? 1 : NH];
    float4 S[VPASS ? NH : 1];
    int n_valid = 0;
    if constexpr (!VPASS) {
      __syncwarp();
      st4(qs + lane * 4, ldg4(a.q + (size_t)e * a.ldq + lane * 4));
      __syncwarp();
      key_contract(qs, sm.W2, lane, U);
    } else {
#pragma unroll
      for (int h = 0; h < NH; ++h) S[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int c0 = s_begin; c0 < s_end; c0 += 32) {
      const int n_t = min(32, s_end - c0);
      __syncwarp();
      {
        bool valid = lane < n_t;
        int ksrc = valid ? __ldg(a.in_src + c0 + lane) : 0;
        float4 xk = ldg4(a.x4 + (size_t)a.lig_idx[ksrc] * 4);
        trip_angles(ang, xi, xj, xk, lane, valid && ksrc != i);
        if (valid && ksrc == i) {   // excluded triplet (i == k): keep finite features, its weight is zero
#pragma unroll
          for (int q = 0; q < 16; ++q) ang[lane * 16 + q] = 0.f;
        }
      }
      __syncwarp();
#pragma unroll 1
      for (int t0 = 0; t0 < n_t; t0 += 4) {
        float4 z[4];
        trip_hidden4(a, side.P, sm, ang, qterm, gam, bet, lane, c0, t0, n_t, z);
        if constexpr (!VPASS) {
          int row[4]; bool ok[4];
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            row[s] = base + (c0 - s_begin) + t0 + s;
            ok[s] = (t0 + s < n_t) && (__ldg(a.in_src + c0 + min(t0 + s, n_t - 1)) != i);
          }
          store_logits4(U, z, lane, a.wbuf, row, ok);
        } else {
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            if (t0 + s < n_t && __ldg(a.in_src + c0 + t0 + s) != i) {
              ++n_valid;
              const float* wr = a.wbuf + (size_t)(base + (c0 - s_begin) + t0 + s) * NH;
#pragma unroll
              for (int h4 = 0; h4 < NH / 4; ++h4) {
                float4 w = ld4(wr + h4 * 4);
                S[h4 * 4 + 0] = fma4(w.x, z[s], S[h4 * 4 + 0]); S[h4 * 4 + 1] = fma4(w.y, z[s], S[h4 * 4 + 1]);
                S[h4 * 4 + 2] = fma4(w.z, z[s], S[h4 * 4 + 2]); S[h4 * 4 + 3] = fma4(w.w, z[s], S[h4 * 4 + 3]);
              }
            }
          }
        }
      }
    }
    if constexpr (!VPASS) {
      __syncwarp();
      group_softmax(a.wbuf + (size_t)base * NH, 0, s_end - s_begin, lane, a.in_src + s_begin, i);
    } else {
      float o[4];
      value_contract(sm.W2, S, lane, o);
#pragma unroll
      for (int chunk = 0; chunk < 4; ++chunk) {
        int c = chunk * 32 + lane;
        float upd = n_valid > 0 ? o[chunk] + __ldg(side.w.b2 + c) : 0.f;
        a.h_bond_out[(size_t)e * H + c] = a.h_bond_in[(size_t)e * H + c] + upd;      // :274
      }
    }
  }
}

void launch_trip_k(const TripArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0) return;
  static DeviceOnce once;
  int bytes = TripSmem::bytes();
  if (!once.done()) { cudaFuncSetAttribute(trip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  trip_kernel<false><<<bnd_grid(a.n_bonds, num_sms), BND_THREADS, bytes, stream>>>(a);
}
void launch_trip_v(const TripArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_bonds <= 0) return;
  static DeviceOnce once;
  int bytes = TripSmem::bytes();
  if (!once.done()) { cudaFuncSetAttribute(trip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  trip_kernel<true><<<bnd_grid(a.n_bonds, num_sms), BND_THREADS, bytes, stream>>>(a);
}

}  // namespace ddb
