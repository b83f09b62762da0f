// K2: 16-head attention over the kNN(32) protein-ligand edges - NodeUpdateLayer / PosUpdateLayer of
// /root/reference/models/encoders/uni_transformer_edge.py:42-74,188-210 evaluated without ever
// materialising the (E,340) MLP inputs or the (E,128) keys / values.
//
// Per destination node (one warp, lane = 4 hidden channels, 4 edges register-blocked per weight load):
//   hidden_e = ReLU(LN(Hi[dst] + Hj[src_e] + Wg[type_e] g(d_e) + Wt[type_e]))        (first Linear decomposed)
//   k pass : logit[e,h] = <U[h], hidden_e>,  U[h] = sum_{c in head h} q[c] W2k[c,:]/sqrt(8)   (key contraction)
//            softmax over the node's edges per head, times e_w  -> wbuf
//   v pass : S[h] += w[e,h] hidden_e ;  out[c] = <W2v[c,:], S[head(c)]> + b2v[c] sum_e w[e,head(c)]
//   pos v  : v[e,h] = <W2xv[h,:], hidden_e> + b ;  dx = mean_h sum_e w[e,h] v[e,h] (x_dst - x_src)
// Edges of a node are ordered ligand-sources-first by the graph kernel, so every 4-edge block has one
// edge type and one set of first-layer weights.  CTAs are persistent (one per SM) and keep all weights
// of the pass in shared memory (~110 KB) for the whole launch.
#include "kernels.cuh"

namespace ddb {

constexpr int ATT_WARPS = 12;
constexpr int ATT_THREADS = ATT_WARPS * 32;

struct KnnSmem {
  float* Wg; float* Wt; float* W2; float* gamma; float* beta; float* warp_scratch;
  static constexpr int kWarpFloats = H + 4 * NG;     // q row + gauss values of the current 4-edge block
  __device__ KnnSmem(float* base, bool with_w2) {
    Wg = base; Wt = Wg + 4 * NG * H; gamma = Wt + 4 * H; beta = gamma + H; W2 = beta + H;
    warp_scratch = W2 + (with_w2 ? H * H : 0);
  }
  static int bytes(bool with_w2) { return (4 * NG * H + 4 * H + 2 * H + (with_w2 ? H * H : 0) + ATT_WARPS * kWarpFloats) * 4; }
};

__device__ __forceinline__ float comp(const float4& v, int c) { return c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : v.w; }

// hidden activations of the 4 edges [e0, e0+4) of `node` (slots >= seg_end are clamped duplicates)
__device__ __forceinline__ void knn_hidden4(const KnnAttnArgs& a, const KnnSmem& sm, float* gbuf, float4 gam, float4 bet,
                                            int lane, int node, float4 xi, float4 hi, int e0, int seg_end, int type,
                                            float4 (&z)[4], float (&rel)[4][3]) {
  int js[4];
  float d[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    int e = min(e0 + s, seg_end - 1);
    js[s] = __ldg(a.nbr + (size_t)node * a.ldn + e);
    float4 xj = ldg4(a.x4 + (size_t)js[s] * 4);
    rel[s][0] = xi.x - xj.x; rel[s][1] = xi.y - xj.y; rel[s][2] = xi.z - xj.z;
    d[s] = sqrtf(rel[s][0] * rel[s][0] + rel[s][1] * rel[s][1] + rel[s][2] * rel[s][2]);
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    int idx = lane + 32 * it;
    if (idx < 4 * NG) {
      int s = idx / NG, g = idx - s * NG;
      float ds = s == 0 ? d[0] : s == 1 ? d[1] : s == 2 ? d[2] : d[3];
      gbuf[idx] = gauss_feat(ds, g);
    }
  }
  __syncwarp();
  const float4 wt = ld4(sm.Wt + type * H + lane * 4);
#pragma unroll
  for (int s = 0; s < 4; ++s) z[s] = add4(add4(hi, ldg4(a.Hj + (size_t)js[s] * a.ldhj + lane * 4)), wt);
  const float* wg = sm.Wg + (size_t)type * NG * H + lane * 4;
#pragma unroll
  for (int gb = 0; gb < NG / 4; ++gb) {
    float4 g4[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) g4[s] = ld4(gbuf + s * NG + gb * 4);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 w = ld4(wg + (gb * 4 + c) * H);
#pragma unroll
      for (int s = 0; s < 4; ++s) z[s] = fma4(comp(g4[s], c), w, z[s]);
    }
  }
  ln_relu_rows<4>(z, gam, bet, lane);
}

__device__ __forceinline__ void load_knn_weights(const KnnAttnArgs& a, const KnnSmem& sm, bool with_w2) {
  cta_copy_f4(sm.Wg, a.w.Wg, 4 * NG * H);
  cta_copy_f4(sm.Wt, a.w.Wt, 4 * H);
  cta_copy_f4(sm.gamma, a.w.gamma, H);
  cta_copy_f4(sm.beta, a.w.beta, H);
  if (with_w2) cta_copy_f4(sm.W2, a.w.W2, H * H);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------ k pass
__global__ void __launch_bounds__(ATT_THREADS, 1) knn_attn_k_kernel(const KnnAttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  KnnSmem sm(smem, true);
  load_knn_weights(a, sm, true);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* qs = sm.warp_scratch + warp * KnnSmem::kWarpFloats;
  float* gbuf = qs + H;
  const float4 gam = ld4(sm.gamma + lane * 4), bet = ld4(sm.beta + lane * 4);

  for (int slot = blockIdx.x * ATT_WARPS + warp; slot < a.n_dst; slot += gridDim.x * ATT_WARPS) {
    const int node = a.dst_list ? a.dst_list[slot] : slot;
    const int deg = a.deg[node];
    if (deg == 0) continue;
    const int nlig = a.nlig[node];
    const bool lig_dst = a.is_lig[node];
    // U[h] = sum_{c in head h} q[c] * W2k[c, lane*4 .. +3]
    __syncwarp();
    st4(qs + lane * 4, ldg4(a.q + (size_t)(a.q_by_slot ? slot : node) * a.ldq + lane * 4));
    __syncwarp();
    float4 U[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      U[h] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c4 = 0; c4 < DH / 4; ++c4) {
        float4 qv = ld4(qs + h * DH + c4 * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          U[h] = fma4(comp(qv, c), ld4(sm.W2 + (size_t)(h * DH + c4 * 4 + c) * H + lane * 4), U[h]);
      }
    }
    const float4 xi = ldg4(a.x4 + (size_t)node * 4);
    const float4 hi = ldg4(a.Hi + (size_t)(a.hi_by_slot ? slot : node) * a.ldhi + lane * 4);
    float* wrow = a.wbuf + (size_t)node * a.ldn * NH;
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const int seg_start = seg ? nlig : 0, seg_end = seg ? deg : nlig;
      const int type = lig_dst ? (seg ? 2 : 0) : (seg ? 3 : 1);   // uni_transformer_edge.py:371-377
#pragma unroll 1
      for (int e0 = seg_start; e0 < seg_end; e0 += 4) {
        float4 z[4];
        float rel[4][3];
        knn_hidden4(a, sm, gbuf, gam, bet, lane, node, xi, hi, e0, seg_end, type, z, rel);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float part[32];
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int h = 0; h < NH; ++h) part[s2 * NH + h] = dot4(U[h], z[p * 2 + s2]);
          warp_reduce_scatter<32>(part, lane);
          int e = e0 + p * 2 + (lane >> 4);
          if (e < seg_end) wrow[e * NH + (lane & 15)] = part[0];
        }
      }
    }
    __syncwarp();
    // softmax over the node's edges, per head (scatter_softmax at :64 / :205), then * e_w (:56 / :199)
    const int h = lane & 15, half = lane >> 4;
    float m = -INFINITY;
    for (int e = half; e < deg; e += 2) m = fmaxf(m, __ldcg(wrow + e * NH + h));
    m = fmaxf(m, __shfl_xor_sync(FULL, m, 16));
    float ssum = 0.f;
    for (int e = half; e < deg; e += 2) ssum += expf(__ldcg(wrow + e * NH + h) - m);
    ssum += __shfl_xor_sync(FULL, ssum, 16);
    for (int e = half; e < deg; e += 2) {
      float w = expf(__ldcg(wrow + e * NH + h) - m) / ssum;
      wrow[e * NH + h] = w * __ldg(a.e_w + (size_t)node * a.ldn + e);
    }
  }
}

// ----------------------------------------------------------------------------------- v pass (nodes)
__global__ void __launch_bounds__(ATT_THREADS, 1) knn_attn_v_node_kernel(const KnnAttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  KnnSmem sm(smem, true);
  load_knn_weights(a, sm, true);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* gbuf = sm.warp_scratch + warp * KnnSmem::kWarpFloats + H;
  const float4 gam = ld4(sm.gamma + lane * 4), bet = ld4(sm.beta + lane * 4);

  for (int slot = blockIdx.x * ATT_WARPS + warp; slot < a.n_dst; slot += gridDim.x * ATT_WARPS) {
    const int node = a.dst_list ? a.dst_list[slot] : slot;
    const int deg = a.deg[node];
    const int nlig = a.nlig[node];
    const bool lig_dst = a.is_lig[node];
    float4 S[NH];
    float wsum[NH];
#pragma unroll
    for (int h = 0; h < NH; ++h) { S[h] = make_float4(0.f, 0.f, 0.f, 0.f); wsum[h] = 0.f; }
    const float4 xi = ldg4(a.x4 + (size_t)node * 4);
    const float4 hi = ldg4(a.Hi + (size_t)(a.hi_by_slot ? slot : node) * a.ldhi + lane * 4);
    const float* wrow = a.wbuf + (size_t)node * a.ldn * NH;
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const int seg_start = seg ? nlig : 0, seg_end = seg ? deg : nlig;
      const int type = lig_dst ? (seg ? 2 : 0) : (seg ? 3 : 1);
#pragma unroll 1
      for (int e0 = seg_start; e0 < seg_end; e0 += 4) {
        float4 z[4];
        float rel[4][3];
        knn_hidden4(a, sm, gbuf, gam, bet, lane, node, xi, hi, e0, seg_end, type, z, rel);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (e0 + s < seg_end) {
#pragma unroll
            for (int h4 = 0; h4 < NH / 4; ++h4) {
              float4 w = ldg4(wrow + (e0 + s) * NH + h4 * 4);
              S[h4 * 4 + 0] = fma4(w.x, z[s], S[h4 * 4 + 0]); wsum[h4 * 4 + 0] += w.x;
              S[h4 * 4 + 1] = fma4(w.y, z[s], S[h4 * 4 + 1]); wsum[h4 * 4 + 1] += w.y;
              S[h4 * 4 + 2] = fma4(w.z, z[s], S[h4 * 4 + 2]); wsum[h4 * 4 + 2] += w.z;
              S[h4 * 4 + 3] = fma4(w.w, z[s], S[h4 * 4 + 3]); wsum[h4 * 4 + 3] += w.w;
            }
          }
        }
      }
    }
    // out[c] = <W2v[c,:], S[c/8]> + b2v[c] * wsum[c/8]; 4 chunks of 32 outputs, lane gets c = chunk*32 + lane
#pragma unroll
    for (int chunk = 0; chunk < 4; ++chunk) {
      float part[32];
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        int c = chunk * 32 + o;
        part[o] = dot4(ld4(sm.W2 + (size_t)c * H + lane * 4), S[c / DH]);
      }
      warp_reduce_scatter<32>(part, lane);
      int hq = lane >> 3;
      float ws = hq == 0 ? wsum[chunk * 4] : hq == 1 ? wsum[chunk * 4 + 1] : hq == 2 ? wsum[chunk * 4 + 2] : wsum[chunk * 4 + 3];
      int c = chunk * 32 + lane;
      a.out_h[(size_t)node * a.ldo + c] = part[0] + __ldg(a.w.b2 + c) * ws;
    }
  }
}

// ------------------------------------------------------------------------------- v pass (positions)
__global__ void __launch_bounds__(ATT_THREADS, 1) knn_attn_v_pos_kernel(const KnnAttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  KnnSmem sm(smem, false);
  load_knn_weights(a, sm, false);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* gbuf = sm.warp_scratch + warp * KnnSmem::kWarpFloats + H;
  const float4 gam = ld4(sm.gamma + lane * 4), bet = ld4(sm.beta + lane * 4);
  float4 Wv[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) Wv[h] = ldg4(a.w.W2 + (size_t)h * H + lane * 4);
  const float b2 = __ldg(a.w.b2 + (lane & 15));

  for (int slot = blockIdx.x * ATT_WARPS + warp; slot < a.n_dst; slot += gridDim.x * ATT_WARPS) {
    const int node = a.dst_list ? a.dst_list[slot] : slot;
    const int deg = a.deg[node];
    const int nlig = a.nlig[node];
    const bool lig_dst = a.is_lig[node];
    float acc[3] = {0.f, 0.f, 0.f};
    const float4 xi = ldg4(a.x4 + (size_t)node * 4);
    const float4 hi = ldg4(a.Hi + (size_t)(a.hi_by_slot ? slot : node) * a.ldhi + lane * 4);
    const float* wrow = a.wbuf + (size_t)node * a.ldn * NH;
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const int seg_start = seg ? nlig : 0, seg_end = seg ? deg : nlig;
      const int type = lig_dst ? (seg ? 2 : 0) : (seg ? 3 : 1);
#pragma unroll 1
      for (int e0 = seg_start; e0 < seg_end; e0 += 4) {
        float4 z[4];
        float rel[4][3];
        knn_hidden4(a, sm, gbuf, gam, bet, lane, node, xi, hi, e0, seg_end, type, z, rel);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          float part[32];
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int h = 0; h < NH; ++h) part[s2 * NH + h] = dot4(Wv[h], z[p * 2 + s2]);
          warp_reduce_scatter<32>(part, lane);
          const int s2 = lane >> 4;
          const int e = e0 + p * 2 + s2;
          if (e < seg_end) {
            float c = __ldg(wrow + e * NH + (lane & 15)) * (part[0] + b2);   // alpha * e_w * v  (:199-208)
            acc[0] = fmaf(c, s2 ? rel[p * 2 + 1][0] : rel[p * 2][0], acc[0]);
            acc[1] = fmaf(c, s2 ? rel[p * 2 + 1][1] : rel[p * 2][1], acc[1]);
            acc[2] = fmaf(c, s2 ? rel[p * 2 + 1][2] : rel[p * 2][2], acc[2]);
          }
        }
      }
    }
    acc[0] = warp_sum(acc[0]); acc[1] = warp_sum(acc[1]); acc[2] = warp_sum(acc[2]);
    if (lane == 0)
      st4(a.out_dx + (size_t)slot * 4, make_float4(acc[0] * (1.f / NH), acc[1] * (1.f / NH), acc[2] * (1.f / NH), 0.f));
  }
}

static int knn_grid(const KnnAttnArgs& a, int num_sms) {
  int need = (a.n_dst + ATT_WARPS - 1) / ATT_WARPS;
  return need < num_sms ? (need > 0 ? need : 1) : num_sms;
}

void launch_knn_attn_k(const KnnAttnArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_dst <= 0) return;
  static DeviceOnce once;
  int bytes = KnnSmem::bytes(true);
  if (!once.done()) { cudaFuncSetAttribute(knn_attn_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  knn_attn_k_kernel<<<knn_grid(a, num_sms), ATT_THREADS, bytes, stream>>>(a);
}
void launch_knn_attn_v_node(const KnnAttnArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_dst <= 0) return;
  static DeviceOnce once;
  int bytes = KnnSmem::bytes(true);
  if (!once.done()) { cudaFuncSetAttribute(knn_attn_v_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  knn_attn_v_node_kernel<<<knn_grid(a, num_sms), ATT_THREADS, bytes, stream>>>(a);
}
void launch_knn_attn_v_pos(const KnnAttnArgs& a, int num_sms, cudaStream_t stream) {
  if (a.n_dst <= 0) return;
  static DeviceOnce once;
  int bytes = KnnSmem::bytes(false);
  if (!once.done()) { cudaFuncSetAttribute(knn_attn_v_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); once.mark(); }
  knn_attn_v_pos_kernel<<<knn_grid(a, num_sms), ATT_THREADS, bytes, stream>>>(a);
}

}  // namespace ddb
