// Tensor-core attention over the kNN(32) edges (see attn_tc.cuh for the structure shared with the triplet kernels).
#include "attn_tc.cuh"

namespace ddb {

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
static_assert(KnnTcSmem::bytes() <= 232448, "shared memory budget");

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
