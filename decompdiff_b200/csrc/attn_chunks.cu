// Softmax groups of more than 32 rows (ligands of 34..65 atoms: bond edges entering an atom, triplets of a bond edge) are
// processed by the tensor-core kernels as two chunks of <= 32 rows.  The kernels here stitch the chunks together:
//   chunk_factor_kernel : per (chunk, head) the factor s_c e^{m_c - M} / sum_c' s_c' e^{m_c' - M} that rescales weights normalised
//                         within a chunk to the softmax over the whole group (m_c, s_c = the chunk's max logit and sum of exp)
//   *_combine_kernel    : adds the per-chunk partial results of the value passes in a fixed order (bit-reproducible) and applies
//                         what the single-chunk path does in its epilogue (residual / bias / masked position update)
#include "kernels.cuh"

namespace ddb {

__global__ void __launch_bounds__(256) chunk_factor_kernel(const float2* __restrict__ stats, const int* __restrict__ pair, int pair_stride,
                                                           int n_vg, float* __restrict__ factor) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = idx / NH, h = idx % NH;
  if (v >= n_vg) return;
  const int p = pair[(size_t)v * pair_stride];
  if (p < 0) return;                                   // single-chunk group: its weights are final
  const float2 a = stats[(size_t)v * NH + h], b = stats[(size_t)p * NH + h];
  const float M = fmaxf(a.x, b.x);
  float f = 0.f;
  if (M > -INFINITY) {
    const float ea = a.y > 0.f ? a.y * expf(a.x - M) : 0.f, eb = b.y > 0.f ? b.y * expf(b.x - M) : 0.f;
    f = ea / (ea + eb);
  }
  factor[(size_t)v * NH + h] = f;
}

void launch_chunk_factors(const float2* stats, const int* pair, int pair_stride, int n_vg, float* factor, cudaStream_t stream) {
  if (n_vg <= 0) return;
  chunk_factor_kernel<<<(n_vg * NH + 255) / 256, 256, 0, stream>>>(stats, pair, pair_stride, n_vg, factor);
}

// one thread per (first chunk of a chunked atom, channel)
__global__ void __launch_bounds__(128) bond_combine_node_kernel(const BondAttnArgs a) {
  const int v = blockIdx.x, c = threadIdx.x;
  const int4 g = a.vg[v];
  if (g.w < 0 || g.w < v) return;                      // single chunk, or the second chunk of a pair
  float* dst = a.out_h + (size_t)a.lig_idx[g.x] * a.ldo + c;
  *dst = *dst + (a.part_h[(size_t)v * H + c] + a.part_h[(size_t)g.w * H + c]);
}
__global__ void __launch_bounds__(128) bond_combine_pos_kernel(const BondAttnArgs a) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= a.n_vg) return;
  const int4 g = a.vg[v];
  if (g.w < 0 || g.w < v) return;
  const int node = a.lig_idx[g.x];
  float4 xi = ldg4(a.x4 + (size_t)node * 4);
  const float4 d0 = ld4(a.part_dx + (size_t)v * 4), d1 = ld4(a.part_dx + (size_t)g.w * 4);
  const float4 de = a.dx_edge ? ld4(a.dx_edge + (size_t)g.x * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float mk = (a.upd_mask == nullptr || a.upd_mask[g.x]) ? 1.f : 0.f;
  xi.x += (de.x + (d0.x + d1.x)) * mk;                 // x + (dx_edge + dx_bond) * mask   (uni_transformer_edge.py:284-285)
  xi.y += (de.y + (d0.y + d1.y)) * mk;
  xi.z += (de.z + (d0.z + d1.z)) * mk;
  st4(a.x4_out + (size_t)node * 4, xi);
}
void launch_bond_combine(const BondAttnArgs& a, bool pos, cudaStream_t stream) {
  if (a.n_vg <= 0) return;
  if (pos) bond_combine_pos_kernel<<<(a.n_vg + 127) / 128, 128, 0, stream>>>(a);
  else bond_combine_node_kernel<<<a.n_vg, 128, 0, stream>>>(a);
}

// one CTA per visiting position; the first chunk of a chunked group finishes the bond edge: h_bond_out = h_bond_in + sum + b2
__global__ void __launch_bounds__(128) trip_combine_kernel(const TripArgs a, const float* __restrict__ b2) {
  const int pos = blockIdx.x, c = threadIdx.x;
  const int p = a.vg_pair[pos];
  if (p < 0 || p < pos) return;
  const int e = a.grp_order[pos];
  // a chunk without a valid row has sum of exp 0 in every head
  const bool any = a.stats[(size_t)pos * NH].y > 0.f || a.stats[(size_t)p * NH].y > 0.f;
  const float upd = any ? (a.part[(size_t)pos * H + c] + a.part[(size_t)p * H + c]) + b2[c] : 0.f;
  a.h_bond_out[(size_t)e * H + c] = a.h_bond_in[(size_t)e * H + c] + upd;          // :274
}
void launch_trip_combine(const TripArgs& a, const float* b2, cudaStream_t stream) {
  if (a.n_groups <= 0) return;
  trip_combine_kernel<<<a.n_groups, 128, 0, stream>>>(a, b2);
}

}  // namespace ddb
