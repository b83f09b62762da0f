// K1: exact kNN graph per complex + the global edge weight.
//
// Replaces torch_geometric.nn.knn_graph (-> torch_cluster.knn) at
// /root/reference/models/encoders/uni_transformer_edge.py:353 and the edge_pred_layer block at :422-427.
//
// A complex has a few hundred atoms, so its coordinates (16 B each) sit in L1/L2 and a brute-force
// sweep per query is the right candidate generator at these sizes: one warp per query node, every lane
// keeps its strided share of candidate keys in registers, 32 rounds of warp-arg-min pick the k nearest.
// Keys are (fp32 bits of d^2) << 32 | index, with d^2 = ((dx*dx)+(dy*dy))+(dz*dz) evaluated without
// FMA contraction, so membership and order are bit-identical to the oracle (ties -> lower index).
#include <cstdlib>

#include "kernels.cuh"

namespace ddb {

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    unsigned long long o = __shfl_xor_sync(FULL, v, m);
    v = o < v ? o : v;
  }
  return v;
}

__device__ __forceinline__ unsigned long long knn_key(float4 xi, const float* __restrict__ x4, int c, int self) {
  float4 xj = ldg4(x4 + (size_t)c * 4);
  float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
  float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)c;
  return c == self ? ~0ull : key;
}

// MAXT > 0: candidate keys cached in registers (graphs up to 32*MAXT nodes); MAXT == 0: recompute per round.
template <int MAXT>
__global__ void __launch_bounds__(256) knn_kernel(const float* __restrict__ x4, const int* __restrict__ node_ptr,
                                                  const int* __restrict__ graph_of, const uint8_t* __restrict__ is_lig,
                                                  int n, int k, int* __restrict__ nbr, int* __restrict__ deg_out,
                                                  int* __restrict__ nlig_out) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const int g = graph_of[i];
  const int s = node_ptr[g], e = node_ptr[g + 1];
  const float4 xi = ldg4(x4 + (size_t)i * 4);
  const int deg = min(k, e - s - 1);

  unsigned long long keys[MAXT > 0 ? MAXT : 1];
  if (MAXT > 0) {
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      int c = s + lane + 32 * t;
      keys[t] = c < e ? knn_key(xi, x4, c, i) : ~0ull;
    }
  }
  unsigned long long last = 0ull;     // keys are unique, so "> last" walks them in order
  bool first = true;
  int mine = -1;                       // neighbour picked in round == lane
  for (int r = 0; r < deg; ++r) {
    unsigned long long best = ~0ull;
    if (MAXT > 0) {
#pragma unroll
      for (int t = 0; t < MAXT; ++t) {
        bool ok = first || keys[t] > last;
        best = (ok && keys[t] < best) ? keys[t] : best;
      }
    } else {
      for (int c = s + lane; c < e; c += 32) {
        unsigned long long kk = knn_key(xi, x4, c, i);
        bool ok = first || kk > last;
        best = (ok && kk < best) ? kk : best;
      }
    }
    best = warp_min_u64(best);
    last = best;
    first = false;
    if (lane == r) mine = (int)(best & 0xffffffffu);
  }
  // stable partition: ligand sources first (so the attention kernels see type-uniform runs)
  const bool has = lane < deg;
  const bool lig = has && is_lig[mine];
  const unsigned mlig = __ballot_sync(FULL, lig), mhas = __ballot_sync(FULL, has);
  const int nlig = __popc(mlig);
  const unsigned below = (1u << lane) - 1u;
  if (has) {
    int pos = lig ? __popc(mlig & below) : nlig + __popc(mhas & ~mlig & below);
    nbr[(size_t)i * KNN + pos] = mine;
  }
  if (lane == 0) { deg_out[i] = deg; nlig_out[i] = nlig; }
}

void launch_knn(const float* x4, const int* node_ptr, const int* graph_of, const uint8_t* is_lig, int n, int k,
                int max_graph_nodes, int* nbr, int* deg, int* nlig, cudaStream_t stream) {
  if (n <= 0) return;
  const int wpb = 8;
  dim3 grid((n + wpb - 1) / wpb), block(wpb * 32);
  if (max_graph_nodes <= 32 * 16) knn_kernel<16><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, nbr, deg, nlig);
  else if (max_graph_nodes <= 32 * 32) knn_kernel<32><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, nbr, deg, nlig);
  else if (max_graph_nodes <= 32 * 64) knn_kernel<64><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, nbr, deg, nlig);
  else knn_kernel<0><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, nbr, deg, nlig);
}

// e_w = sigmoid(MLP_{20->128->1}(gauss(d)))  (uni_transformer_edge.py:422-427), one warp per destination node,
// lane = 4 hidden channels, first-layer weights (20 x 128) held in registers.
//
// e_w depends on the distance only, and protein atoms never move during a sampling run: the value of every protein-protein
// pair is memoised in a per-graph table (NaN = not yet computed), filled on first use by the same arithmetic, so later steps
// only evaluate the MLP for edges that touch a ligand atom (a cache hit returns the bits a recomputation would produce).
__global__ void __launch_bounds__(256) edge_weight_kernel(const float* __restrict__ x4, const int* __restrict__ nbr,
                                                          const int* __restrict__ deg, int n,
                                                          const float* __restrict__ W1t /*[20][128]*/,
                                                          const float* __restrict__ b1, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ w2,
                                                          float b2, float* __restrict__ e_w, EdgeWeightCache c) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const float4 xi = ldg4(x4 + (size_t)i * 4);
  const int d_i = deg[i];
  // ---- lookup phase: lane = edge
  const int j = lane < d_i ? nbr[(size_t)i * KNN + lane] : i;
  const float4 xj = ldg4(x4 + (size_t)j * 4);
  float val = __int_as_float(0x7fc00000);
  long long slot = -1;
  if (c.table != nullptr && lane < d_i) {
    const int g = c.graph_of[i], base = c.node_ptr[g], np = c.n_protein[g];
    const int il = i - base, jl = j - base;
    if (il < np && jl < np) { slot = c.table_base[g] + (long long)il * np + jl; val = __ldcg(c.table + slot); }
  }
  unsigned todo = __ballot_sync(FULL, lane < d_i && val != val);
  if (todo) {
    // ---- evaluation phase: the warp computes the missing edges one at a time
    float4 w[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) w[g] = ldg4(W1t + g * H + lane * 4);
    const float4 bb = ldg4(b1 + lane * 4), gm = ldg4(gamma + lane * 4), bt = ldg4(beta + lane * 4), w2v = ldg4(w2 + lane * 4);
    while (todo) {
      const int e = __ffs(todo) - 1;
      todo &= todo - 1;
      const float xjx = __shfl_sync(FULL, xj.x, e), xjy = __shfl_sync(FULL, xj.y, e), xjz = __shfl_sync(FULL, xj.z, e);
      const float dx = xi.x - xjx, dy = xi.y - xjy, dz = xi.z - xjz;
      const float d = sqrtf(dx * dx + dy * dy + dz * dz);
      const float gl = lane < NG ? gauss_feat(d, lane) : 0.f;
      float4 z[1] = {bb};
#pragma unroll
      for (int g = 0; g < NG; ++g) z[0] = fma4(__shfl_sync(FULL, gl, g), w[g], z[0]);
      ln_relu_rows<1>(z, gm, bt, lane);
      const float logit = warp_sum(dot4(z[0], w2v)) + b2;
      const float r = 1.0f / (1.0f + expf(-logit));
      if (lane == e) { val = r; if (slot >= 0) c.table[slot] = r; }
    }
  }
  if (lane < d_i) e_w[(size_t)i * KNN + lane] = val;
}

void launch_edge_weight(const float* x4, const int* nbr, const int* deg, int n, const float* W1t, const float* b1,
                        const float* gamma, const float* beta, const float* w2, float b2, float* e_w,
                        const EdgeWeightCache& cache, cudaStream_t stream) {
  if (n <= 0) return;
  const int wpb = 8;
  edge_weight_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, stream>>>(x4, nbr, deg, n, W1t, b1, gamma, beta, w2, b2, e_w, cache);
}

// ---------------------------------------------------------------------------------------------------------------------
// Exact receptive field of the outputs.  Only ligand positions / features / bond features leave the network, so the node
// update of layer l (0-based, L layers) is needed only for nodes within L - l hops of a ligand atom along kNN edges
// (destination -> its sources).  level[node] = that hop count (ligand 0, capped at LEVEL_CAP); protein nodes are then listed
// by ascending level behind the (static) ligand block of `dst_list`, so every layer works on a PREFIX whose length is a
// device-side counter - nothing is approximated, rows that cannot influence the outputs are simply not computed.
constexpr int LEVEL_CAP = 7;
__global__ void __launch_bounds__(256) level_init_kernel(const uint8_t* __restrict__ is_lig, int n, int* __restrict__ level) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) level[i] = is_lig[i] ? 0 : LEVEL_CAP;
}
__global__ void __launch_bounds__(256) level_relax_kernel(const int* __restrict__ nbr, const int* __restrict__ deg, int n, int round,
                                                          int* __restrict__ level) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = idx >> 5, lane = idx & 31;
  if (i >= n || level[i] != round || lane >= deg[i]) return;
  atomicMin(level + nbr[idx], round + 1);
}
// Deterministic counting sort of the protein nodes by (level, graph, index): one warp per graph ranks its nodes with ballots, so
// the destination list - and with it the tile composition of the attention kernels - is the same on every run.
__global__ void __launch_bounds__(256) level_count_kernel(const int* __restrict__ level, const int* __restrict__ node_ptr,
                                                          const int* __restrict__ n_protein, int num_graphs, int* __restrict__ cnt) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= num_graphs) return;
  const int base = node_ptr[g], np = n_protein[g];
  int c[LEVEL_CAP + 1];
#pragma unroll
  for (int v = 0; v <= LEVEL_CAP; ++v) c[v] = 0;
  for (int i0 = 0; i0 < np; i0 += 32) {
    const int lv = i0 + lane < np ? level[base + i0 + lane] : -1;
#pragma unroll
    for (int v = 0; v <= LEVEL_CAP; ++v) c[v] += __popc(__ballot_sync(FULL, lv == v));
  }
  if (lane == 0) {
#pragma unroll
    for (int v = 0; v <= LEVEL_CAP; ++v) cnt[g * (LEVEL_CAP + 1) + v] = c[v];
  }
}
// counts[0 .. n_layers): destinations of layer l (ligand block + protein nodes with level <= n_layers - l);
// counts[n_layers .. 2 n_layers): source rows of layer l (level <= n_layers - l + 1); counts[2 n_layers]: level <= 1;
// cnt[(g, v)] is replaced by the first list position of graph g's level-v nodes
__global__ void level_offsets_kernel(int* __restrict__ cnt, int num_graphs, int lig_block, int n_layers, int* __restrict__ counts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int start[LEVEL_CAP + 2];
  start[0] = 0;
  for (int v = 0; v <= LEVEL_CAP; ++v) {
    int run = start[v];
    for (int g = 0; g < num_graphs; ++g) { const int c = cnt[g * (LEVEL_CAP + 1) + v]; cnt[g * (LEVEL_CAP + 1) + v] = lig_block + run; run += c; }
    start[v + 1] = run;
  }
  auto upto = [&](int lv) { return lig_block + start[min(max(lv, 0), LEVEL_CAP) + 1]; };
  for (int l = 0; l < n_layers; ++l) {
    counts[l] = upto(n_layers - l);
    counts[n_layers + l] = upto(n_layers - l + 1);
  }
  counts[2 * n_layers] = upto(1);
}
__global__ void __launch_bounds__(256) level_scatter_kernel(const int* __restrict__ level, const int* __restrict__ node_ptr,
                                                            const int* __restrict__ n_protein, int num_graphs, const int* __restrict__ first,
                                                            int* __restrict__ dst_list) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= num_graphs) return;
  const int base = node_ptr[g], np = n_protein[g];
  int cur[LEVEL_CAP + 1];
#pragma unroll
  for (int v = 0; v <= LEVEL_CAP; ++v) cur[v] = first[g * (LEVEL_CAP + 1) + v];
  for (int i0 = 0; i0 < np; i0 += 32) {
    const int lv = i0 + lane < np ? level[base + i0 + lane] : -1;
#pragma unroll
    for (int v = 0; v <= LEVEL_CAP; ++v) {
      const unsigned m = __ballot_sync(FULL, lv == v);
      if (lv == v) dst_list[cur[v] + __popc(m & ((1u << lane) - 1u))] = base + i0 + lane;
      cur[v] += __popc(m);
    }
  }
}

// counting sort of the protein nodes by `level` behind the ligand block + the per-layer prefix lengths
void launch_level_sort(const int* level, const int* node_ptr, const int* n_protein, int num_graphs, int n_layers, int lig_block,
                       int* cnt /* 8 * num_graphs ints */, int* counts /* 2 * n_layers + 1 */, int* dst_list, cudaStream_t stream) {
  if (num_graphs <= 0) return;
  const int gb = (num_graphs + 7) / 8;
  level_count_kernel<<<gb, 256, 0, stream>>>(level, node_ptr, n_protein, num_graphs, cnt);
  level_offsets_kernel<<<1, 32, 0, stream>>>(cnt, num_graphs, lig_block, n_layers, counts);
  level_scatter_kernel<<<gb, 256, 0, stream>>>(level, node_ptr, n_protein, num_graphs, cnt, dst_list);
}

void launch_receptive_field(const int* nbr, const int* deg, const uint8_t* is_lig, const int* node_ptr, const int* n_protein, int num_graphs,
                            int n, int n_layers, int lig_block, int* level, int* cnt /* 8 * num_graphs ints */,
                            int* counts /* 2 * n_layers + 1 */, int* dst_list, cudaStream_t stream) {
  if (n <= 0) return;
  const int nb = (n + 255) / 256;
  level_init_kernel<<<nb, 256, 0, stream>>>(is_lig, n, level);
  for (int r = 0; r < LEVEL_CAP - 1; ++r) level_relax_kernel<<<(n * 32 + 255) / 256, 256, 0, stream>>>(nbr, deg, n, r, level);
  launch_level_sort(level, node_ptr, n_protein, num_graphs, n_layers, lig_block, cnt, counts, dst_list, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// First-layer cache.  The layer-0 output of a protein node whose 32 sources are all protein atoms depends on nothing that
// changes during a run (static positions, static embeddings, static neighbour set, static edge weights), so it is computed once
// and kept in a buffer only layer 0 writes.  key0 = 1 for the protein nodes layer 0 still has to compute this step (inside the
// receptive field AND (a ligand atom among the sources OR not cached yet)), LEVEL_CAP for the rest; the sort above turns it into
// a destination list whose "level <= 1" prefix is exactly that set.  valid0 remembers which rows hold a static value.
__global__ void __launch_bounds__(256) layer0_key_kernel(const int* __restrict__ level, const int* __restrict__ nlig,
                                                         const uint8_t* __restrict__ is_lig, int n, int n_layers,
                                                         uint8_t* __restrict__ valid0, int* __restrict__ key0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (is_lig[i]) { key0[i] = 0; return; }
  const bool compute = level[i] <= n_layers && (nlig[i] > 0 || !valid0[i]);
  key0[i] = compute ? 1 : LEVEL_CAP;
  if (compute) valid0[i] = nlig[i] == 0;        // the row written this step is reusable iff it has no ligand source
}
void launch_layer0_keys(const int* level, const int* nlig, const uint8_t* is_lig, int n, int n_layers, uint8_t* valid0, int* key0,
                        cudaStream_t stream) {
  if (n <= 0) return;
  layer0_key_kernel<<<(n + 255) / 256, 256, 0, stream>>>(level, nlig, is_lig, n, n_layers, valid0, key0);
}

}  // namespace ddb
