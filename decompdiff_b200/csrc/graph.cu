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
                                                  int n, int k, float r2max, int* __restrict__ nbr, int* __restrict__ deg_out,
                                                  int* __restrict__ nlig_out, const int* __restrict__ node_list,
                                                  const int* __restrict__ n_protein, unsigned long long* __restrict__ skeys,
                                                  int ld, bool hybrid) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= n) return;
  const int i = node_list ? node_list[w] : w;      // optional list of query nodes (ligand atoms when the protein part is cached)
  const int g = graph_of[i];
  const int s = node_ptr[g];
  // static mode (skeys != null): candidates are the graph's PROTEIN atoms only, the sorted keys are the output
  // 'hybrid' graphs (common.py:230-277): a ligand destination gets every other ligand atom of its complex plus its k nearest
  // PROTEIN atoms; protein destinations keep the plain kNN list over all atoms
  const bool hyb_lig = hybrid && is_lig[i];
  const int e = (skeys || hyb_lig) ? s + n_protein[g] : node_ptr[g + 1];
  if (skeys && i >= e) return;                     // ligand node: no static list
  const float4 xi = ldg4(x4 + (size_t)i * 4);
  int deg = hyb_lig ? min(k, e - s) : min(k, e - s - 1);

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
    // 'radius' cut-off: neighbours come nearest first, so the first one beyond r_max ends the list (warp-uniform)
    if (__uint_as_float((unsigned)(best >> 32)) > r2max) { deg = r; break; }
    last = best;
    first = false;
    if (lane == r) mine = (int)(best & 0xffffffffu);
    if (skeys && lane == r) skeys[(size_t)i * KNN + r] = best;
  }
  if (skeys) { if (lane == 0) deg_out[i] = deg; return; }
  if (hyb_lig) {      // ligand sources first (index order, self skipped), then the protein sources nearest first
    const int l0 = e, l1 = node_ptr[g + 1], nl = l1 - l0 - 1;
    for (int c = l0 + lane; c < l1; c += 32)
      if (c != i) nbr[(size_t)i * ld + (c - l0) - (c > i ? 1 : 0)] = c;
    if (lane < deg) nbr[(size_t)i * ld + nl + lane] = mine;
    if (lane == 0) { deg_out[i] = nl + deg; nlig_out[i] = nl; }
    return;
  }
  // stable partition: ligand sources first (so the attention kernels see type-uniform runs)
  const bool has = lane < deg;
  const bool lig = has && is_lig[mine];
  const unsigned mlig = __ballot_sync(FULL, lig), mhas = __ballot_sync(FULL, has);
  const int nlig = __popc(mlig);
  const unsigned below = (1u << lane) - 1u;
  if (has) {
    int pos = lig ? __popc(mlig & below) : nlig + __popc(mhas & ~mlig & below);
    nbr[(size_t)i * ld + pos] = mine;
  }
  if (lane == 0) { deg_out[i] = deg; nlig_out[i] = nlig; }
}

void launch_knn(const float* x4, const int* node_ptr, const int* graph_of, const uint8_t* is_lig, int n, int k,
                int max_graph_nodes, int* nbr, int* deg, int* nlig, cudaStream_t stream, float r_max, const int* node_list,
                const int* n_protein, unsigned long long* skeys, int ld, bool hybrid) {
  const float r2max = r_max > 0.f ? r_max * r_max : INFINITY;
  if (n <= 0) return;
  const int wpb = 8;
  dim3 grid((n + wpb - 1) / wpb), block(wpb * 32);
  if (max_graph_nodes <= 32 * 16) knn_kernel<16><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, r2max, nbr, deg, nlig, node_list, n_protein, skeys, ld, hybrid);
  else if (max_graph_nodes <= 32 * 32) knn_kernel<32><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, r2max, nbr, deg, nlig, node_list, n_protein, skeys, ld, hybrid);
  else if (max_graph_nodes <= 32 * 64) knn_kernel<64><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, r2max, nbr, deg, nlig, node_list, n_protein, skeys, ld, hybrid);
  else knn_kernel<0><<<grid, block, 0, stream>>>(x4, node_ptr, graph_of, is_lig, n, k, r2max, nbr, deg, nlig, node_list, n_protein, skeys, ld, hybrid);
}

// Protein destinations with the cached static part (launch_knn with skeys, once per run: protein atoms never move): the k nearest
// PROTEIN neighbours are known and sorted, so per step only the graph's <= 64 ligand atoms have to be ranked against them.  Ranks
// are counted, not sorted: a key's rank in the union is the number of smaller keys; keys are unique ((d^2 bits, index) as in
// knn_kernel), so the selected set, its order and the tie rule are exactly those of the brute-force kernel.
__global__ void __launch_bounds__(256) knn_merge_kernel(const float* __restrict__ x4, const int* __restrict__ node_ptr,
                                                        const int* __restrict__ graph_of, const int* __restrict__ n_protein,
                                                        const uint8_t* __restrict__ is_lig, int n, int k, float r2max,
                                                        const unsigned long long* __restrict__ skeys, const int* __restrict__ sdeg,
                                                        int* __restrict__ nbr, int* __restrict__ deg_out, int* __restrict__ nlig_out) {
  __shared__ unsigned long long sh[8][96];                 // per warp: 32 static keys | up to 64 ligand keys
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int i = blockIdx.x * (blockDim.x >> 5) + wib;
  if (i >= n || is_lig[i]) return;
  const int g = graph_of[i];
  const int s = node_ptr[g], e = node_ptr[g + 1], l0 = s + n_protein[g], nl = e - l0;
  const float4 xi = ldg4(x4 + (size_t)i * 4);
  unsigned long long* my = sh[wib];
  const unsigned long long sk = lane < sdeg[i] ? skeys[(size_t)i * KNN + lane] : ~0ull;
  unsigned long long lk[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int j = l0 + c * 32 + lane;
    lk[c] = j < e ? knn_key(xi, x4, j, i) : ~0ull;
    if (__uint_as_float((unsigned)(lk[c] >> 32)) > r2max) lk[c] = ~0ull;
    my[32 + c * 32 + lane] = lk[c];
  }
  my[lane] = sk;
  __syncwarp();
  const int nlk = nl > 32 ? 64 : 32;
  int below_s = 0;                                         // ligand keys smaller than this lane's static key
  for (int j = 0; j < nlk; ++j) below_s += my[32 + j] < sk;
  int rank[2], lrank[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    int rs = 0, rl = 0;
    if (c * 32 < nlk) {
      for (int j = 0; j < 32; ++j) rs += my[j] < lk[c];
      for (int j = 0; j < nlk; ++j) rl += my[32 + j] < lk[c];
    }
    lrank[c] = rl; rank[c] = rs + rl;
  }
  const bool sel_s = sk != ~0ull && lane + below_s < k;
  const bool sel0 = lk[0] != ~0ull && rank[0] < k, sel1 = lk[1] != ~0ull && rank[1] < k;
  const int nlig = __popc(__ballot_sync(FULL, sel0)) + __popc(__ballot_sync(FULL, sel1));
  const int nsta = __popc(__ballot_sync(FULL, sel_s));
  if (sel0) nbr[(size_t)i * KNN + lrank[0]] = (int)(lk[0] & 0xffffffffu);          // ligand sources first, nearest first
  if (sel1) nbr[(size_t)i * KNN + lrank[1]] = (int)(lk[1] & 0xffffffffu);
  if (sel_s) nbr[(size_t)i * KNN + nlig + lane] = (int)(sk & 0xffffffffu);          // then the protein sources, nearest first
  if (lane == 0) { deg_out[i] = nlig + nsta; nlig_out[i] = nlig; }
}

void launch_knn_merge(const float* x4, const int* node_ptr, const int* graph_of, const int* n_protein, const uint8_t* is_lig, int n, int k,
                      float r_max, const unsigned long long* skeys, const int* sdeg, int* nbr, int* deg, int* nlig, cudaStream_t stream) {
  if (n <= 0) return;
  const float r2max = r_max > 0.f ? r_max * r_max : INFINITY;
  knn_merge_kernel<<<(n + 7) / 8, 256, 0, stream>>>(x4, node_ptr, graph_of, n_protein, is_lig, n, k, r2max, skeys, sdeg, nbr, deg, nlig);
}

// e_w = sigmoid(MLP_{20->128->1}(gauss(d)))  (uni_transformer_edge.py:422-427), one warp per destination node,
// lane = 4 hidden channels, first-layer weights (20 x 128) in shared memory.
//
// e_w depends on the distance only, and protein atoms never move during a sampling run: the value of every protein-protein
// pair is memoised in a per-graph table (NaN = not yet computed), filled on first use by the same arithmetic, so later steps
// only evaluate the MLP for edges that touch a ligand atom (a cache hit returns the bits a recomputation would produce).
__global__ void __launch_bounds__(256, 4) edge_weight_kernel(const float* __restrict__ x4, const int* __restrict__ nbr,
                                                          const int* __restrict__ deg, int n,
                                                          const float* __restrict__ W1t /*[20][128]*/,
                                                          const float* __restrict__ b1, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ w2,
                                                          float b2, float* __restrict__ e_w, EdgeWeightCache c, int ld) {
  // first-layer weights in shared memory (10 KB per CTA): in registers (80 per thread) they capped the kernel at one CTA per SM,
  // and the kernel is a chain of dependent L2 reads (neighbour row -> positions -> memo table) that only occupancy hides
  __shared__ float4 sW[NG * H / 4];
  for (int t = threadIdx.x; t < NG * H / 4; t += blockDim.x) sW[t] = ldg4(W1t + t * 4);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n) return;
  const float4 xi = ldg4(x4 + (size_t)i * 4);
  const int d_all = deg[i];
  for (int e0 = 0; e0 < d_all; e0 += 32) {      // 32 edges at a time ('hybrid' graphs have longer lists)
    const int d_i = min(d_all - e0, 32);
    // ---- lookup phase: lane = edge
    const int j = lane < d_i ? nbr[(size_t)i * ld + e0 + lane] : i;
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
      const float4 bb = ldg4(b1 + lane * 4), gm = ldg4(gamma + lane * 4), bt = ldg4(beta + lane * 4), w2v = ldg4(w2 + lane * 4);
      // four missing edges per pass: the same per-edge arithmetic as one at a time (every reduction is a plain warp_sum of that
      // edge's values, so the bits do not depend on which edges share a pass), with four independent dependency chains in flight
      while (todo) {
        int e[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { e[u] = todo ? __ffs(todo) - 1 : -1; todo &= todo - 1; }
        float gl[4];
        float4 z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int eu = max(e[u], 0);
          const float xjx = __shfl_sync(FULL, xj.x, eu), xjy = __shfl_sync(FULL, xj.y, eu), xjz = __shfl_sync(FULL, xj.z, eu);
          const float dx = xi.x - xjx, dy = xi.y - xjy, dz = xi.z - xjz;
          const float d = sqrtf(dx * dx + dy * dy + dz * dz);
          gl[u] = lane < NG ? gauss_feat(d, lane) : 0.f;
          z[u] = bb;
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float4 wg = sW[g * (H / 4) + lane];
#pragma unroll
          for (int u = 0; u < 4; ++u) z[u] = fma4(__shfl_sync(FULL, gl[u], g), wg, z[u]);
        }
        float sm[4], sq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) sm[u] = warp_sum((z[u].x + z[u].y) + (z[u].z + z[u].w));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float mu = sm[u] * (1.0f / H);
          z[u].x -= mu; z[u].y -= mu; z[u].z -= mu; z[u].w -= mu;
          sq[u] = warp_sum((z[u].x * z[u].x + z[u].y * z[u].y) + (z[u].z * z[u].z + z[u].w * z[u].w));
        }
        float lg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float rstd = 1.0f / sqrtf(sq[u] * (1.0f / H) + LN_EPS);
          z[u].x = fmaxf(fmaf(z[u].x * rstd, gm.x, bt.x), 0.f);
          z[u].y = fmaxf(fmaf(z[u].y * rstd, gm.y, bt.y), 0.f);
          z[u].z = fmaxf(fmaf(z[u].z * rstd, gm.z, bt.z), 0.f);
          z[u].w = fmaxf(fmaf(z[u].w * rstd, gm.w, bt.w), 0.f);
          lg[u] = warp_sum(dot4(z[u], w2v)) + b2;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float r = 1.0f / (1.0f + expf(-lg[u]));
          if (e[u] >= 0 && lane == e[u]) { val = r; if (slot >= 0) c.table[slot] = r; }
        }
      }
    }
    if (lane < d_i) e_w[(size_t)i * ld + e0 + lane] = val;
  }
}

void launch_edge_weight(const float* x4, const int* nbr, const int* deg, int n, const float* W1t, const float* b1,
                        const float* gamma, const float* beta, const float* w2, float b2, float* e_w,
                        const EdgeWeightCache& cache, cudaStream_t stream, int ld) {
  if (n <= 0) return;
  const int wpb = 8;
  edge_weight_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, stream>>>(x4, nbr, deg, n, W1t, b1, gamma, beta, w2, b2, e_w, cache, ld);
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

// ---------------------------------------------------------------------------------------------------------------------
// The same lists in two launches (graphs of up to GL_MAX_NODES nodes): one CTA per graph does the hop levels in shared memory, the
// first-layer keys and both histograms; a second CTA per graph turns the histograms of all graphs into list positions and scatters
// its nodes (same deterministic order: level, graph, index) together with the per-slot metadata of the attention kernels.
constexpr int GL_MAX_NODES = 6144;
__global__ void __launch_bounds__(1024) graph_levels_kernel(const int* __restrict__ nbr, const int* __restrict__ deg, const int* __restrict__ nlig,
                                                           const uint8_t* __restrict__ is_lig, const int* __restrict__ node_ptr,
                                                           const int* __restrict__ n_protein, int n_layers, int use_l0,
                                                           uint8_t* __restrict__ valid0, int* __restrict__ level, int* __restrict__ key0,
                                                           int* __restrict__ cnt, int* __restrict__ cnt0) {
  __shared__ int lv[GL_MAX_NODES];
  __shared__ int hist[2][LEVEL_CAP + 1];
  const int g = blockIdx.x, base = node_ptr[g], ng = node_ptr[g + 1] - base, np = n_protein[g];
  if (threadIdx.x < 2 * (LEVEL_CAP + 1)) (&hist[0][0])[threadIdx.x] = 0;
  for (int t = threadIdx.x; t < ng; t += blockDim.x) lv[t] = is_lig[base + t] ? 0 : LEVEL_CAP;
  __syncthreads();
  for (int round = 0; round < LEVEL_CAP - 1; ++round) {
    for (int idx = threadIdx.x; idx < ng * 32; idx += blockDim.x) {
      const int t = idx >> 5, j = idx & 31;
      if (lv[t] == round && j < deg[base + t]) atomicMin(&lv[nbr[(size_t)(base + t) * KNN + j] - base], round + 1);
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < ng; t += blockDim.x) {
    const int i = base + t, l = lv[t];
    level[i] = l;
    if (t < np) {
      atomicAdd(&hist[0][l], 1);
      if (use_l0) {
        const bool compute = l <= n_layers && (nlig[i] > 0 || !valid0[i]);
        const int k0 = compute ? 1 : LEVEL_CAP;
        key0[i] = k0;
        if (compute) valid0[i] = nlig[i] == 0;
        atomicAdd(&hist[1][k0], 1);
      }
    } else if (use_l0) {
      key0[i] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x <= LEVEL_CAP) {
    cnt[g * (LEVEL_CAP + 1) + threadIdx.x] = hist[0][threadIdx.x];
    if (use_l0) cnt0[g * (LEVEL_CAP + 1) + threadIdx.x] = hist[1][threadIdx.x];
  }
}

struct GraphListArgs {
  const int *level, *key0, *cnt, *cnt0, *node_ptr, *n_protein, *lig_ptr, *lig_idx, *deg, *nlig;
  int num_graphs, n_layers, lig_block, use_l0;
  int *counts, *counts0, *dst_lvl, *dst_lvl0;
  int2 *meta_lvl, *meta_lvl0, *meta_lig;
};
__global__ void __launch_bounds__(64) graph_lists_kernel(const GraphListArgs a) {
  __shared__ int first[2][LEVEL_CAP + 1];       // list position of this graph's first node of every level / key
  __shared__ int total[2][LEVEL_CAP + 2];       // exclusive prefix over levels of the totals
  const int g = blockIdx.x, lane = threadIdx.x & 31, which = threadIdx.x >> 5;      // warp 0: level list, warp 1: first-layer list
  if (which == 1 && !a.use_l0) return;
  const int* c = which ? a.cnt0 : a.cnt;
  if (lane <= LEVEL_CAP) {
    int before = 0, all = 0;
    for (int gg = 0; gg < a.num_graphs; ++gg) { const int v = c[gg * (LEVEL_CAP + 1) + lane]; all += v; before += gg < g ? v : 0; }
    first[which][lane] = before; total[which][lane + 1] = all;
  }
  __syncwarp();
  if (lane == 0) { total[which][0] = 0; for (int v = 0; v <= LEVEL_CAP; ++v) total[which][v + 1] += total[which][v]; }
  __syncwarp();
  if (lane <= LEVEL_CAP) first[which][lane] += a.lig_block + total[which][lane];
  __syncwarp();
  if (g == 0 && lane == 0) {      // per-layer prefix lengths (see level_offsets_kernel)
    int* counts = which ? a.counts0 : a.counts;
    auto upto = [&](int lv) { return a.lig_block + total[which][min(max(lv, 0), LEVEL_CAP) + 1]; };
    for (int l = 0; l < a.n_layers; ++l) { counts[l] = upto(a.n_layers - l); counts[a.n_layers + l] = upto(a.n_layers - l + 1); }
    counts[2 * a.n_layers] = upto(1);
  }
  const int* key = which ? a.key0 : a.level;
  int* dst = which ? a.dst_lvl0 : a.dst_lvl;
  int2* meta = which ? a.meta_lvl0 : a.meta_lvl;
  const int base = a.node_ptr[g], np = a.n_protein[g];
  int cur[LEVEL_CAP + 1];
#pragma unroll
  for (int v = 0; v <= LEVEL_CAP; ++v) cur[v] = first[which][v];
  for (int i0 = 0; i0 < np; i0 += 32) {
    const int node = base + i0 + lane;
    const int lv = i0 + lane < np ? key[node] : -1;
#pragma unroll
    for (int v = 0; v <= LEVEL_CAP; ++v) {
      const unsigned m = __ballot_sync(FULL, lv == v);
      if (lv == v) {
        const int pos = cur[v] + __popc(m & ((1u << lane) - 1u));
        dst[pos] = node;
        meta[pos] = make_int2(node, a.deg[node] | (a.nlig[node] << 8));
      }
      cur[v] += __popc(m);
    }
  }
  // the ligand block of the list (static node order) and the ligand-only list of the position layers
  for (int r = a.lig_ptr[g] + lane; r < a.lig_ptr[g + 1]; r += 32) {
    const int node = a.lig_idx[r];
    const int2 mm = make_int2(node, a.deg[node] | (a.nlig[node] << 8) | (1 << 16));
    meta[r] = mm;
    if (which == 0) a.meta_lig[r] = mm;
  }
}

bool launch_graph_lists(const int* nbr, const int* deg, const int* nlig, const uint8_t* is_lig, const int* node_ptr, const int* n_protein,
                        const int* lig_ptr, const int* lig_idx, int num_graphs, int max_graph_nodes, int n_layers, int lig_block, bool use_l0,
                        uint8_t* valid0, int* level, int* key0, int* cnt, int* cnt0, int* counts, int* counts0, int* dst_lvl, int* dst_lvl0,
                        int2* meta_lvl, int2* meta_lvl0, int2* meta_lig, cudaStream_t stream) {
  if (num_graphs <= 0 || max_graph_nodes > GL_MAX_NODES) return false;
  graph_levels_kernel<<<num_graphs, 1024, 0, stream>>>(nbr, deg, nlig, is_lig, node_ptr, n_protein, n_layers, use_l0 ? 1 : 0, valid0, level, key0,
                                                      cnt, cnt0);
  GraphListArgs a;
  a.level = level; a.key0 = key0; a.cnt = cnt; a.cnt0 = cnt0; a.node_ptr = node_ptr; a.n_protein = n_protein; a.lig_ptr = lig_ptr;
  a.lig_idx = lig_idx; a.deg = deg; a.nlig = nlig; a.num_graphs = num_graphs; a.n_layers = n_layers; a.lig_block = lig_block;
  a.use_l0 = use_l0 ? 1 : 0; a.counts = counts; a.counts0 = counts0; a.dst_lvl = dst_lvl; a.dst_lvl0 = dst_lvl0;
  a.meta_lvl = meta_lvl; a.meta_lvl0 = meta_lvl0; a.meta_lig = meta_lig;
  graph_lists_kernel<<<num_graphs, 64, 0, stream>>>(a);
  return true;
}

}  // namespace ddb
