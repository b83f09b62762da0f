// Tensor-core attention over the kNN(32) edges: NodeUpdateLayer / PosUpdateLayer key pass, the node value pass and the position
// value pass
// (uni_transformer_edge.py:42-74, 188-210).  Structure shared with the triplet kernel (attn_tc.cuh, attn_tc_trip.cu):
//   rows   : one kNN edge each; 32 rows = the incoming edges of one destination node = one softmax group = one TMEM lane
//            quadrant; 4 destinations form a 128-row tile; thread = (row, 32-channel slice)
//   GEMM 1 : the distance term of the first Linear, sum_g gauss_g(d) Wg[type][g][:], on tcgen05: A2 = Gaussian features of the
//            row placed in the column block of its source class (protein | ligand source, 2 x 20 columns, hi/lo TF32 split,
//            SWIZZLE_32B K-major tile written by the workers), B2 = the two matching type blocks of Wg for the destination
//            class of the tile (destinations are visited one class after the other; B2 is swapped once at the class boundary);
//            tiles without a ligand source need only 3 of the 5 k-steps
//   SIMT   : z = D2 + P_src[j] (row gather, LDG.256) + (P_dst[i] + Wt[type]) (staged per warp), LayerNorm, ReLU, TF32 split
//   GEMM 2 : second Linear, A = hidden activations in TMEM, B = W2 hi/lo resident in shared memory (3xTF32)
//   k pass : logits = <q_i[head], D[row, head]>, softmax over the 32 rows, times e_w -> wbuf
//   v pass : out_h[i] = sum_rows w[row, head(c)] D[row, c] + b2[c] sum_rows w[row, head(c)]
//   pos v  : 16-column second Linear (one scalar per head); dx_i = mean_heads sum_rows w[row, h] (D[row, h] + b2[h]) (x_i - x_j)
// A dedicated warp issues the MMAs (the distance MMA one tile ahead); setmaxnreg moves its registers to the workers.  The
// destination list may be a device-counted prefix (exact receptive-field pruning / first-layer cache, graph.cu).
#include "attn_tc.cuh"

namespace ddb {

constexpr int KT_THREADS = ATC_THREADS + 128;     // 16 worker warps + the issuing warpgroup
constexpr int KT_SYNC = ATC_THREADS + 32;         // participants of the hand-over barriers
constexpr int KT_KB = 5;                          // 40 feature columns = 5 k-steps of 8
constexpr int KT_IMG = KT_KB * 128 * 32;          // one TF32 image of a [128 rows][40 cols] SWIZZLE_32B operand: 20480 bytes
constexpr int KT_COL_D2 = 384;
constexpr int KBAR_A_READY = 6;

struct KnnTcSmem {
  uint8_t *W2, *B2, *A2; float *gamma, *beta, *b2, *hit, *qry; float2* stat; uint64_t* bars; uint32_t* tmem_slot;
  __device__ explicit KnnTcSmem(uint8_t* raw) {
    uint8_t* p = raw;
    W2 = p; p += ATC_W2_BYTES;
    B2 = p; p += 2 * KT_IMG;
    A2 = p; p += 2 * KT_IMG;
    gamma = reinterpret_cast<float*>(p); p += H * 4;
    beta = reinterpret_cast<float*>(p); p += H * 4;
    b2 = reinterpret_cast<float*>(p); p += H * 4;
    hit = reinterpret_cast<float*>(p); p += 16 * 64 * 4;        // per warp: dst-side row slice + Wt[type], for protein | ligand sources
    qry = reinterpret_cast<float*>(p); p += 16 * 64 * 4;        // per warp: 2-deep ring of 32-float query slices
    stat = reinterpret_cast<float2*>(p); p += 2 * 128 * 4 * 8;  // [parity][row][slice] {sum, sum of squares}
    bars = reinterpret_cast<uint64_t*>(p); p += 128;       // two sets of 8: the second phase of a paired launch uses its own
    tmem_slot = reinterpret_cast<uint32_t*>(p);
  }
  static constexpr int bytes() { return ATC_W2_BYTES + 4 * KT_IMG + (3 * H + 16 * 64 * 2 + 2 * 128 * 4 * 2) * 4 + 160; }
};
static_assert(KnnTcSmem::bytes() <= 232448, "shared memory budget");

// K-major SWIZZLE_32B shared-memory matrix descriptor: rows of 32 bytes (8 tf32), 8-row groups 256 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// byte offset of the 16-byte chunk holding columns [4*c4, 4*c4+4) of row r inside one image
__device__ __forceinline__ int kt_chunk_off(int r, int c4) { return (c4 >> 1) * 4096 + r * 32 + (((c4 & 1) ^ ((r >> 2) & 1)) << 4); }

__device__ __forceinline__ float2 kf2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 ku2f(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }
__device__ __forceinline__ void knamed_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void knamed_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// distances of all kNN edges at layer entry: dist[node][lane] = |x_i - x_j| (rel_x of uni_transformer_edge.py:263-265)
__global__ void __launch_bounds__(256) knn_dist_kernel(const float* __restrict__ x4, const int* __restrict__ nbr, const int* __restrict__ deg,
                                                       int n, float* __restrict__ dist) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int node = idx >> 5, lane = idx & 31;
  if (node >= n) return;
  float d = 0.f;
  if (lane < deg[node]) {
    const float4 xi = ldg4(x4 + (size_t)node * 4), xj = ldg4(x4 + (size_t)__ldg(nbr + idx) * 4);
    const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
    d = sqrtf(dx * dx + dy * dy + dz * dz);
  }
  dist[idx] = d;
}
void launch_knn_dist(const float* x4, const int* nbr, const int* deg, int n, float* dist, cudaStream_t stream) {
  if (n <= 0) return;
  knn_dist_kernel<<<(n * 32 + 255) / 256, 256, 0, stream>>>(x4, nbr, deg, n, dist);
}

// per-slot metadata of a destination list, packed so that the attention kernels need ONE load per group and no dependent chain:
// {node id or -1 (padding), deg | nlig << 8 | is_ligand << 16}
__global__ void __launch_bounds__(256) knn_slot_meta_kernel(const int* __restrict__ dst_list, int n_slots, const int* __restrict__ deg,
                                                            const int* __restrict__ nlig, const uint8_t* __restrict__ is_lig,
                                                            int2* __restrict__ out) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const int node = dst_list ? dst_list[slot] : slot;
  out[slot] = node < 0 ? make_int2(-1, 0) : make_int2(node, deg[node] | (nlig[node] << 8) | ((int)is_lig[node] << 16));
}
void launch_knn_slot_meta(const int* dst_list, int n_slots, const int* deg, const int* nlig, const uint8_t* is_lig, int2* out,
                          cudaStream_t stream) {
  if (n_slots <= 0) return;
  knn_slot_meta_kernel<<<(n_slots + 255) / 256, 256, 0, stream>>>(dst_list, n_slots, deg, nlig, is_lig, out);
}

// PASS: 0 = key pass, 1 = node value pass, 2 = position value pass (16-output second Linear, dx per destination slot)
// `first` / `last`: the key pass and a value pass may run back to back inside ONE launch (knn_tc_pair_kernel): same tiles per CTA in
// both phases, so a phase only reads attention weights its own CTA wrote; barriers are re-initialised and the weight images
// swapped in between, TMEM and the register hand-over are set up once
template <int PASS>
__device__ __forceinline__ void knn_tc_body(const KnnAttnArgs& a, const bool first, const bool last) {
  constexpr bool VPASS = PASS != 0, VPOS = PASS == 2;
  constexpr int W2_BYTES = VPOS ? 2 * NH * 128 * 4 : ATC_W2_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  KnnTcSmem sm(smem_raw);
  uint64_t* const bars = sm.bars + (first ? 0 : 8);      // a fresh barrier set per phase (no re-initialisation of used barriers)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, s = (warp >> 2) & 3, r = q * 32 + lane;
  // barriers: [0] W2 (+ first B2) landed, [1] main MMA retired, [2] distance MMA retired, [3] B2 of the second class landed,
  // [4] Gaussian features of a tile written (3 producer warps), [5] D2 of a tile read by every worker warp
  if ((smem_u32(sm.W2) & 1023u) != 0u) __trap();
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), i == 4 ? 3 : i == 5 ? 16 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (first && warp == 0) { __syncwarp(); tmem_alloc(smem_u32(sm.tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int tile_first = a.n_slots_first / 4;          // tiles below this hold destinations of class a.first_class
  auto class_of = [&](int tile) { return tile < tile_first ? a.first_class : 1 - a.first_class; };
  if (tid == 0) {
    const uint32_t bar = smem_u32(&bars[0]);
    mbar_expect_tx(bar, W2_BYTES + 2 * KT_IMG);
    bulk_g2s(smem_u32(sm.W2), a.W2tc, W2_BYTES / 2, bar);
    bulk_g2s(smem_u32(sm.W2) + W2_BYTES / 2, a.W2tc + W2_BYTES / 8, W2_BYTES / 2, bar);
    bulk_g2s(smem_u32(sm.B2), a.B2tc[class_of(blockIdx.x)], 2 * KT_IMG, bar);
  }
  const uint32_t tmem_base = *sm.tmem_slot;
  cta_copy_f4(sm.gamma, a.w.gamma, H);
  cta_copy_f4(sm.beta, a.w.beta, H);
  if (VPASS && !VPOS) cta_copy_f4(sm.b2, a.w.b2, H);
  if (VPOS && tid < NH) sm.b2[tid] = a.w.b2[tid];
  if (first) pdl_wait();      // everything above is set-up on static data; below this line the previous kernels' results are visible
  const int n_dst = a.n_dst_dev ? min(__ldg(a.n_dst_dev), a.n_dst) : a.n_dst;      // device-side count: receptive-field pruning
  const int n_tiles = (n_dst + 3) / 4;
  __syncthreads();
  mbar_wait(smem_u32(&bars[0]), 0);
  const uint32_t bar_mma = smem_u32(&bars[1]), bar_d2 = smem_u32(&bars[2]), bar_a2f = smem_u32(&bars[4]), bar_d2c = smem_u32(&bars[5]);
  const uint32_t w2_smem = smem_u32(sm.W2), a2_smem = smem_u32(sm.A2), b2_smem = smem_u32(sm.B2);

  if (warp >= 16) {
    // ---------------------------------------------------------------- warp 16 issues the MMAs, warps 17..19 produce the distance features
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");      // a no-op in the second phase of a paired launch (already there)
#endif
    if (warp == 16) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int cur_class = class_of(blockIdx.x);
      auto issue_d2 = [&](int tile, int t) {
        if (lane == 0) {
          mbar_wait(bar_a2f, t & 1);                     // the producers have written the features of this tile
          if (t > 0) mbar_wait(bar_d2c, (t - 1) & 1);     // every worker warp has read D2 of the tile before
          tc_fence_after();
          const int cls = class_of(tile);
          if (cls != cur_class) {                     // class boundary: every distance MMA issued so far has retired (workers
            cur_class = cls;                          // waited on it), so B2 can be replaced
            const uint32_t bar = smem_u32(&bars[3]);
            mbar_expect_tx(bar, 2 * KT_IMG);
            bulk_g2s(b2_smem, a.B2tc[cls], 2 * KT_IMG, bar);
            mbar_wait(bar, 0);
          }
          const uint64_t a_hi = umma_desc_sw32(a2_smem), a_lo = umma_desc_sw32(a2_smem + KT_IMG);
          const uint64_t b_hi = umma_desc_sw32(b2_smem), b_lo = umma_desc_sw32(b2_smem + KT_IMG);
          // tiles without a ligand source (most protein destinations): the ligand column block is all zero, 3 of the 5 k-steps
          // cover the protein block (columns 20..23 of k-step 2 are zeros written by the workers)
          int any_lig = 0;
          for (int g4 = 0; g4 < 4; ++g4) {
            const int slot = tile * 4 + g4;
            if (slot < n_dst) any_lig |= (__ldg(a.slot_meta + slot).y >> 8) & 0xff;
          }
          const int n_kb = any_lig ? KT_KB : 3;
#pragma unroll 1
          for (int kb = 0; kb < n_kb; ++kb) {
            const uint64_t o = (uint64_t)((kb * 4096) >> 4);
            umma_tf32_ss(tmem_base + KT_COL_D2, a_hi + o, b_hi + o, idesc, kb ? 1u : 0u);
            umma_tf32_ss(tmem_base + KT_COL_D2, a_lo + o, b_hi + o, idesc, 1u);
            umma_tf32_ss(tmem_base + KT_COL_D2, a_hi + o, b_lo + o, idesc, 1u);
          }
          umma_commit(bar_d2);
        }
        __syncwarp();
      };
      if ((int)blockIdx.x < n_tiles) issue_d2(blockIdx.x, 0);
      int t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        if (tile + (int)gridDim.x < n_tiles) issue_d2(tile + gridDim.x, t + 1);
        knamed_sync(KBAR_A_READY, KT_SYNC);           // hidden activations are in TMEM, D of the previous tile is in registers
        if (lane == 0) { tc_fence_after(); if (VPOS) atc_issue_mma_n<NH>(tmem_base, w2_smem, bar_mma); else atc_issue_mma(tmem_base, w2_smem, bar_mma); }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- producers: distance of a row -> its 20 Gaussians -> A2.
      // Thread = row.  Warp 17 + p writes quadrant p; quadrant 3 is shared by 4-column chunk ({0,1} | {2,3} | {4}).  A chunk goes to
      // the column block of the row's source class, the other block gets zeros.  (On the worker warps this was ~70 instructions per
      // thread and tile, twice that on the slice-0 warps the rest of the quadrant then waited for.)
      const int pw = warp - 17, step = gridDim.x;
      auto meta_of = [&](int tile, int qq) {
        const int slot = tile * 4 + qq;
        return tile < n_tiles && slot < n_dst ? __ldg(a.slot_meta + slot) : make_int2(-1, 0);
      };
      auto put_chunk = [&](int rr, int c, float d, bool src_lig) {
        float g[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float u = d - c_gauss_offset[c * 4 + i]; g[i] = expf(-0.5f * u * u); }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) tf32_split(g[i], hi[i], lo[i]);
        const uint4 vh = make_uint4(hi[0], hi[1], hi[2], hi[3]), vl = make_uint4(lo[0], lo[1], lo[2], lo[3]), z4 = make_uint4(0u, 0u, 0u, 0u);
        const int off_p = kt_chunk_off(rr, c), off_l = kt_chunk_off(rr, 5 + c);
        *reinterpret_cast<uint4*>(sm.A2 + off_p) = src_lig ? z4 : vh;
        *reinterpret_cast<uint4*>(sm.A2 + KT_IMG + off_p) = src_lig ? z4 : vl;
        *reinterpret_cast<uint4*>(sm.A2 + off_l) = src_lig ? vh : z4;
        *reinterpret_cast<uint4*>(sm.A2 + KT_IMG + off_l) = src_lig ? vl : z4;
      };
      const int c3_lo = pw * 2, c3_hi = pw == 2 ? 5 : pw * 2 + 2;
      int2 m0 = meta_of(blockIdx.x, pw), m3 = meta_of(blockIdx.x, 3);
      int t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += step, ++t) {
        const float d0 = m0.x >= 0 ? __ldg(a.dist + (size_t)m0.x * KNN + lane) : 0.f;
        const float d3 = m3.x >= 0 ? __ldg(a.dist + (size_t)m3.x * KNN + lane) : 0.f;
        const bool l0 = lane < ((m0.y >> 8) & 0xff), l3 = lane < ((m3.y >> 8) & 0xff);
        m0 = meta_of(tile + step, pw); m3 = meta_of(tile + step, 3);
        if (t > 0) mbar_wait(bar_d2, (t - 1) & 1);      // the distance MMA of the tile before has read A2
#pragma unroll
        for (int c = 0; c < 5; ++c) put_chunk(pw * 32 + lane, c, d0, l0);
#pragma unroll
        for (int c = 0; c < 5; ++c) if (c >= c3_lo && c < c3_hi) put_chunk(96 + lane, c, d3, l3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // A2 was written through the generic proxy
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a2f) : "memory");
      }
    }
  } else {
#ifndef DDB_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
#endif
    // ---------------------------------------------------------------- 16 worker warps: thread = (row r, channel slice s)
    struct Grp {            // one destination group, 2 registers: node id (-1: padding) and deg | nlig << 8 | is_ligand << 16
      int node, pk;
      __device__ bool valid() const { return node >= 0; }
      __device__ int deg() const { return pk & 0xff; }
      __device__ int nlig() const { return (pk >> 8) & 0xff; }
      __device__ bool lig_dst() const { return (pk >> 16) != 0; }
    };
    auto load_group = [&](int tile) {
      Grp g; g.node = -1; g.pk = 0;
      const int slot = tile * 4 + q;
      if (tile < n_tiles && slot < n_dst) { const int2 m = __ldg(a.slot_meta + slot); g.node = m.x; g.pk = m.y; }
      return g;
    };
    float* const whit = sm.hit + warp * 64;         // [2 source classes][32]
    float* const wqry = sm.qry + warp * 64;         // [2-deep ring][32]
    const int step = gridDim.x;
    // type bias of the first Linear for this thread's channel, all four edge types (L1 is a few KB next to 226 KB of shared
    // memory, so a per-tile __ldg would be an L2 round trip)
    const float wt0 = __ldg(a.w.Wt + 0 * H + s * 32 + lane), wt1 = __ldg(a.w.Wt + 1 * H + s * 32 + lane),
                wt2 = __ldg(a.w.Wt + 2 * H + s * 32 + lane), wt3 = __ldg(a.w.Wt + 3 * H + s * 32 + lane);

    int it = 0;
    Grp g = load_group(blockIdx.x), g_n = load_group(blockIdx.x + step);
    // row state of the current tile
    auto hidx = [&](const Grp& gg, int tile) { return a.hi_by_slot ? tile * 4 + q : gg.node; };
    int j = lane < g.deg() ? __ldg(a.nbr + (size_t)g.node * KNN + lane) : 0;
    float hi_cur = g.valid() ? __ldg(a.Hi + (size_t)hidx(g, blockIdx.x) * a.ldhi + s * 32 + lane) : 0.f;
    float4 pv[8];
    if ((int)blockIdx.x < n_tiles) {
      const float* prow = a.Hj + (size_t)j * a.ldhj + s * 32;
#pragma unroll
      for (int i8 = 0; i8 < 4; ++i8) ldg8(prow + i8 * 8, pv[2 * i8], pv[2 * i8 + 1]);
    }
    int prev_node = -1, prev_slot = 0; bool prev_ok = false; float prev_ew = 0.f;
    float4 prev_rel = make_float4(0.f, 0.f, 0.f, 0.f);
    TL_DECL
    for (int tile = blockIdx.x; tile < n_tiles; tile += step, ++it) {
      TL_MARK(0);
      const bool rowok = lane < g.deg();            // deg is 0 for padding groups
      // ---- requests for later: rows of the next tile, group metadata two tiles ahead, this tile's query / edge weight
      const int j_n = lane < g_n.deg() ? ldna_i(a.nbr + (size_t)g_n.node * KNN + lane) : 0;
      const float hi_n = g_n.valid() ? ldna_f(a.Hi + (size_t)hidx(g_n, tile + step) * a.ldhi + s * 32 + lane) : 0.f;
      const Grp g_nn = load_group(tile + 2 * step);
      float4 rel = make_float4(0.f, 0.f, 0.f, 0.f);
      if (VPOS && s == 0 && rowok) {          // x_i - x_j of this thread's edge (rel_x, :198-199)
        const float4 xi = ldg4(a.x4 + (size_t)g.node * 4), xj = ldg4(a.x4 + (size_t)j * 4);
        rel = make_float4(xi.x - xj.x, xi.y - xj.y, xi.z - xj.z, 0.f);
      }
      float qry_v = 0.f, ew = 0.f;
      if (!VPASS) {
        if (g.valid()) qry_v = ldna_f(a.q + (size_t)(a.q_by_slot ? tile * 4 + q : g.node) * a.ldq + s * 32 + lane);
        if (rowok) ew = ldna_f(a.e_w + (size_t)g.node * KNN + lane);
      }
      // dst-side term of the first Linear + type bias, for protein and for ligand sources (uni_transformer_edge.py:371-377)
      {
        const bool ld = g.lig_dst();
        whit[lane] = hi_cur + (ld ? wt2 : wt3);            // protein sources: type 2 into a ligand, 3 into a protein destination
        whit[32 + lane] = hi_cur + (ld ? wt0 : wt1);       // ligand sources:  type 0 / 1
        if (!VPASS) wqry[(it & 1) * 32 + lane] = qry_v;
      }
      // ---- first Linear: z = P_src[j] (prefetched) + (P_dst[i] + Wt) (staged) + D2 (distance MMA, issued one iteration ago)
      float2 z[16];
      {
        TL_MARK(1);
        mbar_wait(bar_d2, it & 1);
        TL_MARK(2);
        tc_fence_after();
        __syncwarp();
        const float* hs = whit + (lane < g.nlig() ? 32 : 0);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 hv = ld4(hs + i4 * 4);
          z[i4 * 2] = __fadd2_rn(kf2(pv[i4].x, pv[i4].y), kf2(hv.x, hv.y));
          z[i4 * 2 + 1] = __fadd2_rn(kf2(pv[i4].z, pv[i4].w), kf2(hv.z, hv.w));
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + KT_COL_D2 + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = __fadd2_rn(z[i], ku2f(v[2 * i], v[2 * i + 1]));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_d2c) : "memory");      // D2 may be overwritten
      }
      TL_MARK(3);
      TL_MARK(4);
      // ---- LayerNorm with one exchange between the 4 slice-warps of the quadrant, ReLU
      {
        float2 s1 = kf2(0.f, 0.f), s2 = kf2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) s1 = __fadd2_rn(s1, z[i]);
        const float part = s1.x + s1.y;
        // two-pass statistics need two exchanges; a one-pass variance about the slice mean needs only one: exchange
        // {sum, centred sum of squares} and combine with the parallel-variance formula
        const float mu_s = part * (1.0f / 32.0f);
        const float2 nm = kf2(-mu_s, -mu_s);
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float2 dz = __fadd2_rn(z[i], nm); s2 = __ffma2_rn(dz, dz, s2); }
        float2* st = sm.stat + ((it & 1) * 128 + r) * 4;
        st[s] = make_float2(part, s2.x + s2.y);
        TL_MARK(5);
        quad_barrier(q);
        TL_MARK(6);
        const float4 t01 = *reinterpret_cast<const float4*>(st), t23 = *reinterpret_cast<const float4*>(st + 2);
        const float mu = ((t01.x + t01.z) + (t23.x + t23.z)) * (1.0f / H);
        const float d0 = t01.x * (1.0f / 32.0f) - mu, d1 = t01.z * (1.0f / 32.0f) - mu, d2 = t23.x * (1.0f / 32.0f) - mu, d3 = t23.z * (1.0f / 32.0f) - mu;
        const float m2 = ((t01.y + t01.w) + (t23.y + t23.w)) + 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
        const float rstd = rsqrtf(m2 * (1.0f / H) + LN_EPS);
        const float2 rs2 = kf2(rstd, rstd), nm2 = kf2(-mu * rstd, -mu * rstd);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 gm = ld4(sm.gamma + s * 32 + i4 * 4), bt = ld4(sm.beta + s * 32 + i4 * 4);
          float2 u0 = __ffma2_rn(z[i4 * 2], rs2, nm2), u1 = __ffma2_rn(z[i4 * 2 + 1], rs2, nm2);      // (z - mu) * rstd
          u0 = __ffma2_rn(u0, kf2(gm.x, gm.y), kf2(bt.x, bt.y));
          u1 = __ffma2_rn(u1, kf2(gm.z, gm.w), kf2(bt.z, bt.w));
          z[i4 * 2] = kf2(fmaxf(u0.x, 0.f), fmaxf(u0.y, 0.f));
          z[i4 * 2 + 1] = kf2(fmaxf(u1.x, 0.f), fmaxf(u1.y, 0.f));
        }
      }
      uint32_t v[32];
      float lg[4] = {0.f, 0.f, 0.f, 0.f};
      if (VPASS) {
        // ---- value passes: TF32 split BEFORE the wait on the tensor core (hi = z truncated to TF32, lo = z - hi, exact); between
        // "main MMA of the previous tile retired" and "main MMA of this tile may start" only the A store and the D load remain
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          hi[2 * i] = __float_as_uint(z[i].x) & 0xffffe000u;
          hi[2 * i + 1] = __float_as_uint(z[i].y) & 0xffffe000u;
          const float2 l = __fadd2_rn(z[i], kf2(-__uint_as_float(hi[2 * i]), -__uint_as_float(hi[2 * i + 1])));
          lo[2 * i] = __float_as_uint(l.x); lo[2 * i + 1] = __float_as_uint(l.y);
        }
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        if (it > 0) { mbar_wait(bar_mma, (it - 1) & 1); tc_fence_after(); }
        tmem_st32(lane_addr + ATC_COL_AHI + s * 32, hi);
        tmem_st32(lane_addr + ATC_COL_ALO + s * 32, lo);
        if (it > 0) {
          if (VPOS) { if (s == 0) tmem_ld16(lane_addr + ATC_COL_D, reinterpret_cast<uint32_t (&)[16]>(v)); }
          else tmem_ld32(lane_addr + ATC_COL_D + s * 32, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (it > 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        knamed_arrive(KBAR_A_READY, KT_SYNC);
      } else {
        // ---- key pass (measured: holding hi | lo across the wait costs more in spills than it saves): drain D into the 4 head
        // logits, then split and store 16 columns at a time
        if (it > 0) {
          mbar_wait(bar_mma, (it - 1) & 1);
          tc_fence_after();
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D + s * 32, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float* qr = wqry + ((it - 1) & 1) * 32;
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            const float4 q0 = ld4(qr + hh * 8), q1 = ld4(qr + hh * 8 + 4);
            float2 acc = __fmul2_rn(kf2(q0.x, q0.y), ku2f(v[hh * 8], v[hh * 8 + 1]));
            acc = __ffma2_rn(kf2(q0.z, q0.w), ku2f(v[hh * 8 + 2], v[hh * 8 + 3]), acc);
            acc = __ffma2_rn(kf2(q1.x, q1.y), ku2f(v[hh * 8 + 4], v[hh * 8 + 5]), acc);
            acc = __ffma2_rn(kf2(q1.z, q1.w), ku2f(v[hh * 8 + 6], v[hh * 8 + 7]), acc);
            lg[hh] = prev_ok ? acc.x + acc.y : -INFINITY;
          }
        }
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 zz = z[half * 8 + i];
            hi[2 * i] = __float_as_uint(zz.x) & 0xffffe000u;
            hi[2 * i + 1] = __float_as_uint(zz.y) & 0xffffe000u;
            const float2 l = __fadd2_rn(zz, kf2(-__uint_as_float(hi[2 * i]), -__uint_as_float(hi[2 * i + 1])));
            lo[2 * i] = __float_as_uint(l.x); lo[2 * i + 1] = __float_as_uint(l.y);
          }
          tmem_st16(lane_addr + ATC_COL_AHI + s * 32 + half * 16, hi);
          tmem_st16(lane_addr + ATC_COL_ALO + s * 32 + half * 16, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        knamed_arrive(KBAR_A_READY, KT_SYNC);
      }
      // ---- row gather of the next tile (consumed one iteration from now)
      {
        const float* prow = a.Hj + (size_t)j_n * a.ldhj + s * 32;
#pragma unroll
        for (int i8 = 0; i8 < 4; ++i8) ldg8(prow + i8 * 8, pv[2 * i8], pv[2 * i8 + 1]);
      }
      // ---- epilogue of the previous tile from registers while the tensor core works
      if (it > 0) {
        if (!VPASS) {
          float ex[4];
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) {
            float m;
            asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(lg[hh]));
            ex[hh] = prev_ok ? __expf(lg[hh] - m) : 0.f;
          }
          float sum[4] = {ex[0], ex[1], ex[2], ex[3]};
          warp_allreduce4(sum, lane);
          float w[4];
#pragma unroll
          for (int hh = 0; hh < 4; ++hh) w[hh] = sum[hh] > 0.f ? __fdividef(ex[hh], sum[hh]) * prev_ew : 0.f;
          if (prev_ok) st4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + s * 4, make_float4(w[0], w[1], w[2], w[3]));
        } else if (VPOS) {
          if (s == 0) {           // 16 head outputs of this row; c = sum_h (alpha e_w)[h] (D[h] + b2[h])   (:199-208)
            float cpos = 0.f;
#pragma unroll
            for (int h4 = 0; h4 < 4; ++h4) {
              const float4 wp = prev_ok ? ldna_c4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + h4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
              cpos = fmaf(wp.x, __uint_as_float(v[h4 * 4 + 0]) + sm.b2[h4 * 4 + 0], cpos);
              cpos = fmaf(wp.y, __uint_as_float(v[h4 * 4 + 1]) + sm.b2[h4 * 4 + 1], cpos);
              cpos = fmaf(wp.z, __uint_as_float(v[h4 * 4 + 2]) + sm.b2[h4 * 4 + 2], cpos);
              cpos = fmaf(wp.w, __uint_as_float(v[h4 * 4 + 3]) + sm.b2[h4 * 4 + 3], cpos);
            }
            const float ax = warp_sum(cpos * prev_rel.x), ay = warp_sum(cpos * prev_rel.y), az = warp_sum(cpos * prev_rel.z);
            if (lane == 0 && prev_node >= 0) st4(a.out_dx + (size_t)prev_slot * 4, make_float4(ax * (1.f / NH), ay * (1.f / NH), az * (1.f / NH), 0.f));
          }
        } else {
          float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (prev_ok) w4 = ldna_c4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + s * 4);
          float val[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float wh = (i < 8) ? w4.x : (i < 16) ? w4.y : (i < 24) ? w4.z : w4.w;
            val[i] = wh * __uint_as_float(v[i]);
          }
          warp_reduce_scatter<32>(val, lane);
          float ws[4] = {w4.x, w4.y, w4.z, w4.w};
          warp_allreduce4(ws, lane);
          if (prev_node >= 0) {
            const int c = s * 32 + lane;
            a.out_h[(size_t)prev_node * a.ldo + c] = val[0] + sm.b2[c] * ws[lane >> 3];
          }
        }
      }
      prev_node = g.node; prev_ok = rowok; prev_ew = ew; prev_rel = rel; prev_slot = tile * 4 + q;
      g = g_n; g_n = g_nn; j = j_n; hi_cur = hi_n;
      TL_MARK(10);
    }
    TL_FLUSH(PASS == 1 ? 1 : 0);
    // ---- epilogue of the last tile
    if (it > 0) {
      mbar_wait(bar_mma, (it - 1) & 1);
      tc_fence_after();
      if (!VPASS) {
        float4 w4 = atc_logits_softmax(tmem_base, q, s, wqry + ((it - 1) & 1) * 32 - s * 32, prev_ok);
        if (prev_ok) st4(a.wbuf + ((size_t)prev_node * KNN + lane) * NH + s * 4,
                         make_float4(w4.x * prev_ew, w4.y * prev_ew, w4.z * prev_ew, w4.w * prev_ew));
      } else if (VPOS) {
        if (s == 0) {
          uint32_t v16[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + ATC_COL_D, v16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float cpos = 0.f;
          if (prev_ok) {
#pragma unroll
            for (int h = 0; h < NH; ++h) cpos = fmaf(a.wbuf[((size_t)prev_node * KNN + lane) * NH + h], __uint_as_float(v16[h]) + sm.b2[h], cpos);
          }
          const float ax = warp_sum(cpos * prev_rel.x), ay = warp_sum(cpos * prev_rel.y), az = warp_sum(cpos * prev_rel.z);
          if (lane == 0 && prev_node >= 0) st4(a.out_dx + (size_t)prev_slot * 4, make_float4(ax * (1.f / NH), ay * (1.f / NH), az * (1.f / NH), 0.f));
        }
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
  }
  tc_fence_before();
  __syncthreads();      // also orders this phase's global writes (attention weights) before the next phase's reads within the CTA
  if (last && warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int PASS>
__global__ void __launch_bounds__(KT_THREADS, 1) knn_tc_kernel(const KnnAttnArgs a) { knn_tc_body<PASS>(a, true, true); }
// key pass + value pass of one edge family in one launch (P2 = 1: node update, 2: position update)
template <int P2>
__global__ void __launch_bounds__(KT_THREADS, 1) knn_tc_pair_kernel(const KnnAttnArgs ak, const KnnAttnArgs av) {
  knn_tc_body<0>(ak, true, false);
  knn_tc_body<P2>(av, false, true);
}

void launch_knn_tc_pair(const KnnAttnArgs& ak, const KnnAttnArgs& av, bool pos, int num_sms, cudaStream_t stream) {
  if (ak.n_dst <= 0) return;
  static DeviceOnce once;
  const int bytes = KnnTcSmem::bytes();
  if (!once.done()) {
    cudaFuncSetAttribute(knn_tc_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(knn_tc_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    once.mark();
  }
  const int grid = atc_grid((ak.n_dst + 3) / 4, num_sms);
  if (pos) launch_pdl(knn_tc_pair_kernel<2>, dim3(grid), dim3(KT_THREADS), bytes, stream, ak, av);
  else launch_pdl(knn_tc_pair_kernel<1>, dim3(grid), dim3(KT_THREADS), bytes, stream, ak, av);
}

void launch_knn_tc(const KnnAttnArgs& a, int pass, int num_sms, cudaStream_t stream) {      // pass: 0 key, 1 node value, 2 position value
  if (a.n_dst <= 0) return;
  static DeviceOnce once;
  const int bytes = KnnTcSmem::bytes();
  if (!once.done()) {
    cudaFuncSetAttribute(knn_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(knn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(knn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    once.mark();
  }
  const int grid = atc_grid((a.n_dst + 3) / 4, num_sms);
  if (pass == 2) launch_pdl(knn_tc_kernel<2>, dim3(grid), dim3(KT_THREADS), bytes, stream, a);
  else if (pass == 1) launch_pdl(knn_tc_kernel<1>, dim3(grid), dim3(KT_THREADS), bytes, stream, a);
  else launch_pdl(knn_tc_kernel<0>, dim3(grid), dim3(KT_THREADS), bytes, stream, a);
}

#ifdef DDB_TIMELINE
extern "C" int ddb_debug_timeline_knn(unsigned long long* out /* 2*2*16 */) {
  return (int)cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * 64);
}
#endif

// host-side packing of the distance-term weights for one destination class: B2[n = channel][k], k < 20: Wg[type_p][k][n],
// 20 <= k < 40: Wg[type_l][k-20][n]; hi | lo images in the SWIZZLE_32B K-major layout of the kernel
void pack_wg_tc(const float* Wg /* [4*20][128] */, int type_p, int type_l, float* out /* 2 * KT_IMG / 4 floats */) {
  for (int i = 0; i < 2 * KT_IMG / 4; ++i) out[i] = 0.f;
  float* hi = out;
  float* lo = out + KT_IMG / 4;
  for (int n = 0; n < 128; ++n)
    for (int k = 0; k < 40; ++k) {
      const int type = k < 20 ? type_p : type_l, g = k < 20 ? k : k - 20;
      const float w = Wg[(size_t)(type * NG + g) * H + n];
      const float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
      const int off = ((k >> 3) * 4096 + n * 32 + ((((k >> 2) & 1) ^ ((n >> 2) & 1)) << 4) + (k & 3) * 4) / 4;
      hi[off] = h; lo[off] = l;
    }
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
