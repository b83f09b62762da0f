// Tensor-core projection GEMM for sm_100a:  C = act(prologue(A)[M,128] @ W[128,N] + bias) (+ R)
// tcgen05.mma (kind::tf32), A operand AND accumulators in TMEM, B (weights) streamed through shared memory.
//
// fp32 parity: plain TF32 inputs break the rtol 1e-4 contract, so every product is evaluated as
// the 3xTF32 split  a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  (a_hi = rna_tf32(a), a_lo = rna_tf32(a - a_hi)), three
// MMAs into the same fp32 accumulator; the dropped a_lo*b_lo term is O(2^-22).
//
// Structure (one CTA = 16 worker warps + 1 MMA-issuing warp + 1 weight-producer warp, per 128-row tile x a range of 128-column
// output tiles):
//   * worker thread (row r = 32q + lane, channel slice s) of warp w = 4s + q stages 32 channels of one row: gather / add /
//     LayerNorm+ReLU prologue (row statistics exchanged through smem between the 4 warps of a quadrant), hi/lo split,
//     tcgen05.st into TMEM (A_hi columns 0..127, A_lo columns 128..255; row -> lane, k -> column)
//   * B is pre-split and pre-swizzled on the host into K-blocks: for every 128-column output tile, 4 blocks of 32 k-values,
//     each 32 KB (hi | lo, [128 n-rows][128 B], K-major, 128B swizzle); the producer warp streams them with ONE cp.async.bulk
//     each into a 4-stage ring (mbarrier complete_tx), re-using a stage as soon as the MMAs that read it have retired
//   * the issuer warp waits for a stage, issues 12 tcgen05.mma (M128 N128 K8, A from TMEM, 64 tensor cycles each - wide enough
//     to hide the issue cost of a single thread, which N=32 tiles did not) and commits to the stage's "empty" barrier;
//     after the 4th K-block it also commits the accumulator's "full" barrier.  Two accumulators (TMEM columns 256..511)
//     alternate, so the epilogue of tile t overlaps the MMAs of tile t+1.  No CTA-wide barrier inside the tile loop.
//   * epilogue: warp (s, q) drains rows 32q.. of columns 32s.. of the accumulator: tcgen05.ld, release, smem transpose, bias /
//     residual / activation, coalesced global stores
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "gemm.cuh"
#include "tc_common.cuh"

namespace ddb {

constexpr int TC_BM = 128;              // rows per CTA (UMMA M)
constexpr int TC_BN = 128;              // columns per output tile (UMMA N)
constexpr int TC_STAGES = 4;            // weight ring depth (K-blocks in flight)
constexpr int TC_B_PART = TC_BN * 128;                  // 16 KB: [128 n-rows][32 tf32 = 128 B] hi or lo
constexpr int TC_B_STAGE = 2 * TC_B_PART;               // 32 KB per K-block (hi | lo)
constexpr int TC_WORKERS = 512;                         // 16 staging / epilogue warps
constexpr int TC_THREADS = TC_WORKERS + 64;             // + MMA-issuing warp + weight-producer warp
constexpr int TC_EPI_LD = 36;                           // padded row of a per-warp 32x32 transpose tile (floats)
constexpr int TC_EPI_BYTES = 16 * 32 * TC_EPI_LD * 4;   // 72 KB
constexpr int TC_STAT_BYTES = 2 * 128 * 4 * 4;          // LayerNorm partial sums [row][slice] x {sum, centred squares}
constexpr int TC_SMEM = TC_STAGES * TC_B_STAGE + TC_EPI_BYTES + TC_STAT_BYTES + 128 /*barriers*/;
constexpr int TC_COL_AHI = 0, TC_COL_ALO = 128, TC_COL_D = 256;   // TMEM column map (512 allocated): D0 256.., D1 384..
constexpr int TC_BAR_A_READY = 5;
static_assert(TC_SMEM <= 232448, "shared memory budget");

__device__ __forceinline__ float ssp(float x) { return (x > 20.f ? x : log1pf(expf(x))) - 0.69314718055994530942f; }
__device__ __forceinline__ void tc_quad_barrier(int q) { asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(128) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// One launch serves up to GEMM_MAX_BATCH independent problems (a flat grid, problem after problem): the projections of a layer phase share a
// launch, so the 15-CTA ligand problems run next to the wide ones instead of paying a launch of their own.
__global__ void __launch_bounds__(TC_THREADS, 1) gemm128_tc_kernel(const GemmBatch gb) {
  // flat grid: the CTAs of problem 0 come first, then those of problem 1, ... (no empty CTAs; the launcher gives every problem a
  // share of the SMs in proportion to its work, so the whole batch is one wave of persistent CTAs)
  int pz = 0, bid = blockIdx.x;
  while (pz + 1 < GEMM_MAX_BATCH && bid >= gb.gx[pz] * gb.gy[pz]) { bid -= gb.gx[pz] * gb.gy[pz]; ++pz; }
  const int bx = bid % gb.gx[pz], by = bid / gb.gx[pz];
  const GemmArgs& a = gb.p[pz];
  const float* __restrict__ Wtc = gb.Wtc[pz];
  const int tiles_per_cta = gb.per[pz], grid_x = gb.gx[pz];
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sB = smem_raw;                                                // TC_STAGES x (hi | lo)
  float* sEpi = reinterpret_cast<float*>(sB + TC_STAGES * TC_B_STAGE);   // per-warp 32 x 36 transpose tiles
  float* sStatA = sEpi + 16 * 32 * TC_EPI_LD;
  float* sStatB = sStatA + 128 * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStatB + 128 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();      // the row count below may be a device-side counter of the previous kernels
  const int M = a.M_dev ? min(__ldg(a.M_dev), a.M) : a.M;      // device-side row count: exact receptive-field pruning
  const int row_tiles = (M + TC_BM - 1) / TC_BM;
  if (bx >= row_tiles) return;
  // persistent over row tiles: bx, bx + grid_x, ...  (one TMEM allocation / barrier set-up per CTA; with a
  // single output tile the whole weight stays resident in the ring and is streamed once).
  const int my_rows = (row_tiles - bx + grid_x - 1) / grid_x;
  const int tile0 = by * tiles_per_cta;
  const int n_tiles = min(tiles_per_cta, a.N / TC_BN - tile0);
  auto bar_full = [&](int i) { return smem_u32(&bars[i]); };                        // K-block landed in stage i
  auto bar_empty = [&](int i) { return smem_u32(&bars[TC_STAGES + i]); };           // MMAs reading stage i retired
  auto bar_dfull = [&](int i) { return smem_u32(&bars[2 * TC_STAGES + i]); };       // accumulator i complete
  auto bar_dempty = [&](int i) { return smem_u32(&bars[2 * TC_STAGES + 2 + i]); };  // accumulator i drained (16 warp arrivals)

  if ((smem_u32(sB) & 1023u) != 0u) __trap();      // SWIZZLE_128B operands need a 1024-byte aligned base
  if (tid == 0) {
    for (int i = 0; i < 2 * TC_STAGES + 2; ++i) mbar_init(smem_u32(&bars[i]), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_dempty(i), 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) { __syncwarp(); tmem_alloc(smem_u32(tmem_slot), 512); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_blocks = n_tiles * 4;                 // K-blocks per row tile
  const bool resident = n_blocks <= TC_STAGES;      // the ring holds the whole weight: load once, never refill

  if (warp == 17) {
    // ------------------------------------------------------------------------------------ weight producer warp
    if (lane == 0) {
      const float* src = Wtc + (size_t)tile0 * 4 * (TC_B_STAGE / 4);
      int gi = 0;                                   // running K-block index over all row tiles of this CTA
      for (int rt = 0; rt < (resident ? 1 : my_rows); ++rt)
        for (int i = 0; i < n_blocks; ++i, ++gi) {
          const int st = gi % TC_STAGES, use = gi / TC_STAGES;
          if (use > 0) mbar_wait(bar_empty(st), (use - 1) & 1);
          mbar_expect_tx(bar_full(st), TC_B_STAGE);
          bulk_g2s(smem_u32(sB + st * TC_B_STAGE), src + (size_t)i * (TC_B_STAGE / 4), TC_B_STAGE, bar_full(st));
        }
    }
    __syncwarp();
  } else if (warp == 16) {
    // ------------------------------------------------------------------------------------ MMA issuer warp
    // instruction descriptor: D=F32, A=B=TF32, both K-major, N=128, M=128 (cute::UMMA::InstrDescriptor)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint64_t desc0 = umma_desc_sw128(smem_u32(sB));     // descriptors of other addresses differ in the low field only
    int gi = 0, tg = 0;                             // running K-block / output-tile counters over all row tiles
    for (int rt = 0; rt < my_rows; ++rt) {
      asm volatile("bar.sync %0, %1;" ::"r"(TC_BAR_A_READY), "r"(TC_WORKERS + 32) : "memory");   // A tile of this row tile is in TMEM
      if (lane == 0) {
        tc_fence_after();
        for (int t = 0; t < n_tiles; ++t, ++tg) {
          const int db = tg & 1;
          if (tg >= 2) mbar_wait(bar_dempty(db), ((tg >> 1) - 1) & 1);      // the epilogue two tiles back has drained this accumulator
          const uint32_t d = tmem_base + TC_COL_D + db * TC_BN;
#pragma unroll 1
          for (int kb = 0; kb < 4; ++kb, ++gi) {
            const int st = resident ? (t * 4 + kb) : gi % TC_STAGES;
            mbar_wait(bar_full(st), resident ? 0u : (uint32_t)((gi / TC_STAGES) & 1));
            tc_fence_after();
            const uint64_t d_hi = desc0 + (uint64_t)((st * TC_B_STAGE) >> 4), d_lo = d_hi + (uint64_t)(TC_B_PART >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {   // UMMA K = 8 tf32: 8 TMEM columns of A, 32 bytes inside the 128-byte swizzle atom of B
              const uint32_t acol = kb * 32 + kk * 8;
              umma_tf32_ts(d, tmem_base + TC_COL_AHI + acol, d_hi + (uint64_t)(kk * 2), idesc, (kb | kk) ? 1u : 0u);
              umma_tf32_ts(d, tmem_base + TC_COL_ALO + acol, d_hi + (uint64_t)(kk * 2), idesc, 1u);
              umma_tf32_ts(d, tmem_base + TC_COL_AHI + acol, d_lo + (uint64_t)(kk * 2), idesc, 1u);
            }
            if (!resident) umma_commit(bar_empty(st));      // stage may be refilled
          }
          umma_commit(bar_dfull(db));             // accumulator complete (implies tcgen05.fence::before_thread_sync)
        }
      }
      __syncwarp();
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------ worker warps
    const int q = warp & 3, s = warp >> 2;                  // TMEM lane quadrant, channel slice
    float* tile = sEpi + warp * 32 * TC_EPI_LD;
    int tg = 0;
    // rows of a row tile -> registers: this thread owns channels [32s, 32s+32) of row 32q + lane.  The loads of row tile rt + 1 are
    // issued BEFORE the epilogues of row tile rt, so the gather latency (two dependent L2 round trips with a row map) is off the
    // path between the last MMA of one row tile and the first MMA of the next.
    float z[32];
    auto load_rows = [&](int rt) {
      const int m = (bx + rt * grid_x) * TC_BM + q * 32 + lane;
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = 0.f;
      const int ar = m < M ? (a.a_rows ? a.a_rows[m] : m) : -1;
      if (ar >= 0) {
        const float* src = a.A + (size_t)ar * a.lda + s * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {      // LDG.256: a thread walks its 128-byte row slice in 4 instead of 8 L1 tag look-ups per line
          float4 v, w;
          ldg8(src + i * 8, v, w);
          z[8 * i] = v.x; z[8 * i + 1] = v.y; z[8 * i + 2] = v.z; z[8 * i + 3] = v.w;
          z[8 * i + 4] = w.x; z[8 * i + 5] = w.y; z[8 * i + 6] = w.z; z[8 * i + 7] = w.w;
        }
        if (a.A2) {
          int r2 = a.a2_rows[m];
          if (r2 >= 0) {
            const float* s2 = a.A2 + (size_t)r2 * a.lda2 + s * 32;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 v, w;
              ldg8(s2 + i * 8, v, w);
              z[8 * i] += v.x; z[8 * i + 1] += v.y; z[8 * i + 2] += v.z; z[8 * i + 3] += v.w;
              z[8 * i + 4] += w.x; z[8 * i + 5] += w.y; z[8 * i + 6] += w.z; z[8 * i + 7] += w.w;
            }
          }
        }
      }
    };
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    auto ln_relu = [&]() {       // LayerNorm + ReLU over the full row: two-pass statistics via smem
      if (a.ln_gamma == nullptr) return;
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) p += z[i];
      sStatA[r * 4 + s] = p;
      tc_quad_barrier(q);
      float4 t = ld4(sStatA + r * 4);
      const float mu = ((t.x + t.y) + (t.z + t.w)) * (1.0f / H);
      p = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) { z[i] -= mu; p = fmaf(z[i], z[i], p); }
      sStatB[r * 4 + s] = p;
      tc_quad_barrier(q);
      t = ld4(sStatB + r * 4);
      const float rstd = 1.0f / sqrtf(((t.x + t.y) + (t.z + t.w)) * (1.0f / H) + LN_EPS);
#pragma unroll
      for (int i4 = 0; i4 < 8; ++i4) {
        const float4 g = ldg4(a.ln_gamma + s * 32 + i4 * 4), b = ldg4(a.ln_beta + s * 32 + i4 * 4);
        z[i4 * 4 + 0] = fmaxf(fmaf(z[i4 * 4 + 0] * rstd, g.x, b.x), 0.f);
        z[i4 * 4 + 1] = fmaxf(fmaf(z[i4 * 4 + 1] * rstd, g.y, b.y), 0.f);
        z[i4 * 4 + 2] = fmaxf(fmaf(z[i4 * 4 + 2] * rstd, g.z, b.z), 0.f);
        z[i4 * 4 + 3] = fmaxf(fmaf(z[i4 * 4 + 3] * rstd, g.w, b.w), 0.f);
      }
    };
    auto store_a = [&]() {       // hi / lo split -> TMEM, 16 columns at a time, then the hand-over to the issuing warp
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { float h = tf32_rna(z[half * 16 + i]); hi[i] = __float_as_uint(h); lo[i] = __float_as_uint(tf32_rna(z[half * 16 + i] - h)); }
        tmem_st16(lane_addr + TC_COL_AHI + s * 32 + half * 16, hi);
        tmem_st16(lane_addr + TC_COL_ALO + s * 32 + half * 16, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      asm volatile("bar.arrive %0, %1;" ::"r"(TC_BAR_A_READY), "r"(TC_WORKERS + 32) : "memory");
    };
    load_rows(0);
    ln_relu();
    store_a();
    for (int rt = 0; rt < my_rows; ++rt) {
      const int row0 = (bx + rt * grid_x) * TC_BM;
      const bool more = rt + 1 < my_rows;
      if (more) load_rows(rt + 1);      // in flight during the epilogues below
      // ---- epilogue: warp (s, q) drains rows 32q.., columns 32s.. of every output tile.  With the LAST output tile of a row tile
      // the A operand is free as well (its commit covers every MMA of the row tile): the next row tile's A is stored and handed
      // over between the accumulator load and the global stores, so the tensor core restarts under this epilogue.
      for (int t = 0; t < n_tiles; ++t, ++tg) {
        const int db = tg & 1;
        const bool last_tile = t + 1 == n_tiles;
        if (last_tile && more) ln_relu();      // needs the prefetched rows; runs under the MMAs of the last output tile
        mbar_wait(bar_dfull(db), (tg >> 1) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(lane_addr + TC_COL_D + db * TC_BN + s * 32, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dempty(db));           // this warp's part of the accumulator may be overwritten
        if (last_tile && more) { tc_fence_after(); store_a(); }
        // thread = row -> transpose through smem so that 8 lanes write one 128-byte segment of a row
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          st4(tile + lane * TC_EPI_LD + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        __syncwarp();
        const int n0 = (tile0 + t) * TC_BN + s * 32 + (lane & 7) * 4;
        const float4 bias4 = a.bias ? ldg4(a.bias + n0) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rl = it * 4 + (lane >> 3);
          const int m = row0 + q * 32 + rl;
          const int cr = m < M ? (a.c_rows ? a.c_rows[m] : m) : -1;
          if (cr >= 0) {
            float4 o = add4(ld4(tile + rl * TC_EPI_LD + (lane & 7) * 4), bias4);
            if (a.R) o = add4(o, ld4(a.R + (size_t)cr * a.ldr + n0));
            if (a.act == 1) { o.x = ssp(o.x); o.y = ssp(o.y); o.z = ssp(o.z); o.w = ssp(o.w); }
            st4(a.C + (size_t)cr * a.ldc + n0, o);
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

void launch_gemm128_tc_batch(const GemmArgs* args, const float* const* Wtc, int n, int num_sms, cudaStream_t stream) {
  static DeviceOnce attr_set;
  if (!attr_set.done()) {
    cudaFuncSetAttribute(gemm128_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    attr_set.mark();
  }
  GemmBatch gb;
  int np = 0;
  // every problem gets SMs in proportion to its work (row tiles x (A staging + output tiles), in units of one output tile of MMAs)
  double total_work = 0.0;
  for (int i = 0; i < n; ++i)
    if (args[i].M > 0 && args[i].N > 0) total_work += (double)((args[i].M + TC_BM - 1) / TC_BM) * (1.5 + args[i].N / TC_BN);
  int cta_total = 0;
  for (int i = 0; i < n && np < GEMM_MAX_BATCH; ++i) {
    const GemmArgs& a = args[i];
    if (a.M <= 0 || a.N <= 0) continue;
    const int row_tiles = (a.M + TC_BM - 1) / TC_BM, tiles = a.N / TC_BN;
    const int share = std::max(1, (int)(num_sms * ((double)row_tiles * (1.5 + tiles)) / total_work));
    // Split the output tiles over `nsplit` CTAs per row tile.  Cost model in units of one output tile of MMA work: every CTA
    // pays ~1.5 units to stage its A tile, then tiles/nsplit units; the problem owns `share` SMs (1 CTA per SM).
    int best = 1; double best_cost = 1e30;
    for (int ns = 1; ns <= tiles; ++ns) {
      if (tiles % ns) continue;
      // CTAs are persistent over row tiles: share / ns of them per column range, each walking ceil(row_tiles / that) row tiles
      const int per_col = std::max(1, std::min(row_tiles, share / ns));
      const double rows_each = (double)((row_tiles + per_col - 1) / per_col);
      const double cost = 1.0 + rows_each * (1.5 + (double)tiles / ns);
      if (cost < best_cost - 1e-9) { best_cost = cost; best = ns; }
    }
    static const int force_ns = getenv("DDB_GEMM_NS") ? atoi(getenv("DDB_GEMM_NS")) : 0;      // experiment: fixed column split
    if (force_ns > 0 && tiles % force_ns == 0) best = force_ns;
    const int per = tiles / best;
    gb.p[np] = a; gb.Wtc[np] = Wtc[i]; gb.per[np] = per;
    gb.gx[np] = std::max(1, std::min(row_tiles, share / best)); gb.gy[np] = (tiles + per - 1) / per;
    cta_total += gb.gx[np] * gb.gy[np];
    ++np;
  }
  if (np == 0) return;
  for (int i = np; i < GEMM_MAX_BATCH; ++i) { gb.gx[i] = 0; gb.gy[i] = 0; gb.per[i] = 0; gb.Wtc[i] = nullptr; }
  launch_pdl(gemm128_tc_kernel, dim3(cta_total), dim3(TC_THREADS), TC_SMEM, stream, gb);
}

void launch_gemm128_tc(const GemmArgs& a, const float* Wtc, int num_sms, cudaStream_t stream) {
  launch_gemm128_tc_batch(&a, &Wtc, 1, num_sms, stream);
}

// host-side packing of a K-major weight Wt[128][N] into the image the kernel streams: for every 128-column output tile, 4
// K-blocks of 32 k-values, each hi | lo with rows = output column n (128 of them), 128-byte rows, 128B swizzle
void pack_gemm_tc(const float* Wt, int N, float* out /* N*128*2 floats */) {
  const int tiles = N / TC_BN;
  for (int t = 0; t < tiles; ++t)
    for (int kb = 0; kb < 4; ++kb) {
      float* hi = out + (size_t)(t * 4 + kb) * (TC_B_STAGE / 4);
      float* lo = hi + TC_B_PART / 4;
      for (int nl = 0; nl < TC_BN; ++nl)
        for (int kk = 0; kk < 32; ++kk) {
          const float w = Wt[(size_t)(kb * 32 + kk) * N + t * TC_BN + nl];
          const float h = host_tf32_rna(w), l = host_tf32_rna(w - h);
          const int off = (nl * 128 + (((kk >> 2) ^ (nl & 7)) << 4) + (kk & 3) * 4) / 4;
          hi[off] = h;
          lo[off] = l;
        }
    }
}

}  // namespace ddb
