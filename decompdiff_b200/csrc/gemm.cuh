// fp32 SIMT GEMM for the per-node / per-edge projections:  C = act(prologue(A)[M,128] @ Wt[128,N] + bias) (+ R)
//
// Why fp32 FMA and not tensor cores: SURVEY.md section 7 measured that TF32 / BF16 inputs break the
// rtol 1e-4 / atol 1e-5 contract of the path (3.6x .. 208x out of tolerance); only fp32 FMA or a
// 3xTF32 split is admissible.  K is always 128 (hidden_dim) so the whole A tile is staged once,
// which lets the gather / add / LayerNorm+ReLU prologues run on complete rows in shared memory.
#pragma once
#include "common.cuh"

namespace ddb {

struct GemmArgs {
  const float* A = nullptr; int lda = 0;          // M x 128 rows
  const int* a_rows = nullptr;                    // optional gather: A row of output row m
  const float* A2 = nullptr; int lda2 = 0;        // optional rows added to A before LN (requires a2_rows)
  const int* a2_rows = nullptr;
  const float* ln_gamma = nullptr;                // optional LayerNorm+ReLU prologue (both or none)
  const float* ln_beta = nullptr;
  const float* Wt = nullptr; int ldw = 0;         // 128 x N, row k contiguous over n
  const float* bias = nullptr;                    // N or null
  const float* R = nullptr; int ldr = 0;          // optional residual, rows indexed like C
  float* C = nullptr; int ldc = 0;
  const int* c_rows = nullptr;                    // optional scatter of output (and residual) rows
  int M = 0, N = 0;                               // N multiple of 128
  const int* M_dev = nullptr;                     // optional: the row count lives on the device (<= M, which sizes the grid);
                                                  // rows whose a_rows / c_rows entry is negative are padding and are skipped
  int act = 0;                                    // 0 none, 1 shifted softplus
};

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 32, GEMM_THREADS = 256;
constexpr int GEMM_SMEM = (H * GEMM_BM + 2 * GEMM_BK * GEMM_BN) * 4;   // 64 KB A + 32 KB W stages

constexpr int GEMM_MAX_BATCH = 4;
struct GemmBatch {       // kernel argument of the batched tensor-core launch (gemm_tc.cu)
  GemmArgs p[GEMM_MAX_BATCH]; const float* Wtc[GEMM_MAX_BATCH]; int per[GEMM_MAX_BATCH], gx[GEMM_MAX_BATCH], gy[GEMM_MAX_BATCH];
};

void launch_gemm128(const GemmArgs& a, cudaStream_t stream);
// up to GEMM_MAX_BATCH independent problems in one launch
void launch_gemm128_tc_batch(const GemmArgs* args, const float* const* Wtc, int n, int num_sms, cudaStream_t stream);
// tcgen05 / TMEM 3xTF32 version (gemm_tc.cu); Wtc = weight packed by pack_gemm_tc (2*128*N floats)
void launch_gemm128_tc(const GemmArgs& a, const float* Wtc, int num_sms, cudaStream_t stream);
void pack_gemm_tc(const float* Wt, int N, float* out);

}  // namespace ddb
