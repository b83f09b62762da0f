// Embeddings, classification heads, the reverse-diffusion step (K4) and the drift guidance (K5).
// Reference: /root/reference/models/decompdiff.py:219-256,295-297,315-338 (embeddings / heads),
// :601-689 (reverse step), models/transitions.py:65-161, utils/guidance_funcs.py:24-78.
#include "kernels.cuh"

namespace ddb {

// h[lig_idx[a]] = base[a] + Wv[v[a]] (+ Wt * t / T)   (ligand_atom_emb on [onehot(v) | aux (| t / T)] with the node indicator column;
// the time column is the 'simple' time embedding, models/decompdiff.py:224-230: t is the run's device-side time index, or one
// entry per graph when forward() was given explicit time steps)
__global__ void embed_ligand_kernel(const float* __restrict__ base, const float* __restrict__ Wv,
                                    const int64_t* __restrict__ v, int n, const int* __restrict__ lig_idx,
                                    float* __restrict__ h, const float* __restrict__ Wt, const int* __restrict__ t_dev,
                                    const int* __restrict__ t_graph, const int* __restrict__ graph_of_lig, float num_timesteps) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;     // one float4 per thread
  if (idx >= n * (H / 4)) return;
  int a = idx / (H / 4), c4 = idx - a * (H / 4);
  int vt = (int)v[a];
  float4 o = add4(ldg4(base + (size_t)a * H + c4 * 4), ldg4(Wv + (size_t)vt * H + c4 * 4));
  if (Wt != nullptr) {
    const int t = t_graph != nullptr ? t_graph[graph_of_lig[lig_idx[a]]] : *t_dev;      // graph_of_lig: graph of a merged node
    const float tf = __fdiv_rn((float)t, num_timesteps);      // time_step / self.num_timesteps
    o = fma4(tf, ldg4(Wt + c4 * 4), o);
  }
  st4(h + (size_t)lig_idx[a] * H + c4 * 4, o);
}
void launch_embed_ligand(const float* base, const float* Wv, const int64_t* v, int n, const int* lig_idx, float* h,
                         cudaStream_t stream, const float* Wt, const int* t_dev, const int* t_graph, const int* graph_of_lig,
                         int num_timesteps) {
  if (n <= 0) return;
  int total = n * (H / 4);
  embed_ligand_kernel<<<(total + 255) / 256, 256, 0, stream>>>(base, Wv, v, n, lig_idx, h, Wt, t_dev, t_graph, graph_of_lig,
                                                               (float)num_timesteps);
}

__global__ void embed_bond_kernel(const float* __restrict__ table, const int64_t* __restrict__ btype, int n_bonds,
                                  float* __restrict__ h_bond) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_bonds * (H / 4)) return;
  int e = idx / (H / 4), c4 = idx - e * (H / 4);
  st4(h_bond + (size_t)e * H + c4 * 4, ldg4(table + (size_t)btype[e] * H + c4 * 4));
}
void launch_embed_bond(const float* table, const int64_t* btype, int n_bonds, float* h_bond, cudaStream_t stream) {
  if (n_bonds <= 0) return;
  int total = n_bonds * (H / 4);
  embed_bond_kernel<<<(total + 255) / 256, 256, 0, stream>>>(table, btype, n_bonds, h_bond);
}

__global__ void set_ligand_x_kernel(const float* __restrict__ x_lig, int n, const int* __restrict__ lig_idx,
                                    float* __restrict__ x4) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  st4(x4 + (size_t)lig_idx[a] * 4, make_float4(x_lig[a * 3], x_lig[a * 3 + 1], x_lig[a * 3 + 2], 0.f));
}
void launch_set_ligand_x(const float* x_lig, int n, const int* lig_idx, float* x4, cudaStream_t stream) {
  if (n <= 0) return;
  set_ligand_x_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x_lig, n, lig_idx, x4);
}
__global__ void get_ligand_x_kernel(const float* __restrict__ x4, int n, const int* __restrict__ lig_idx,
                                    float* __restrict__ out) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n) return;
  float4 p = ld4(x4 + (size_t)lig_idx[a] * 4);
  out[a * 3] = p.x; out[a * 3 + 1] = p.y; out[a * 3 + 2] = p.z;
}
void launch_get_ligand_x(const float* x4, int n, const int* lig_idx, float* out, cudaStream_t stream) {
  if (n <= 0) return;
  get_ligand_x_kernel<<<(n + 255) / 256, 256, 0, stream>>>(x4, n, lig_idx, out);
}

// logits[r, :C] = hidden[r, :] @ W[C,128]^T + b     (second Linear of v_inference / bond_inference)
__global__ void __launch_bounds__(256) head_logits_kernel(const float* __restrict__ hidden, int ld, int rows,
                                                          const float* __restrict__ W, const float* __restrict__ b, int C,
                                                          float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float4 hv = ld4(hidden + (size_t)r * ld + lane * 4);
  for (int o = 0; o < C; ++o) {
    float s = warp_sum(dot4(hv, ldg4(W + (size_t)o * H + lane * 4)));
    if (lane == 0) logits[(size_t)r * C + o] = s + __ldg(b + o);
  }
}
void launch_head_logits(const float* hidden, int ld, int rows, const float* W, const float* b, int C, float* logits,
                        cudaStream_t stream) {
  if (rows <= 0) return;
  head_logits_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(hidden, ld, rows, W, b, C, logits);
}

// ------------------------------------------------------------------------------------ reverse step
__device__ __forceinline__ float log_add_exp(float a, float b) {   // transitions.py:91-93
  float m = fmaxf(a, b);
  return m + logf(expf(a - m) + expf(b - m));
}

template <int MAXC>
__device__ __forceinline__ int categorical_step(const float* __restrict__ logits, int C, int cur, int t, int tm1,
                                                const float* la, const float* l1ma, const float* lac, const float* l1mac,
                                                const float* prior, const float* __restrict__ u, float* recon_out,
                                                float* prob_out) {
  float l[MAXC], un[MAXC];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) if (c < C) { l[c] = logits[c]; mx = fmaxf(mx, l[c]); }
  float se = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) if (c < C) se += expf(l[c] - mx);
  const float lse = mx + logf(se);                       // F.log_softmax (decompdiff.py:617,629)
  const float log_eps = logf(1e-30f);                    // log(clamp(onehot, 1e-30))  (transitions.py:70)
  const float la_t = la[t], l1ma_t = l1ma[t], lac_p = lac[tm1], l1mac_p = l1mac[tm1];
  float umax = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) if (c < C) {
    float recon = l[c] - lse;
    if (recon_out) recon_out[c] = recon;
    float a = log_add_exp(recon + lac_p, l1mac_p + prior[c]);                     // q_v_pred at t-1  (:135-144)
    float b = log_add_exp((c == cur ? 0.f : log_eps) + la_t, l1ma_t + prior[c]);  // one step at t    (:123-133)
    un[c] = a + b;
    umax = fmaxf(umax, un[c]);
  }
  float us = 0.f;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) if (c < C) us += expf(un[c] - umax);
  const float ulse = umax + logf(us);                    // torch.logsumexp (:160)
  int best = 0; float bestv = -INFINITY;
#pragma unroll
  for (int c = 0; c < MAXC; ++c) if (c < C) {
    float p = un[c] - ulse;
    if (prob_out) prob_out[c] = p;
    float g = -logf(-logf(u[c] + 1e-30f) + 1e-30f);      // Gumbel noise (:79-81)
    float s = g + p;
    if (s > bestv) { bestv = s; best = c; }
  }
  return best;
}

__global__ void __launch_bounds__(256) reverse_step_kernel(const StepArgs a) {
  const int t = *a.t_dev;
  const int slot = *a.t_start_dev - t;
  const int tm1 = t - 1 < 0 ? 0 : t - 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < a.n_lig) {
    const int r = idx;
    const bool frozen = a.upd_mask != nullptr && a.upd_mask[r] == 0;
    float* v0 = a.v0_traj ? a.v0_traj + ((size_t)slot * a.n_lig + r) * a.C : nullptr;
    float* vt = a.vt_traj ? a.vt_traj + ((size_t)slot * a.n_lig + r) * a.C : nullptr;
    const int cur = (int)a.v[r];
    int nxt = categorical_step<16>(a.v_logits + (size_t)r * a.C, a.C, cur, t, tm1, a.a_log_alpha, a.a_log_1m_alpha,
                                   a.a_log_cumprod, a.a_log_1m_cumprod, a.a_prior, a.u_atom + (size_t)r * a.C, v0, vt);
    if (frozen) nxt = cur;
    // Gaussian posterior (:612-615, :679-683)
    const float c0 = a.c0[t], ct = a.ct[t];
    const float sig = (t == 0 ? 0.f : 1.f) * expf(0.5f * a.logvar[t]);
    float nx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float xt = a.x[r * 3 + d];
      float x0 = a.x0[r * 3 + d];
      // model_mean_type 'noise' (:602-605): eps = pred - x_t, x_0 = sqrt(1/acp) x_t - sqrt(1/acp - 1) eps  (two products, one
      // difference, as torch evaluates it - no contraction)
      if (a.recip) x0 = __fsub_rn(__fmul_rn(a.recip[t], xt), __fmul_rn(a.recipm1[t], __fsub_rn(x0, xt)));
      float mean = c0 * x0 + ct * xt;
      if (a.grad) mean -= a.grad[r * 3 + d];
      float v = mean + sig * a.eps[r * 3 + d] * a.prior_std[r * 3 + d];
      nx[d] = frozen ? xt : v;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      a.x[r * 3 + d] = nx[d];
      if (a.pos_traj) a.pos_traj[((size_t)slot * a.n_lig + r) * 3 + d] = nx[d] + a.offset_lig[r * 3 + d];
    }
    a.v[r] = nxt;
    if (a.v_traj) a.v_traj[(size_t)slot * a.n_lig + r] = nxt;
  } else if (idx - a.n_lig < a.n_bonds) {
    const int e = idx - a.n_lig;
    float* bt = a.bt_traj ? a.bt_traj + ((size_t)slot * a.n_bonds + e) * a.Cb : nullptr;
    const int cur = (int)a.bond[e];
    int nxt = categorical_step<8>(a.b_logits + (size_t)e * a.Cb, a.Cb, cur, t, tm1, a.b_log_alpha, a.b_log_1m_alpha,
                                  a.b_log_cumprod, a.b_log_1m_cumprod, a.b_prior, a.u_bond + (size_t)e * a.Cb, nullptr, bt);
    a.bond[e] = nxt;
    if (a.bond_traj) a.bond_traj[(size_t)slot * a.n_bonds + e] = nxt;
  }
}
void launch_reverse_step(const StepArgs& a, cudaStream_t stream) {
  int total = a.n_lig + a.n_bonds;
  if (total <= 0) return;
  reverse_step_kernel<<<(total + 255) / 256, 256, 0, stream>>>(a);
}

__global__ void advance_time_kernel(int* t) { *t -= 1; }
void launch_advance_time(int* t_dev, cudaStream_t stream) { advance_time_kernel<<<1, 1, 0, stream>>>(t_dev); }

// ---------------------------------------------------------------------------------------- guidance
// One CTA per complex.  grad = d/dx_t [ sum_g clash_g  +  (1/B) sum_g armsca_g ]   (decompdiff.py:638-677)
__global__ void __launch_bounds__(256) guidance_kernel(const GuidanceArgs a) {
  const int g = blockIdx.x;
  const int l0 = a.lig_ptr[g], l1 = a.lig_ptr[g + 1];
  const int n = l1 - l0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < n * 3; i += blockDim.x) a.grad[(size_t)l0 * 3 + i] = 0.f;
  __syncthreads();
  if (n == 0) return;
  const float coef_t = (a.scale_armsca || a.scale_clash) ? a.score_coef[*a.t_dev] : 1.f;
  if (a.enable_clash) {
    // compute_batch_clash_loss: per complex mean_i relu(gamma + sigma log(1e-3 + sum_j exp(-|p_j - x_i|^2 / sigma)))
    const int p0 = a.full_ptr[g], p1 = a.full_ptr[g + 1];
    const float inv_sigma = 1.0f / a.sigma;
    for (int i = warp; i < n; i += nwarps) {
      const int r = l0 + i;
      const float xo = a.x[r * 3] + a.offset_lig[r * 3], yo = a.x[r * 3 + 1] + a.offset_lig[r * 3 + 1],
                  zo = a.x[r * 3 + 2] + a.offset_lig[r * 3 + 2];
      float S = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
      for (int j = p0 + lane; j < p1; j += 32) {
        float4 p = ldg4(a.full_pos4 + (size_t)j * 4);
        float dx = p.x - xo, dy = p.y - yo, dz = p.z - zo;
        float e = expf(-(dx * dx + dy * dy + dz * dz) * inv_sigma);
        S += e; vx = fmaf(e, dx, vx); vy = fmaf(e, dy, vy); vz = fmaf(e, dz, vz);
      }
      S = warp_sum(S); vx = warp_sum(vx); vy = warp_sum(vy); vz = warp_sum(vz);
      if (lane == 0) {
        float G = -a.sigma * logf(1e-3f + S);
        if (a.gamma - G > 0.f) {
          float c = 2.0f / ((1e-3f + S) * (float)n) * (a.scale_clash ? coef_t : 1.f);
          a.grad[r * 3] += c * vx; a.grad[r * 3 + 1] += c * vy; a.grad[r * 3 + 2] += c * vz;
        }
      }
    }
  }
  __syncthreads();
  if (a.enable_armsca) {
    // compute_batch_armsca_prox_loss: mean over arm ids of hinge(min arm<->scaffold distance), / num_graphs
    __shared__ int s_narm, s_nsca;
    if (threadIdx.x == 0) {
      int mx = -1, nsca = 0;
      for (int i = 0; i < n; ++i) { int m = a.decomp_index[l0 + i]; if (m < 0) ++nsca; else mx = max(mx, m); }
      s_narm = mx + 1; s_nsca = nsca;
    }
    __syncthreads();
    const int narm = s_narm;
    if (narm > 0 && s_nsca > 0) {
      for (int arm = warp; arm < narm; arm += nwarps) {
        // lanes scan (arm atom, scaffold atom) pairs; keep the closest pair (lowest pair index on ties)
        float best = INFINITY; int bp = -1, bs = -1;
        for (int pi = 0; pi < n; ++pi) {
          if (a.decomp_index[l0 + pi] != arm) continue;
          for (int si = lane; si < n; si += 32) {
            if (a.decomp_index[l0 + si] >= 0) continue;
            float dx = a.x[(l0 + pi) * 3] - a.x[(l0 + si) * 3], dy = a.x[(l0 + pi) * 3 + 1] - a.x[(l0 + si) * 3 + 1],
                  dz = a.x[(l0 + pi) * 3 + 2] - a.x[(l0 + si) * 3 + 2];
            float d = sqrtf(dx * dx + dy * dy + dz * dz);
            if (d < best) { best = d; bp = pi; bs = si; }
          }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
          float ob = __shfl_xor_sync(FULL, best, m);
          int op = __shfl_xor_sync(FULL, bp, m), os = __shfl_xor_sync(FULL, bs, m);
          bool take = ob < best || (ob == best && op >= 0 && (bp < 0 || os < bs || (os == bs && op < bp)));
          if (take) { best = ob; bp = op; bs = os; }
        }
        if (lane == 0 && bp >= 0) {
          float dm = (a.min_d - best > 0.f ? -1.f : 0.f) + (best - a.max_d > 0.f ? 1.f : 0.f);
          if (dm != 0.f && best > 0.f) {
            float c = dm / (best * (float)narm * (float)a.num_graphs) * (a.scale_armsca ? coef_t : 1.f);
            int rp = l0 + bp, rs = l0 + bs;
            float dx = a.x[rp * 3] - a.x[rs * 3], dy = a.x[rp * 3 + 1] - a.x[rs * 3 + 1], dz = a.x[rp * 3 + 2] - a.x[rs * 3 + 2];
            atomicAdd(a.grad + rp * 3, c * dx); atomicAdd(a.grad + rp * 3 + 1, c * dy); atomicAdd(a.grad + rp * 3 + 2, c * dz);
            atomicAdd(a.grad + rs * 3, -c * dx); atomicAdd(a.grad + rs * 3 + 1, -c * dy); atomicAdd(a.grad + rs * 3 + 2, -c * dz);
          }
        }
      }
    }
  }
}
void launch_guidance(const GuidanceArgs& a, cudaStream_t stream) {
  if (a.num_graphs <= 0) return;
  guidance_kernel<<<a.num_graphs, 256, 0, stream>>>(a);
}

}  // namespace ddb
