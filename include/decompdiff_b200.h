/*
 * decompdiff_b200 - C ABI of the B200-native DecompDiff sampling hot path.
 *
 * The reference (bytedance/DecompDiff) is pure Python and has no FFI of its own
 * (SURVEY.md section 8b): the seam it offers is three Python-level contracts -
 *   (1) the model API      models/decompdiff.py:77,213,553  (DecompScorePosNet3D)
 *   (2) the refine-net API models/encoders/uni_transformer_edge.py:394
 *   (3) the collated-batch attribute contract scripts/sample_diffusion_decomp.py:317-352
 * This header is the plain-C boundary underneath them.  Every entry point takes
 * raw pointers and sizes only (no torch types), returns an int status
 * (0 = ok, see ddb_status) and never throws.  `ddb_last_error()` returns the text
 * of the last failure on the calling thread.
 *
 * Conventions
 *   - "host"   pointers are ordinary CPU memory, read during the call only.
 *   - "device" pointers are CUDA device memory on the current device; kernels are
 *     launched on the `stream` argument (a cudaStream_t passed as void*), nothing
 *     synchronises unless stated, so every per-step call is CUDA-graph capturable.
 *     ddb_forward / ddb_reverse_step fork part of each layer onto a stream the batch
 *     owns and join it back with events before they return: all of their work is
 *     ordered before whatever the caller enqueues on `stream` next, and under stream
 *     capture the fork becomes parallel branches of the caller's graph.
 *   - fp32 everywhere the reference is fp32; indices are int64 at the boundary
 *     exactly as the reference passes them (torch.long).
 */
#ifndef DECOMPDIFF_B200_H
#define DECOMPDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ddb_status {
  DDB_OK = 0,
  DDB_ERR_INVALID = 1,      /* bad argument / unsupported configuration (reference: ValueError) */
  DDB_ERR_MISSING = 2,      /* a required state_dict tensor was never set                        */
  DDB_ERR_CUDA = 3,         /* CUDA runtime failure; text in ddb_last_error()                    */
  DDB_ERR_STATE = 4         /* call order violated (e.g. forward before set_state)              */
} ddb_status;

typedef struct ddb_model ddb_model;   /* weights of one DecompScorePosNet3D                     */
typedef struct ddb_batch ddb_batch;   /* static topology + workspace + evolving state of a batch */

/* `model:` section of configs/training.yml:16-57 (only the fields that shape the path). */
typedef struct ddb_config {
  int32_t hidden_dim;          /* 128 (kernels are specialised for it)                        */
  int32_t n_heads;             /* 16                                                           */
  int32_t knn;                 /* 32  (<= 32)                                                  */
  int32_t num_layers;          /* 6                                                            */
  int32_t num_blocks;          /* 1                                                            */
  int32_t num_classes;         /* 8   ligand atom types ('basic')                              */
  int32_t num_bond_classes;    /* 5                                                            */
  int32_t protein_feature_dim; /* 29 = 27 + 2                                                  */
  int32_t ligand_feature_dim;  /* 10 = 8 + 2                                                   */
  int32_t num_timesteps;       /* 1000                                                         */
} ddb_config;

const char* ddb_last_error(void);
const char* ddb_version(void);

/* ---------------------------------------------------------------- model ------------------
 * Replaces: DecompScorePosNet3D.__init__ + load_state_dict(ckpt['model'], strict=True)
 * (scripts/sample_diffusion_decomp.py:537-544).  Tensors are handed over under their
 * reference state_dict names (616 keys, e.g.
 * "refine_net.base_block.0.bond_layer.hk_func.net.0.weight"), row-major fp32, host memory. */
int ddb_model_create(ddb_model** out, const ddb_config* cfg);
int ddb_model_set_tensor(ddb_model* m, const char* name, const float* host_data, int64_t numel);
/* Re-packs the weights into the kernels' layouts and uploads them.  Fails with
 * DDB_ERR_MISSING naming the first absent key (strict=True behaviour). */
int ddb_model_finalize(ddb_model* m);
void ddb_model_destroy(ddb_model* m);
/* Graph construction of the refine net (uni_transformer_edge.py:349-359): mode 0 = 'knn' (k nearest neighbours per node within
 * its graph, the shipped configuration), mode 1 = 'radius': the k nearest neighbours within r_max.  Upstream 'radius' raises
 * (it reads an attribute `self.r` that is never set, :351), so mode 1 is DEFINED here as radius_graph(r = r_max,
 * max_num_neighbors = k) with nearest-first truncation.  Mode 2 = 'hybrid' (batch_hybrid_edge_connection with add_p_index,
 * models/common.py:230-277): the ligand atoms of a complex are fully connected, every ligand atom also receives its k nearest
 * PROTEIN atoms (a complex with ligand atoms but fewer than k protein atoms is rejected with DDB_ERR_INVALID, as torch.topk raises
 * upstream), protein destinations keep their k nearest atoms of the whole complex.  A ligand destination then has n_lig - 1 + k
 * incoming edges, so this mode runs the kNN edge family on the fp32 FMA kernels with wide neighbour rows (no receptive-field
 * pruning, no neighbour cache).  Call before any batch is created. */
int ddb_model_set_cutoff(ddb_model* m, int32_t mode, float r_max);
/* model_mean_type (models/decompdiff.py:84, :602-611): 0 = 'C0' (the network output is x_0; the shipped configuration),
 * 1 = 'noise' (the network output minus x_t is the predicted noise, x_0 = sqrt(1/acp_t) x_t - sqrt(1/acp_t - 1) eps,
 * _predict_x0_from_eps :353-356).  Selects the branch of the posterior step in ddb_reverse_step; call before ddb_model_finalize
 * (the two extra schedule tables "sqrt_recip_alphas_cumprod" / "sqrt_recipm1_alphas_cumprod" are then required). */
int ddb_model_set_mean_type(ddb_model* m, int32_t noise);
/* Time embedding (models/decompdiff.py:168-183, 224-236): 0 = none (time_emb_dim = 0, the shipped configuration), 1 = 'simple'
 * (the ligand feature vector gets one more column, time_step / num_timesteps; "ligand_atom_emb.weight" then has
 * ligand_feature_dim = classes + aux + 1 input columns and ddb_config.ligand_feature_dim must say so).  The 'sin' mode cannot run
 * upstream for real batches (it concatenates a per-GRAPH embedding to per-ATOM features, :231-232) and is not offered.  Call
 * before ddb_model_finalize. */
int ddb_model_set_time_emb(ddb_model* m, int32_t mode);
/* Stand-alone refine net (get_refine_net('uni_o2_bond', config), models/encoders/__init__.py:27-43): call before
 * ddb_model_finalize; only the "refine_net.*" tensors are then required and the model serves ddb_refine_batch_create /
 * ddb_refine_forward only. */
int ddb_model_set_refine_only(ddb_model* m, int32_t on);

/* ---------------------------------------------------------------- batch ------------------
 * Replaces the per-call set-up of sample_diffusion / forward that does not depend on t:
 * center_pos (decompdiff.py:20-32,567), protein_atom_emb (:238), compose_context ordering
 * (common.py:167-194), bond index remap (:291), BondUpdateLayer.triplets
 * (uni_transformer_edge.py:103-123).  All pointers are HOST memory.
 *   batch_protein / batch_ligand : graph id per atom, ascending (PyG collate order)
 *   bond_index  : (2, n_bonds) row-major, [0]=src [1]=dst, ligand-atom numbering
 *   ligand_atom_mask : optional (NULL = all ones): 0 freezes an atom's position (decompdiff.py:285)
 *   center_mode : 0 = 'none', 1 = 'protein'                                                    */
int ddb_batch_create(ddb_batch** out, const ddb_model* m, int32_t num_graphs,
                     int64_t n_protein, const float* protein_pos, const float* protein_v,
                     const int64_t* batch_protein,
                     int64_t n_ligand, const int64_t* batch_ligand, const float* ligand_v_aux,
                     int64_t n_bonds, const int64_t* bond_index,
                     const uint8_t* ligand_atom_mask, int32_t center_mode);
void ddb_batch_destroy(ddb_batch* b);
/* per-graph offset subtracted by center_pos, (num_graphs,3) host out */
int ddb_batch_get_offset(const ddb_batch* b, float* offset_out);

/* Evolving state (x_t, v_t, b_t).  DEVICE pointers; positions are in the caller's
 * (un-centred) frame - the batch applies / removes the centring offset itself. */
int ddb_batch_set_state(ddb_batch* b, const float* ligand_pos, const int64_t* ligand_v,
                        const int64_t* bond_type, void* stream);
int ddb_batch_get_state(const ddb_batch* b, float* ligand_pos, int64_t* ligand_v,
                        int64_t* bond_type, void* stream);

/* ---------------------------------------------------------------- forward ----------------
 * Replaces DecompScorePosNet3D.forward (decompdiff.py:213-351) on the batch's current state:
 * out_pos (n_ligand,3) centred frame as in the reference, out_v_logits (n_ligand,num_classes),
 * out_bond_logits (n_bonds,num_bond_classes).  DEVICE pointers, any may be NULL.              */
int ddb_forward(ddb_batch* b, float* out_pos, float* out_v_logits, float* out_bond_logits,
                void* stream);

/* Same with return_all=True (decompdiff.py:343-350): out_v_logits_input (n_ligand,num_classes) receives v_inference of the
 * INPUT ligand embedding - the reference's 'layer_pred_ligand_v'[0]; its [1] is out_v_logits ('layer_pred_ligand_pos' is
 * [input positions, out_pos]; the lists have one entry per BLOCK boundary, uni_transformer_edge.py:436-438, num_blocks = 1). */
int ddb_forward_ex(ddb_batch* b, float* out_pos, float* out_v_logits, float* out_bond_logits,
                   float* out_v_logits_input, void* stream);

/* ---------------------------------------------------------------- refine-net seam ---------
 * Replaces UniTransformerO2TwoUpdateGeneralBond.forward(h, x, group_idx, bond_index, h_bond, mask_ligand, mask_ligand_atom,
 * batch) (uni_transformer_edge.py:394-443) - the narrowest swap point of the reference (contract 2).
 * ddb_refine_batch_create: HOST pointers; nodes in merged order (batch ascending, within a graph protein nodes before ligand
 * nodes - what compose_context produces, common.py:172-191); mask_ligand / mask_ligand_atom (n_nodes) uint8 (the latter may be
 * NULL = mask_ligand); bond_index (2, n_bonds) in MERGED node numbering as the reference passes it (decompdiff.py:291).
 * ddb_refine_forward: DEVICE pointers; h (n_nodes,128), x (n_nodes,3), h_bond (n_bonds,128) in, the same shapes out
 * ({'h','x','h_bond'} of the reference; every node row is computed - no receptive-field pruning on this path). */
int ddb_refine_batch_create(ddb_batch** out, const ddb_model* m, int32_t num_graphs, int64_t n_nodes,
                            const int64_t* batch, const uint8_t* mask_ligand, const uint8_t* mask_ligand_atom,
                            int64_t n_bonds, const int64_t* bond_index);
int ddb_refine_forward(ddb_batch* b, const float* h, const float* x, const float* h_bond,
                       float* h_out, float* x_out, float* h_bond_out, void* stream);

/* ---------------------------------------------------------------- reverse step -----------
 * Replaces one iteration of the loop in sample_diffusion (decompdiff.py:576-689): forward,
 * Gaussian posterior, categorical posteriors + Gumbel-argmax, optional drift, x_{t-1}.
 *   prior_std_atom : (n_ligand,3) device, prior_stds[ligand_decomp_batch]
 *   u_atom (n_ligand,num_classes), u_bond (n_bonds,num_bond_classes) : U[0,1) draws,
 *   eps_pos (n_ligand,3) : N(0,1) draws - the three draws the reference makes per step, in
 *   its order (transitions.py:79 via decompdiff.py:620, :633, :680).
 * The time index lives on the device (ddb_batch_set_time) so one captured CUDA graph can be
 * replayed for every step; each call uses t and then decrements it.
 * Trajectory outputs (device, any may be NULL), all written at row `step slot` =
 * (t_start - t): pos_traj (S,n,3) un-centred, v_traj (S,n) i64, v0_traj / vt_traj (S,n,C),
 * bond_traj (S,Eb) i64, bt_traj (S,Eb,Cb).                                                     */
typedef struct ddb_step_io {
  const float* prior_std_atom;
  const float* u_atom;
  const float* u_bond;
  const float* eps_pos;
  float* pos_traj; int64_t* v_traj; float* v0_traj; float* vt_traj;
  int64_t* bond_traj; float* bt_traj;
} ddb_step_io;
int ddb_batch_set_time(ddb_batch* b, int32_t t_start, void* stream);
/* forward() with explicit time steps (DecompScorePosNet3D.forward(..., time_step), one int64 per graph, host memory): only the
 * 'simple' time embedding reads them.  ddb_batch_set_time returns the batch to the run's single device-side time index. */
int ddb_batch_set_time_steps(ddb_batch* b, const int64_t* time_step, void* stream);
int ddb_reverse_step(ddb_batch* b, const ddb_step_io* io, void* stream);

/* Drift guidance (decompdiff.py:638-677, utils/guidance_funcs.py:24-78); host pointers, copied.
 *   armsca_prox: ligand_decomp_index (n_ligand) arm id or -1, min_d/max_d
 *   clash      : full protein cloud in the un-centred frame, sigma / gamma (surface_ct)
 * Pass enable_* = 0 to switch a term off.                                                      */
int ddb_batch_set_guidance(ddb_batch* b,
                           int32_t enable_armsca, const int64_t* ligand_decomp_index,
                           float min_d, float max_d,
                           int32_t enable_clash, int64_t n_full, const float* full_protein_pos,
                           const int64_t* full_batch_protein, float sigma, float gamma);
/* drift option `scale: True` (decompdiff.py:657-658, :668-669): multiply the armsca_prox / clash gradient by
 * pos_score_coef[t] (the coefficient goes to 0 as t -> 0).  Off by default (the shipped sampling_drift.yml has no `scale`). */
int ddb_batch_set_guidance_scale(ddb_batch* b, int32_t scale_armsca, int32_t scale_clash);

/* ---------------------------------------------------------------- building blocks --------
 * Stand-alone entry points for the kernels (unit-testable seams; DEVICE pointers).            */
/* kNN graph, replaces torch_geometric.nn.knn_graph at uni_transformer_edge.py:353.
 * x4: (n,4) fp32 (xyz + pad); node_ptr: (num_graphs+1) int32 CSR of nodes per graph;
 * is_ligand: (n) uint8.  Outputs: nbr (n,K) int32 source node ids (ligand sources first,
 * then by ascending distance, ties by index), deg (n) int32, nlig (n) int32.                  */
int ddb_knn_graph(const float* x4, const int32_t* node_ptr, const uint8_t* is_ligand,
                  int32_t num_graphs, int32_t n, int32_t k,
                  int32_t* nbr, int32_t* deg, int32_t* nlig, void* stream);
/* C[M,N] = act(A[M,128] @ Wt[128,N] + bias): the GEMM used for all node / edge projections.
 * impl 0 = fp32 FMA kernel, 1 = tcgen05 / TMEM 3xTF32 kernel (synchronises; test seam).      */
int ddb_gemm128(const float* A, int32_t lda, const float* Wt, int32_t ldw, const float* bias,
                float* C, int32_t ldc, int32_t M, int32_t N, int32_t act, int32_t impl, void* stream);

/* Introspection for tests / profiling: device pointers to internal buffers of the last forward.
 * name in {"h","x","h_bond","nbr","deg","nlig","e_w","grad"}; rows/cols describe the layout.          */
int ddb_batch_debug_buffer(const ddb_batch* b, const char* name, const void** ptr,
                           int64_t* rows, int64_t* cols);
/* Per-layer intermediates (test seam): DEVICE buffers (num_layers, n_nodes, 128), (num_layers, n_nodes, 4) [xyz + pad]
 * and (num_layers, n_bonds, 128) that receive h / x / h_bond after every layer of the following forwards; NULL disables a tap.
 * With receptive-field pruning (sampling batches) protein rows outside a layer's receptive field hold stale values. */
int ddb_batch_set_layer_tap(ddb_batch* b, float* h_layers, float* x_layers, float* h_bond_layers);
/* Per-kernel device timing of eager (non-captured) forward / reverse-step calls: CUDA events recorded on the launch
 * stream around every kernel, accumulated per category.  enable=1 makes each call end with a stream synchronise. */
int ddb_batch_profile(ddb_batch* b, int32_t enable, int32_t reset);
int32_t ddb_profile_num_categories(void);
const char* ddb_profile_category_name(int32_t i);
int ddb_batch_profile_read(const ddb_batch* b, double* ms_out, int64_t* count_out);
/* stream-ordered device-to-device copy (lets a host language read a debug buffer without its own CUDA binding) */
int ddb_copy_device(void* dst, const void* src, int64_t bytes, void* stream);
/* number of kernel launches issued by the last ddb_forward / ddb_reverse_step on this batch    */
int64_t ddb_batch_last_launch_count(const ddb_batch* b);
/* kNN-attention destination rows of the last forward: `full` = num_layers * nodes (what the reference computes,
 * uni_transformer_edge.py:259-287 on every node), `executed` = what ran after the exact receptive-field pruning and the
 * first-layer cache.  Synchronous (reads device counters). */
int ddb_batch_executed_rows(const ddb_batch* b, int64_t* executed, int64_t* full);
/* bytes copied host->device by ddb_batch_create / ddb_batch_set_guidance for this batch */
int64_t ddb_batch_h2d_bytes(const ddb_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* DECOMPDIFF_B200_H */
