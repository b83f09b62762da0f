"""TEST INFRASTRUCTURE ONLY - the CPU oracle for the DecompDiff sampling hot path.

A plain torch-fp32 (CPU) restatement of the reference algorithm in the *reference formulation*
(materialised per-edge / per-triplet inputs, scatter-softmax, scatter-sum), written from the
reference's behaviour and citing the lines each function follows.  It is a pure function of a
reference-format `state_dict` (616 keys) and the tensors the reference API takes.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this module; the product path (`decompdiff_b200/`) never does and fails loudly when its
CUDA library is missing.

PARITY PIN: the reference ships no tests or golden vectors.  This restatement
is pinned against the UNMODIFIED reference executed in the build container through
`oracle/ref_shims.py`; the committed fixtures `tests/golden/*.pt` were produced by
`oracle/make_golden.py` from the reference itself, and `tests/test_oracle_golden.py` checks this
file against them (max |err| ~1e-6, i.e. fp32 re-association noise).  Schedule known-answers from
SURVEY.md section 4 are checked in `tests/test_schedules.py`.

All citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .ref_shims import knn_graph, scatter_softmax, scatter_sum, scatter_mean, scatter_min

GAUSS_OFFSETS = [0, 1, 1.25, 1.5, 1.75, 2, 2.25, 2.5, 2.75, 3, 3.5, 4, 4.5, 5, 5.5, 6, 7, 8, 9, 10]
SSP_SHIFT = math.log(2.0)


# ----------------------------------------------------------------------------
# schedules (models/decompdiff.py:96-131, models/transitions.py:12-28,55-57,97-120)
# ----------------------------------------------------------------------------
def cosine_alpha_schedule(timesteps: int, s: float) -> np.ndarray:
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    alphas = np.clip(ac[1:] / ac[:-1], a_min=0.001, a_max=1.)
    return np.sqrt(alphas)


def schedule_tables(cfg) -> Dict[str, torch.Tensor]:
    """All constant tables the reference keeps as non-trainable Parameters, same key names."""
    T = cfg['num_diffusion_timesteps']
    if cfg['beta_schedule'] == 'sigmoid':
        b = np.linspace(-6, 6, T)
        betas = 1 / (np.exp(-b) + 1) * (cfg['beta_end'] - cfg['beta_start']) + cfg['beta_start']
        alphas = 1. - betas
    elif cfg['beta_schedule'] == 'cosine':
        alphas = cosine_alpha_schedule(T, cfg['pos_beta_s']) ** 2
        betas = 1. - alphas
    elif cfg['beta_schedule'] == 'linear':
        betas = np.linspace(cfg['beta_start'], cfg['beta_end'], T, dtype=np.float64)
        alphas = 1. - betas
    else:
        raise NotImplementedError(cfg['beta_schedule'])
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1., ac[:-1])
    post_var = betas * (1. - ac_prev) / (1. - ac)
    f32 = lambda a: torch.from_numpy(np.asarray(a)).float()
    out = {
        'betas': f32(betas), 'alphas_cumprod': f32(ac), 'alphas_cumprod_prev': f32(ac_prev),
        'sqrt_alphas_cumprod': f32(np.sqrt(ac)),
        'sqrt_one_minus_alphas_cumprod': f32(np.sqrt(1. - ac)),
        'sqrt_recip_alphas_cumprod': f32(np.sqrt(1. / ac)),
        'sqrt_recipm1_alphas_cumprod': f32(np.sqrt(1. / ac - 1)),
        'posterior_mean_c0_coef': f32(betas * np.sqrt(ac_prev) / (1. - ac)),
        'posterior_mean_ct_coef': f32((1. - ac_prev) * np.sqrt(alphas) / (1. - ac)),
        'posterior_var': f32(post_var),
        'pos_score_coef': f32(betas / np.sqrt(alphas)),
    }
    # decompdiff.py:130 - log of the *fp32* posterior variance with entry 0 replaced by entry 1
    pv32 = out['posterior_var'].numpy()
    out['posterior_logvar'] = f32(np.log(np.append(pv32[1], pv32[1:])))
    for name, K in (('atom_type_trans', cfg['num_classes']), ('bond_type_trans', cfg['num_bond_classes'])):
        la = np.log(cosine_alpha_schedule(T, cfg['v_beta_s']))
        lac = np.cumsum(la)
        l1m = lambda a: np.log(1 - np.exp(a) + 1e-40)
        out[f'{name}.log_alphas_v'] = f32(la)
        out[f'{name}.log_one_minus_alphas_v'] = f32(l1m(la))
        out[f'{name}.log_alphas_cumprod_v'] = f32(lac)
        out[f'{name}.log_one_minus_alphas_cumprod_v'] = f32(l1m(lac))
        out[f'{name}.prior_probs'] = f32(-np.log(K).repeat(K)[None, :])
    return out


# ----------------------------------------------------------------------------
# small blocks (models/common.py)
# ----------------------------------------------------------------------------
def mlp(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Linear -> LayerNorm(eps 1e-5) -> ReLU -> Linear (common.py:85-105)."""
    z = F.linear(x, sd[f'{prefix}.net.0.weight'], sd[f'{prefix}.net.0.bias'])
    z = F.layer_norm(z, (z.size(-1),), sd[f'{prefix}.net.1.weight'], sd[f'{prefix}.net.1.bias'], 1e-5)
    return F.linear(F.relu(z), sd[f'{prefix}.net.3.weight'], sd[f'{prefix}.net.3.bias'])


def gaussian_smearing(d: torch.Tensor) -> torch.Tensor:
    """Fixed 20 offsets, coeff = -0.5/(o1-o0)^2 = -0.5 (common.py:16-31)."""
    off = torch.tensor(GAUSS_OFFSETS, dtype=d.dtype, device=d.device)
    return torch.exp(-0.5 * (d.reshape(-1, 1) - off.view(1, -1)) ** 2)


def angular_encoding(theta: torch.Tensor) -> torch.Tensor:
    """[theta, sin(f theta), cos(f theta)], f = [1,2,3,1,1/2,1/3] (common.py:34-54)."""
    f = torch.tensor([1., 2., 3., 1., 1. / 2, 1. / 3], dtype=theta.dtype, device=theta.device)
    t = theta.unsqueeze(-1)
    return torch.cat([t, torch.sin(t * f), torch.cos(t * f)], dim=-1)


def shifted_softplus(x):
    return F.softplus(x) - SSP_SHIFT


def compose_context(h_p, h_l, pos_p, pos_l, batch_p, batch_l, ligand_atom_mask=None):
    """Stable sort of [protein; ligand] by graph id (common.py:167-194)."""
    batch_ctx = torch.cat([batch_p, batch_l], 0)
    sort_idx = torch.sort(batch_ctx, stable=True).indices
    n_p, n_l = batch_p.numel(), batch_l.numel()
    mask_ligand = torch.cat([torch.zeros(n_p, dtype=torch.bool), torch.ones(n_l, dtype=torch.bool)])[sort_idx]
    if ligand_atom_mask is None:
        mask_ligand_atom = mask_ligand
    else:
        mask_ligand_atom = torch.cat([torch.zeros(n_p, dtype=torch.bool), ligand_atom_mask.bool()])[sort_idx]
    inv = torch.empty_like(sort_idx)
    inv[sort_idx] = torch.arange(sort_idx.numel())
    l_index_in_ctx = inv[n_p:]  # position of every ligand atom in the merged order
    return (torch.cat([h_p, h_l], 0)[sort_idx], torch.cat([pos_p, pos_l], 0)[sort_idx],
            batch_ctx[sort_idx], mask_ligand, mask_ligand_atom, l_index_in_ctx)


# ----------------------------------------------------------------------------
# score network (models/encoders/uni_transformer_edge.py)
# ----------------------------------------------------------------------------
def node_update(sd, p, h, edge_feat, src, dst, e_w, n_heads):
    """NodeUpdateLayer.forward (:42-74), out_fc off."""
    N = h.size(0)
    kv = torch.cat([edge_feat, h[dst], h[src]], -1)
    k = mlp(sd, f'{p}.hk_func', kv).view(-1, n_heads, 128 // n_heads)
    v = mlp(sd, f'{p}.hv_func', kv)
    if e_w is not None:
        v = v * e_w.view(-1, 1)
    v = v.view(-1, n_heads, 128 // n_heads)
    q = mlp(sd, f'{p}.hq_func', h).view(-1, n_heads, 128 // n_heads)
    alpha = scatter_softmax((q[dst] * k / np.sqrt(k.shape[-1])).sum(-1), dst, dim=0)
    out = scatter_sum(alpha.unsqueeze(-1) * v, dst, dim=0, dim_size=N)
    return out.view(-1, 128)


def pos_update(sd, p, h, rel_x, edge_feat, src, dst, e_w, n_heads):
    """PosUpdateLayer.forward (:188-210)."""
    N = h.size(0)
    kv = torch.cat([edge_feat, h[dst], h[src]], -1)
    k = mlp(sd, f'{p}.xk_func', kv).view(-1, n_heads, 128 // n_heads)
    v = mlp(sd, f'{p}.xv_func', kv)
    if e_w is not None:
        v = v * e_w.view(-1, 1)
    v = v.unsqueeze(-1) * rel_x.unsqueeze(1)
    q = mlp(sd, f'{p}.xq_func', h).view(-1, n_heads, 128 // n_heads)
    alpha = scatter_softmax((q[dst] * k / np.sqrt(k.shape[-1])).sum(-1), dst, dim=0)
    out = scatter_sum(alpha.unsqueeze(-1) * v, dst, dim=0, dim_size=N)
    return out.mean(1)


def bond_triplets(bsrc: torch.Tensor, bdst: torch.Tensor, num_nodes: int):
    """All (k->j, j->i) edge pairs with k != i (BondUpdateLayer.triplets, :103-123).
    Returns (idx_i, idx_j, idx_k, idx_kj, idx_ji); grouped by ji, k ascending."""
    Eb = bsrc.numel()
    order = torch.sort(bdst * num_nodes + bsrc, stable=True).indices  # edges sorted by (dst, src)
    counts = torch.bincount(bdst, minlength=num_nodes)
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    # for edge e = (j -> i): all edges entering j
    j = bsrc
    start, length = ptr[j], counts[j]
    total = int(length.sum())
    idx_ji = torch.repeat_interleave(torch.arange(Eb), length)
    offs = torch.arange(total) - torch.repeat_interleave(
        torch.cat([length.new_zeros(1), length.cumsum(0)[:-1]]), length)
    idx_kj = order[torch.repeat_interleave(start, length) + offs]
    idx_k = bsrc[idx_kj]
    idx_i, idx_j = bdst[idx_ji], bsrc[idx_ji]
    keep = idx_i != idx_k
    return idx_i[keep], idx_j[keep], idx_k[keep], idx_kj[keep], idx_ji[keep]


def bond_update(sd, p, h, h_bond, pos, bsrc, bdst, n_heads):
    """BondUpdateLayer.forward with include_h_node=True (:125-167)."""
    Eb = h_bond.size(0)
    idx_i, idx_j, idx_k, idx_kj, idx_ji = bond_triplets(bsrc, bdst, h.size(0))
    dist = (pos[bdst] - pos[bsrc]).pow(2).sum(-1).sqrt()
    pos_i = pos[idx_i]
    pos_ji, pos_ki = pos[idx_j] - pos_i, pos[idx_k] - pos_i
    a = (pos_ji * pos_ki).sum(-1)
    b = torch.linalg.cross(pos_ji, pos_ki, dim=-1).norm(dim=-1)
    angle = torch.atan2(b, a)
    r_feat = gaussian_smearing(dist)
    a_feat = angular_encoding(angle)
    kv = torch.cat([h_bond[idx_kj], r_feat[idx_kj], r_feat[idx_ji], a_feat, h[idx_k], h[idx_j]], -1)
    q_in = torch.cat([h_bond[idx_ji], h[idx_i]], -1)
    k = mlp(sd, f'{p}.hk_func', kv).view(-1, n_heads, 128 // n_heads)
    v = mlp(sd, f'{p}.hv_func', kv).view(-1, n_heads, 128 // n_heads)
    q = mlp(sd, f'{p}.hq_func', q_in).view(-1, n_heads, 128 // n_heads)
    alpha = scatter_softmax((q * k / np.sqrt(k.shape[-1])).sum(-1), idx_ji, dim=0, dim_size=Eb)
    out = scatter_sum(alpha.unsqueeze(-1) * v, idx_ji, dim=0, dim_size=Eb)
    return out.view(-1, 128)


def edge_types(src, dst, mask_ligand):
    """_build_edge_type (:361-377): 0 l->l, 1 l->p, 2 p->l, 3 p->p (src kind first)."""
    n_src, n_dst = mask_ligand[src], mask_ligand[dst]
    t = torch.full_like(src, 3)
    t[n_src & n_dst] = 0
    t[n_src & ~n_dst] = 1
    t[~n_src & n_dst] = 2
    return t


def attention_layer(sd, p, h, x, edge_type_1h, src, dst, h_bond, bsrc, bdst, mask_upd, e_w, n_heads):
    """AttentionLayerO2TwoUpdateNodeGeneral.forward (:259-287)."""
    rel_x = x[dst] - x[src]
    dist = torch.norm(rel_x, p=2, dim=-1, keepdim=True)
    g = gaussian_smearing(dist)
    dist_feat = (edge_type_1h.unsqueeze(-1) * g.unsqueeze(1)).reshape(g.size(0), -1)  # idx = type*20+g
    edge_feat = torch.cat([dist_feat, edge_type_1h], -1)
    new_h_edge = node_update(sd, f'{p}.node_layer_with_edge', h, edge_feat, src, dst, e_w, n_heads)
    new_h_bond_nodes = node_update(sd, f'{p}.node_layer_with_bond', h, h_bond, bsrc, bdst, None, n_heads)
    new_h_bond = h_bond + bond_update(sd, f'{p}.bond_layer', h, h_bond, x, bsrc, bdst, n_heads)
    new_h = h + F.linear(new_h_edge + new_h_bond_nodes, sd[f'{p}.lin_node.weight'], sd[f'{p}.lin_node.bias'])
    dx_edge = pos_update(sd, f'{p}.pos_layer_with_edge', new_h, rel_x, edge_feat, src, dst, e_w, n_heads)
    rel_bond_x = x[bdst] - x[bsrc]
    dx_bond = pos_update(sd, f'{p}.pos_layer_with_bond', new_h, rel_bond_x, new_h_bond, bsrc, bdst, None, n_heads)
    x = x + (dx_edge + dx_bond) * mask_upd[:, None]
    return new_h, new_h_bond, x


def hybrid_edges(x, k, mask_ligand, batch):
    """batch_hybrid_edge_connection(add_p_index=True) (models/common.py:230-277): per complex, ligand atoms fully connected; every
    ligand atom receives its k nearest PROTEIN atoms (torch.topk, :241-242); protein destinations receive their k nearest atoms
    of the whole complex (knn_graph over [protein; ligand], edges into ligand atoms dropped, :262-268)."""
    out = []
    for g in range(int(batch.max()) + 1):
        lig = ((batch == g) & mask_ligand).nonzero()[:, 0]
        pro = ((batch == g) & ~mask_ligand).nonzero()[:, 0]
        dst = torch.repeat_interleave(lig, len(lig))
        src = lig.repeat(len(lig))
        keep = dst != src
        ll = torch.stack([src[keep], dst[keep]])
        d = torch.norm(x[lig].unsqueeze(1) - x[pro].unsqueeze(0), p=2, dim=-1)
        near = torch.topk(d, k=k, largest=False, dim=1).indices          # raises when the complex has fewer than k protein atoms
        pl = torch.stack([pro[near], lig.unsqueeze(1).repeat(1, k)]).view(2, -1)
        all_idx = torch.cat([pro, lig])
        pe = knn_graph(x[all_idx], k=k)
        pe = pe[:, pe[1] < len(pro)]
        out.append(torch.cat([ll, pl, torch.stack([all_idx[pe[0]], all_idx[pe[1]]])], -1))
    return torch.cat(out, -1)


def refine_net(sd, cfg, h, x, bond_index, h_bond, mask_ligand, mask_ligand_atom, batch,
               return_all=False, knn_edge_index=None):
    """UniTransformerO2TwoUpdateGeneralBond.forward (:394-443); cutoff_mode 'knn' (shipped), 'hybrid', and the product's 'radius'."""
    p = 'refine_net'
    n_heads = cfg['n_heads']
    all_x, all_h, all_hb = [x], [h], [h_bond]
    edge_index = None
    for _ in range(cfg['num_blocks']):
        if cfg.get('cutoff_mode', 'knn') == 'hybrid':
            edge_index = hybrid_edges(x, cfg['knn'], mask_ligand.bool(), batch)
        else:
            edge_index = knn_graph(x, k=cfg['knn'], batch=batch) if knn_edge_index is None else knn_edge_index
        if cfg.get('cutoff_mode', 'knn') == 'radius':
            # upstream raises here (`self.r` undefined, :351); the product DEFINES the mode as the k nearest neighbours within r_max
            # (include/decompdiff_b200.h: ddb_model_set_cutoff) and this restates that definition
            s_, d_ = edge_index
            dd = x[d_] - x[s_]
            d2 = (dd[:, 0] * dd[:, 0] + dd[:, 1] * dd[:, 1]) + dd[:, 2] * dd[:, 2]
            keep = d2 <= torch.tensor(cfg['r_max'], dtype=torch.float32) ** 2
            edge_index = torch.stack([s_[keep], d_[keep]])
        src, dst = edge_index
        et = F.one_hot(edge_types(src, dst, mask_ligand), num_classes=4).to(h.dtype)
        dist = torch.norm(x[dst] - x[src], p=2, dim=-1, keepdim=True)
        e_w = torch.sigmoid(mlp(sd, f'{p}.edge_pred_layer', gaussian_smearing(dist)))
        for l in range(cfg['num_layers']):
            h, h_bond, x = attention_layer(sd, f'{p}.base_block.{l}', h, x, et, src, dst, h_bond,
                                           bond_index[0], bond_index[1], mask_ligand_atom.to(x.dtype), e_w, n_heads)
            if return_all:
                all_x.append(x), all_h.append(h), all_hb.append(h_bond)
    out = {'x': x, 'h': h, 'h_bond': h_bond, 'edge_index': edge_index}
    if return_all:
        out.update(all_x=all_x, all_h=all_h, all_h_bond=all_hb)
    return out


# ----------------------------------------------------------------------------
# DecompScorePosNet3D.forward (models/decompdiff.py:213-351)
# ----------------------------------------------------------------------------
def forward(sd, cfg, protein_pos, protein_v, batch_protein, init_ligand_pos, init_ligand_v,
            init_ligand_v_aux, batch_ligand, ligand_fc_bond_index, init_ligand_fc_bond_type,
            ligand_atom_mask=None, return_all=False, knn_edge_index=None, time_step=None, **_unused):
    nc, nb = cfg['num_classes'], cfg['num_bond_classes']
    lig_feat = torch.cat([F.one_hot(init_ligand_v, nc).float(), init_ligand_v_aux], -1)
    if cfg.get('time_emb_dim', 0) > 0:      # 'simple' time embedding (decompdiff.py:224-229)
        if cfg.get('time_emb_mode', 'simple') != 'simple':
            raise NotImplementedError
        lig_feat = torch.cat([lig_feat, (time_step / cfg['num_diffusion_timesteps'])[batch_ligand].unsqueeze(-1)], -1)
    h_p = F.linear(protein_v, sd['protein_atom_emb.weight'], sd['protein_atom_emb.bias'])
    h_l = F.linear(lig_feat, sd['ligand_atom_emb.weight'], sd['ligand_atom_emb.bias'])
    h_p = torch.cat([h_p, torch.zeros(h_p.size(0), 1)], -1)   # node indicator 0 / 1 (:252-256)
    h_l = torch.cat([h_l, torch.ones(h_l.size(0), 1)], -1)
    h, x, batch, mask_l, mask_la, l_idx = compose_context(
        h_p, h_l, protein_pos, init_ligand_pos, batch_protein, batch_ligand, ligand_atom_mask)
    bond_index = l_idx[ligand_fc_bond_index]
    h_bond = F.linear(F.one_hot(init_ligand_fc_bond_type, nb).float(),
                      sd['ligand_bond_emb.weight'], sd['ligand_bond_emb.bias'])
    out = refine_net(sd, cfg, h, x, bond_index, h_bond, mask_l, mask_la, batch,
                     return_all=return_all, knn_edge_index=knn_edge_index)
    fh = out['h'][mask_la]
    v_logits = F.linear(shifted_softplus(F.linear(fh, sd['v_inference.0.weight'], sd['v_inference.0.bias'])),
                        sd['v_inference.2.weight'], sd['v_inference.2.bias'])
    b_logits = F.linear(shifted_softplus(F.linear(out['h_bond'], sd['bond_inference.0.weight'],
                                                  sd['bond_inference.0.bias'])),
                        sd['bond_inference.2.weight'], sd['bond_inference.2.bias'])
    preds = {'pred_ligand_pos': out['x'][mask_la], 'pred_ligand_v': v_logits, 'pred_bond': b_logits}
    if return_all:
        preds.update(all_x=out['all_x'], all_h=out['all_h'], all_h_bond=out['all_h_bond'],
                     edge_index=out['edge_index'], l_index_in_ctx=l_idx, mask_ligand=mask_l)
    return preds


# ----------------------------------------------------------------------------
# categorical transitions (models/transitions.py)
# ----------------------------------------------------------------------------
def index_to_log_onehot(x, K):
    return torch.log(F.one_hot(x, K).float().clamp(min=1e-30))  # :65-71


def log_add_exp(a, b):
    m = torch.max(a, b)
    return m + torch.log(torch.exp(a - m) + torch.exp(b - m))  # :91-93


def q_v_posterior(tab, name, log_v0, log_vt, t, batch):
    """DiscreteTransition.q_v_posterior (:153-161) with q_v_pred (:135-144), one-step (:123-133)."""
    ex = lambda key, tt: tab[f'{name}.{key}'][tt][batch].unsqueeze(-1)
    prior = tab[f'{name}.prior_probs']
    tm1 = torch.where(t - 1 < 0, torch.zeros_like(t), t - 1)
    log_qvt1_v0 = log_add_exp(log_v0 + ex('log_alphas_cumprod_v', tm1),
                              ex('log_one_minus_alphas_cumprod_v', tm1) + prior)
    one_step = log_add_exp(log_vt + ex('log_alphas_v', t), ex('log_one_minus_alphas_v', t) + prior)
    un = log_qvt1_v0 + one_step
    return un - torch.logsumexp(un, dim=-1, keepdim=True)


def gumbel_argmax(logits, uniform):
    """log_sample_categorical (:78-84) with the uniform draw passed in."""
    g = -torch.log(-torch.log(uniform + 1e-30) + 1e-30)
    return (g + logits).argmax(dim=-1)


# ----------------------------------------------------------------------------
# drift guidance (utils/guidance_funcs.py:24-78, models/decompdiff.py:638-677)
# ----------------------------------------------------------------------------
def armsca_prox_energy(pos, batch_ligand, decomp_index, min_d, max_d, num_graphs_div=None):
    """`num_graphs_div`: the batch size the energy is divided by (:78) when `pos` holds only some pockets of a larger batch."""
    total = torch.tensor(0.)
    num_graphs = int(batch_ligand.max()) + 1
    n_valid = 0
    for g in range(num_graphs):
        sel = batch_ligand == g
        p, m = pos[sel], decomp_index[sel]
        arm = m != -1
        arm_pos, sca_pos = p[arm], p[~arm]
        if len(arm_pos) > 0 and len(sca_pos) > 0:
            pd = torch.norm(arm_pos.unsqueeze(1) - sca_pos.unsqueeze(0), p=2, dim=-1)
            min_all, _ = scatter_min(pd, m[arm], dim=0)
            md, _ = min_all.min(-1)
            total = total + torch.mean(torch.clamp(min_d - md, min=0) + torch.clamp(md - max_d, min=0))
            n_valid += 1
    return total / (num_graphs_div or num_graphs), n_valid  # the 1/num_graphs quirk of :78


def clash_energy(full_protein_pos, lig_pos, full_batch_protein, batch_ligand, sigma, surface_ct):
    total = torch.tensor(0.)
    for g in range(int(batch_ligand.max()) + 1):
        pp, lp = full_protein_pos[full_batch_protein == g], lig_pos[batch_ligand == g]
        e = torch.exp(-torch.sum((pp.view(1, -1, 3) - lp.view(-1, 1, 3)) ** 2, dim=2) / float(sigma))
        G = -sigma * torch.log(1e-3 + e.sum(dim=1))
        total = total + torch.mean(torch.clamp(surface_ct - G, min=0))
    return total


def guidance_grad(xt, offset_l, energy_drift_opt, batch_ligand, decomp_index,
                  full_protein_pos=None, full_batch_protein=None, num_graphs_div=None, score_coef_t=None):
    """Sum of energy gradients w.r.t. x_t (centred frame), shipped drift types only."""
    total = torch.zeros_like(xt)
    for drift in energy_drift_opt:
        x = xt.detach().clone().requires_grad_(True)
        if drift['type'] == 'armsca_prox':
            e, n_valid = armsca_prox_energy(x, batch_ligand, decomp_index, drift['min_d'], drift['max_d'], num_graphs_div)
            if n_valid > 0:
                g = torch.autograd.grad(e, x)[0]
                total = total + (g * score_coef_t if drift.get('scale', False) else g)      # :657-658
        elif drift['type'] == 'clash':
            e = clash_energy(full_protein_pos, x + offset_l, full_batch_protein, batch_ligand,
                             drift['sigma'], drift['gamma'])
            g = torch.autograd.grad(e, x)[0]
            total = total + (g * score_coef_t if drift.get('scale', False) else g)          # :668-669
        else:
            raise ValueError(drift['type'])
    return total


# ----------------------------------------------------------------------------
# reverse loop (models/decompdiff.py:552-703)
# ----------------------------------------------------------------------------
def reverse_step(tab, cfg, preds, ligand_pos, ligand_v, ligand_bond, t, batch_ligand, batch_bond,
                 prior_stds_atom, u_atom, u_bond, eps_pos, grad=None, ligand_atom_mask=None):
    """One posterior step given the network predictions and the three noise draws
    (u_atom (n,8) uniform, u_bond (Eb,5) uniform, eps_pos (n,3) normal) - decompdiff.py:601-685."""
    nc, nb = cfg['num_classes'], cfg['num_bond_classes']
    x0 = preds['pred_ligand_pos']
    if cfg.get('model_mean_type', 'C0') == 'noise':      # decompdiff.py:602-605, _predict_x0_from_eps :353-356
        eps_pred = x0 - ligand_pos
        x0 = tab['sqrt_recip_alphas_cumprod'][t][batch_ligand].unsqueeze(-1) * ligand_pos - \
            tab['sqrt_recipm1_alphas_cumprod'][t][batch_ligand].unsqueeze(-1) * eps_pred
    elif cfg.get('model_mean_type', 'C0') != 'C0':
        raise ValueError
    c0 = tab['posterior_mean_c0_coef'][t][batch_ligand].unsqueeze(-1)
    ct = tab['posterior_mean_ct_coef'][t][batch_ligand].unsqueeze(-1)
    mean = c0 * x0 + ct * ligand_pos
    logvar = tab['posterior_logvar'][t][batch_ligand].unsqueeze(-1)
    nonzero = (1 - (t == 0).float())[batch_ligand].unsqueeze(-1)
    log_v_recon = F.log_softmax(preds['pred_ligand_v'], dim=-1)
    log_v_prob = q_v_posterior(tab, 'atom_type_trans', log_v_recon, index_to_log_onehot(ligand_v, nc), t, batch_ligand)
    v_next = gumbel_argmax(log_v_prob, u_atom)
    log_b_recon = F.log_softmax(preds['pred_bond'], dim=-1)
    log_b_prob = q_v_posterior(tab, 'bond_type_trans', log_b_recon, index_to_log_onehot(ligand_bond, nb), t, batch_bond)
    b_next = gumbel_argmax(log_b_prob, u_bond)
    if grad is not None:
        mean = mean - grad
    pos_next = mean + nonzero * (0.5 * logvar).exp() * eps_pos * prior_stds_atom
    if ligand_atom_mask is not None:
        keep = ligand_atom_mask == 0
        v_next[keep] = ligand_v[keep]
        pos_next[keep] = ligand_pos[keep]
    return dict(pos=pos_next, v=v_next, bond=b_next, log_v_recon=log_v_recon, log_v_prob=log_v_prob,
                log_b_prob=log_b_prob)


@torch.no_grad()
def sample_diffusion(sd, cfg, protein_pos, protein_v, batch_protein, init_ligand_pos, init_ligand_v,
                     ligand_v_aux, batch_ligand, prior_stds, ligand_decomp_batch, ligand_decomp_index,
                     ligand_fc_bond_index, init_ligand_fc_bond_type, batch_ligand_bond,
                     num_steps=None, center_pos_mode='protein', energy_drift_opt=None,
                     full_protein_pos=None, full_batch_protein=None, ligand_atom_mask=None,
                     noise=None, generator=None, keep_traj=True, first_t=None, num_graphs_div=None, **_unused):
    """DecompScorePosNet3D.sample_diffusion.  `noise` = optional list (one per step, first step
    first) of dicts {'u_atom','u_bond','eps_pos'}; otherwise drawn with torch in the reference's
    order rand(n,8) -> rand(Eb,5) -> randn(n,3) from `generator` (or the global CPU generator).
    Test hooks beyond the reference: `first_t` starts the loop at that time index instead of T-1 (teacher-forced single steps),
    `num_graphs_div` is the batch size of the armsca quirk when the inputs are a slice of a larger batch."""
    tab = sd        # the schedule tables are part of the state_dict (models/common.py:280-283)
    T = cfg['num_diffusion_timesteps']
    num_steps = T if num_steps is None else num_steps
    B = int(batch_protein.max()) + 1
    if center_pos_mode == 'protein':
        offset = scatter_mean(protein_pos, batch_protein, dim=0)
        protein_pos = protein_pos - offset[batch_protein]
        ligand_pos = init_ligand_pos - offset[batch_ligand]
        offset_l = offset[batch_ligand]
    elif center_pos_mode in ('none', None):
        ligand_pos, offset_l = init_ligand_pos, torch.zeros_like(init_ligand_pos)
    else:
        raise NotImplementedError(center_pos_mode)
    ligand_v, ligand_bond = init_ligand_v, init_ligand_fc_bond_type
    traj = dict(pos_traj=[], v_traj=[], bond_traj=[], v0_traj=[], vt_traj=[], bt_traj=[])
    n, Eb = ligand_pos.size(0), ligand_bond.numel()
    t_hi = T if first_t is None else first_t + 1
    for s, i in enumerate(reversed(range(t_hi - num_steps, t_hi))):
        t = torch.full((B,), i, dtype=torch.long)
        preds = forward(sd, cfg, protein_pos, protein_v, batch_protein, ligand_pos, ligand_v, ligand_v_aux,
                        batch_ligand, ligand_fc_bond_index, ligand_bond, ligand_atom_mask=ligand_atom_mask, time_step=t)
        if noise is not None:
            u_a, u_b, eps = noise[s]['u_atom'], noise[s]['u_bond'], noise[s]['eps_pos']
        else:
            u_a = torch.rand(n, cfg['num_classes'], generator=generator)
            u_b = torch.rand(Eb, cfg['num_bond_classes'], generator=generator)
            eps = torch.randn(n, 3, generator=generator)
        grad = None
        if energy_drift_opt is not None:
            with torch.enable_grad():
                grad = guidance_grad(ligand_pos, offset_l, energy_drift_opt, batch_ligand, ligand_decomp_index,
                                     full_protein_pos, full_batch_protein, num_graphs_div,
                                     tab['pos_score_coef'][t][batch_ligand].unsqueeze(-1))
        st = reverse_step(tab, cfg, preds, ligand_pos, ligand_v, ligand_bond, t, batch_ligand,
                          batch_ligand_bond, prior_stds[ligand_decomp_batch], u_a, u_b, eps, grad,
                          ligand_atom_mask)
        ligand_pos, ligand_v, ligand_bond = st['pos'], st['v'], st['bond']
        if keep_traj:
            traj['v0_traj'].append(st['log_v_recon']), traj['vt_traj'].append(st['log_v_prob'])
            traj['bt_traj'].append(st['log_b_prob']), traj['bond_traj'].append(ligand_bond.clone())
            traj['pos_traj'].append(ligand_pos + offset_l), traj['v_traj'].append(ligand_v.clone())
    out = dict(pos=ligand_pos + offset_l, v=ligand_v, bond=ligand_bond)
    out.update(traj)
    return out


def config_dict(cfg) -> dict:
    d = dict(cfg)
    d.setdefault('num_classes', 8)
    return d
