"""Unit parity of the stand-alone kernels behind the C ABI: kNN graph (bit-exact vs the oracle) and the projection
GEMM (fp32 FMA kernel and the tcgen05 3xTF32 kernel vs an fp64 reference)."""
import ctypes as C

import pytest
import torch

from decompdiff_b200 import _lib
from oracle import ref_shims

pytestmark = pytest.mark.gpu
P = lambda t: C.c_void_p(t.data_ptr())


@pytest.mark.parametrize('impl', [0, 1], ids=['fp32_fma', 'tcgen05_3xtf32'])
@pytest.mark.parametrize('M,N', [(1, 128), (127, 128), (128, 256), (1000, 640), (25600, 640), (333, 1280)])
def test_gemm128(impl, M, N):
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, 128, generator=g) * 2.0
    Wt = (torch.rand(128, N, generator=g) * 2 - 1) / 128 ** 0.5
    bias = torch.randn(N, generator=g) * 0.1
    want = (A.double() @ Wt.double() + bias.double())
    Ad, Wd, bd = A.cuda(), Wt.cuda(), bias.cuda()
    out = torch.full((M, N), float('nan'), device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().ddb_gemm128(P(Ad), 128, P(Wd), N, P(bd), P(out), N, M, N, 0, impl, st))
    torch.cuda.synchronize()
    err = (out.cpu().double() - want).abs()
    tol = 1e-5 + 1e-4 * want.abs()
    # fp32 FMA sits at fp32 noise; the 3xTF32 split (11-bit hi + 11-bit lo) carries ~1e-6 relative error per product,
    # still well inside the path's tolerance
    ratio = float((err / tol).max())
    print(f'gemm impl={impl} M={M} N={N}: max err {float(err.max()):.2e}, {ratio:.3f} x tol')
    assert ratio < (0.2 if impl == 0 else 0.4), ratio
    assert float(err.max()) < 2e-5


def test_gemm128_softplus_epilogue_matches_between_kernels():
    g = torch.Generator().manual_seed(3)
    A, Wt, bias = torch.randn(300, 128, generator=g).cuda(), (torch.randn(128, 128, generator=g) * 0.1).cuda(), torch.randn(128, generator=g).cuda()
    outs = []
    for impl in (0, 1):
        out = torch.empty(300, 128, device='cuda')
        _lib.check(_lib.lib().ddb_gemm128(P(A), 128, P(Wt), 128, P(bias), P(out), 128, 300, 128, 1, impl, torch.cuda.current_stream().cuda_stream))
        outs.append(out)
    torch.cuda.synchronize()
    want = torch.nn.functional.softplus(A.double() @ Wt.double() + bias.double()) - torch.log(torch.tensor(2.0, dtype=torch.float64))
    for o in outs:
        assert float((o.double() - want).abs().max()) < 2e-5


@pytest.mark.parametrize('sizes,k', [([400] * 8, 32), ([14, 33, 5, 1, 700], 32), ([1100, 40], 32), ([2300], 32), ([50, 60], 7)])
def test_knn_graph_bit_exact(sizes, k):
    g = torch.Generator().manual_seed(sum(sizes))
    n = sum(sizes)
    x = torch.randn(n, 3, generator=g) * 8.0
    x[1] = x[0]                                  # coincident atoms: distance 0, ties broken by index
    if n > 40:
        x[30:36] = torch.round(x[30:36])          # lattice points: exact distance ties
    is_lig = (torch.rand(n, generator=g) < 0.1)
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32)
    x4 = torch.cat([x, torch.zeros(n, 1)], 1).cuda().contiguous()
    nbr = torch.full((n, 32), -1, dtype=torch.int32, device='cuda')
    deg = torch.empty(n, dtype=torch.int32, device='cuda')
    nlig = torch.empty(n, dtype=torch.int32, device='cuda')
    ptr_d, lig_d = ptr.cuda(), is_lig.to(torch.uint8).cuda()      # keep the device tensors alive across the call
    _lib.check(_lib.lib().ddb_knn_graph(P(x4), P(ptr_d), P(lig_d), len(sizes), n, k, P(nbr), P(deg), P(nlig),
                                        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    src, dst = ref_shims.knn_graph(x, k=k, batch=batch)      # oracle: nearest first, ties -> lower index
    nbr, deg, nlig = nbr.cpu().long(), deg.cpu().long(), nlig.cpu().long()
    want_deg = torch.bincount(dst, minlength=n)
    assert torch.equal(deg, want_deg)
    start = 0
    for i in range(n):
        d = int(deg[i])
        want = src[start:start + d]
        start += d
        got = nbr[i, :d]
        lig_first = torch.cat([want[is_lig[want]], want[~is_lig[want]]])      # stable partition, ligand sources first
        assert torch.equal(got, lig_first), (i, got.tolist(), lig_first.tolist())
        assert int(nlig[i]) == int(is_lig[want].sum())


def test_cached_graph_build_equals_brute_force(model_cpu, monkeypatch):
    """The static protein-neighbour cache + rank-merge (graph.cu: knn_merge_kernel) and the two-launch level lists must reproduce
    the brute-force kNN / multi-launch lists bit for bit: same neighbour lists, same trajectory."""
    from decompdiff_b200 import synthetic as syn
    kw = syn.make_batch(n_pockets=5, n_protein=[370, 40, 25, 600, 90], arm_sizes=(8, 8), n_scaffold=14, seed=71)
    n, Eb = kw['init_ligand_pos'].size(0), kw['init_ligand_fc_bond_type'].numel()
    noise = syn.step_noise(n, Eb, 4, seed=3)

    def run():
        r = model_cpu.begin_sampling(**kw, num_steps=4, center_pos_mode='protein')
        r.advance(4, noise=noise)
        out = r.finish(traj_on_device=True)
        return out, r.eb.debug_buffer('nbr').cpu(), r.eb.debug_buffer('deg').cpu(), r.eb.debug_buffer('nlig').cpu()

    fast = run()
    monkeypatch.setenv('DDB_NO_KNN_CACHE', '1')
    monkeypatch.setenv('DDB_OLD_LISTS', '1')
    slow = run()
    assert torch.equal(fast[2], slow[2]) and torch.equal(fast[3], slow[3])
    deg = fast[2].view(-1)
    mask = torch.arange(32)[None, :] < deg[:, None]
    assert torch.equal(fast[1][mask], slow[1][mask])
    for k in ('pos', 'v', 'bond'):
        assert torch.equal(fast[0][k], slow[0][k]), k
    assert torch.equal(fast[0]['pos_traj'], slow[0]['pos_traj'])
