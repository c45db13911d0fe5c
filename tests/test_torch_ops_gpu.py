"""GPU: the dispatcher registrations `torch.ops.escgnn.*` (esc_gnn_b200/torch_ops.py) give the same values and gradients as the
autograd Functions the drop-in modules use, and as a plain-torch restatement."""
import pytest
import torch

from oracle import model_ref
from tests.test_model_gpu import product_batch

pytestmark = pytest.mark.gpu


def test_registered_ops_match_module_path_and_torch():
    from esc_gnn_b200 import ops, torch_ops
    b = product_batch(2, 900, 24)
    idx = ops.graph_index(b)
    E, N = b.edge_index.size(1), b.x.size(0)
    g = torch.Generator(device='cuda').manual_seed(5)
    H = 64
    # ---- bag_embed
    W = torch.randn(1800, H, device='cuda', generator=g, requires_grad=True)
    W2 = W.detach().clone().requires_grad_(True)
    out = torch.ops.escgnn.bag_embed(W, b.pos_index, b.pos_enc, idx.rec_ptr, E)
    ref = model_ref.bag_embed(W2, b.pos_index, b.pos_enc, b.pos_batch)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-4)
    go = torch.randn(E, H, device='cuda', generator=g)
    out.backward(go); ref.backward(go)
    torch.testing.assert_close(W.grad, W2.grad, rtol=1e-4, atol=1e-3)
    # ---- gine_aggregate
    x = torch.randn(N, H, device='cuda', generator=g, requires_grad=True)
    ef = torch.randn(E, H, device='cuda', generator=g, requires_grad=True)
    eps = torch.tensor([0.3], device='cuda', requires_grad=True)
    x2, ef2, eps2 = (t.detach().clone().requires_grad_(True) for t in (x, ef, eps))
    y = torch.ops.escgnn.gine_aggregate(x, ef, eps, *torch_ops.index_tensors(idx))
    y2 = ops.gine_aggregate(x2, ef2, eps2, idx)
    assert torch.equal(y, y2)
    want = (1 + eps.detach()) * x.detach() + torch.zeros_like(x).index_add_(0, b.edge_index[1], torch.relu(x.detach()[b.edge_index[0]] + ef.detach()))
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)
    gy = torch.randn(N, H, device='cuda', generator=g)
    y.backward(gy); y2.backward(gy)
    for a, c in ((x, x2), (ef, ef2), (eps, eps2)):
        assert torch.equal(a.grad, c.grad)
    # ---- segment_pool
    for mean in (False, True):
        xp = torch.randn(N, H, device='cuda', generator=g, requires_grad=True)
        xq = xp.detach().clone().requires_grad_(True)
        p = torch.ops.escgnn.segment_pool(xp, idx.graph_ptr, idx.num_graphs, mean)
        q = (ops.global_mean_pool if mean else ops.global_add_pool)(xq, idx)
        assert torch.equal(p, q)
        gp = torch.randn_like(p)
        p.backward(gp); q.backward(gp)
        assert torch.equal(xp.grad, xq.grad)
    # ---- gemm / linear
    a = torch.randn(700, 96, device='cuda', generator=g)
    w = torch.randn(160, 96, device='cuda', generator=g)
    bias = torch.randn(160, device='cuda', generator=g)
    c = torch.ops.escgnn.gemm(a, False, w, False, bias)
    torch.testing.assert_close(c.double(), a.double() @ w.double().t() + bias.double(), rtol=1e-5, atol=1e-4)
    ar = a.clone().requires_grad_(True)
    yl = torch_ops.linear(ar, w, bias)
    yl.sum().backward()
    torch.testing.assert_close(ar.grad.double(), torch.ones(700, 160, device='cuda', dtype=torch.float64) @ w.double(), rtol=1e-5, atol=1e-4)
    # ---- edge_distance
    pos = torch.randn(N, 3, device='cuda', generator=g)
    d = torch.ops.escgnn.edge_distance(pos, b.edge_index, False)
    torch.testing.assert_close(d.view(-1), (pos[b.edge_index[1]] - pos[b.edge_index[0]]).norm(dim=1), rtol=1e-5, atol=1e-6)
    # ---- the dispatcher knows the schemas and refuses CPU tensors (no CPU implementation is registered)
    assert 'escgnn::gine_aggregate' in str(torch.ops.escgnn.gine_aggregate.default._schema)
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.escgnn.segment_pool(torch.zeros(4, 8), torch.tensor([0, 4], dtype=torch.int32), 1, False)
