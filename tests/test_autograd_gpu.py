"""GPU: SURVEY.md section 8(f) row 4 -- the accelerated volume constructors and the soft-argmin inside an autograd graph.
Expected gradients = PyTorch autograd through the CPU oracle (the reference's own slice-assignment loops are differentiable aten
code).  Bars: <= 1e-5 of the gradient scale (fp32 sums of <= D*K products in a different order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cost_volume as ocv      # noqa: E402
from oracle import regression as oreg      # noqa: E402


@pytest.fixture(scope="module")
def ops():
    import __graft_entry__
    __graft_entry__.build()
    from openstereo_b200 import ops
    return ops


def rnd(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def rel(got, want):
    return ((got.detach().cpu() - want).abs().max() / want.abs().max().clamp(min=1e-12)).item()


def grads(fn, inputs, weight):
    leaves = [t.clone().requires_grad_(True) for t in inputs]
    out = fn(*leaves)
    (out * weight.to(out.device)).sum().backward()
    return out.detach(), [t.grad for t in leaves]


@pytest.mark.parametrize("shape,d,g", [((2, 24, 5, 37), 9, 4), ((1, 40, 3, 128), 48, 8), ((1, 6, 2, 5), 8, 3)])
def test_gwc_volume_backward(ops, shape, d, g):
    l, r = rnd(1, *shape), rnd(2, *shape)
    wgt = rnd(3, shape[0], g, d, shape[2], shape[3])
    want_out, want = grads(lambda a, b: ocv.build_gwc_volume(a, b, d, g), (l, r), wgt)
    got_out, got = grads(lambda a, b: ops.build_gwc_volume(a, b, d, g), (l.cuda(), r.cuda()), wgt)
    assert rel(got_out, want_out) <= 1e-5
    assert rel(got[0], want[0]) <= 1e-5 and rel(got[1], want[1]) <= 1e-5


def test_gwc_volume_backward_one_sided(ops):
    """Only the left features require grad (a frozen right branch): the right gradient is neither computed nor returned."""
    l, r = rnd(4, 1, 16, 4, 32).cuda().requires_grad_(True), rnd(5, 1, 16, 4, 32).cuda()
    ops.build_gwc_volume(l, r, 6, 4).sum().backward()
    assert l.grad is not None and r.grad is None and l.grad.abs().sum() > 0


@pytest.mark.parametrize("mask_left", [True, False])
def test_concat_volume_backward(ops, mask_left):
    shape, d = (2, 5, 4, 21), 7
    l, r = rnd(6, *shape), rnd(7, *shape)
    wgt = rnd(8, shape[0], 2 * shape[1], d, shape[2], shape[3])
    _, want = grads(lambda a, b: ocv.build_concat_volume(a, b, d, mask_left=mask_left), (l, r), wgt)
    _, got = grads(lambda a, b: ops.build_concat_volume(a, b, d, mask_left=mask_left), (l.cuda(), r.cuda()), wgt)
    assert rel(got[0], want[0]) <= 1e-6 and rel(got[1], want[1]) <= 1e-6


def test_correlation_and_coex_backward(ops):
    l, r = rnd(9, 2, 12, 3, 30), rnd(10, 2, 12, 3, 30)
    wgt = rnd(11, 2, 8, 3, 30)
    _, want = grads(lambda a, b: ocv.correlation_volume(a, b, 8), (l, r), wgt)
    _, got = grads(lambda a, b: ops.correlation_volume(a, b, 8), (l.cuda(), r.cuda()), wgt)
    assert rel(got[0], want[0]) <= 1e-5 and rel(got[1], want[1]) <= 1e-5
    wgt = rnd(12, 2, 3, 7, 3, 30)
    _, want = grads(lambda a, b: ocv.coex_cost_volume(a, b, 6, 3), (l, r), wgt)
    _, got = grads(lambda a, b: ops.coex_cost_volume(a, b, 6, 3), (l.cuda(), r.cuda()), wgt)
    assert rel(got[0], want[0]) <= 1e-5 and rel(got[1], want[1]) <= 1e-5


@pytest.mark.parametrize("normalize", [True, False])
def test_softargmin_backward(ops, normalize):
    cost = rnd(13, 2, 48, 6, 20, scale=3.0)
    wgt = rnd(14, 2, 6, 20)
    if normalize:
        ref_fn = lambda c: oreg.disparity_regression(torch.softmax(c, 1), 48, keepdim=False)
        my_fn = lambda c: ops.softargmin(c, 48, keepdim=False)
    else:
        prob = torch.softmax(cost, 1)
        cost = prob
        ref_fn = lambda c: oreg.disparity_regression(c, 48, keepdim=False)
        my_fn = lambda c: ops.disparity_regression(c, 48, keepdim=False)
    want_out, want = grads(ref_fn, (cost,), wgt)
    got_out, got = grads(my_fn, (cost.cuda(),), wgt)
    assert rel(got_out, want_out) <= 1e-5 and rel(got[0], want[0]) <= 2e-5
