"""Host logic of mcgm_b200.optim that needs no GPU: FlatBuffers turns the parameters and gradients of a
network into views of two flat fp32 buffers without changing names, shapes or values; zero_grad keeps
the views; GradAllReduce shares the gradient buffer; FusedAdam refuses CPU parameters (no CPU path)."""
import pytest
import torch


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Linear(5, 7))


def test_flat_buffers_keep_names_shapes_values_and_alias():
    from mcgm_b200.optim import FlatBuffers, flat_buffers
    net = _net()
    before = {k: v.clone() for k, v in net.state_dict().items()}
    params = list(net.parameters())
    fb = FlatBuffers(params)
    after = net.state_dict()
    assert list(after) == list(before)
    for k in before:
        assert torch.equal(after[k], before[k]) and after[k].shape == before[k].shape
    # every parameter and gradient is a view of the flat buffers, on a 256-byte boundary
    for p, o in zip(fb.params, fb.offsets):
        assert p.data_ptr() == fb.param.data_ptr() + 4 * o and o % 64 == 0
        assert p.grad.data_ptr() == fb.grad.data_ptr() + 4 * o
    # writing the flat buffer writes the parameters; autograd accumulates into the flat gradient
    fb.param.zero_()
    assert all(float(p.abs().sum()) == 0.0 for p in net.parameters())
    net[2].weight.data.fill_(1.0)
    net[2](torch.ones(2, 5)).sum().backward()
    assert float(fb.grad.abs().sum()) > 0
    views = [p.grad for p in fb.params]
    fb.zero_grad()
    assert float(fb.grad.abs().sum()) == 0.0 and all(p.grad is g for p, g in zip(fb.params, views))
    # the same parameters resolve to the same buffers (FusedAdam and GradAllReduce share them)
    assert flat_buffers(params) is fb


def test_grad_allreduce_shares_the_flat_gradient_buffer():
    from mcgm_b200.optim import flat_buffers
    from mcgm_b200.train_gan import GradAllReduce
    net = _net()
    fb = flat_buffers(list(net.parameters()))
    red = GradAllReduce(net, bucket_bytes=64)
    assert red.flat is fb.grad
    covered = sorted((s, e) for s, e, _ in red.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == fb.numel
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))          # buckets tile the buffer


def test_fused_adam_has_no_cpu_path():
    from mcgm_b200.optim import FusedAdam
    with pytest.raises(ValueError, match="Not valid device"):
        FusedAdam(_net().parameters(), lr=1e-3)
    from mcgm_b200 import train_gan
    opt = train_gan.make_optimizer(_net(), 2e-4, (0.0, 0.9))                # CPU parameters -> torch.optim.Adam
    assert isinstance(opt, torch.optim.Adam)


def test_deferred_wgrad_chunking_never_aliases_a_parameter():
    """ops._wgrad_chunks: launches of <= 4 same-shape items; the two passes (real / fake batch) of one layer
    never share a launch, whatever order the backward pass queued them in."""
    from mcgm_b200 import ops
    params = [object() for _ in range(3)]
    items = [(None, None, 3, 3, p, None, None) for p in params] + [(None, None, 3, 3, p, None, None) for p in params]
    chunks = ops._wgrad_chunks(items)
    assert sum(len(c) for c in chunks) == 6 and all(len(c) <= 4 for c in chunks)
    for c in chunks:
        assert len({id(it[4]) for it in c}) == len(c)
    # 8 items of 4 layers (the 8x8 convs of a D update): two full launches
    params = [object() for _ in range(4)]
    chunks = ops._wgrad_chunks([(None, None, 3, 3, p, None, None) for p in params * 2])
    assert [len(c) for c in chunks] == [4, 4]
