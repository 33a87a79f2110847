"""world_size-2 gloo test of the data-parallel plumbing: contiguous batch shards, ONE all-reduce over a
flat gradient buffer, result identical to the single-process gradient.  Runs on CPU: the solve itself is replaced by a
tiny differentiable stand-in because the product has no CPU path; what is tested is the host-side sharding logic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(11, 6, generator=g), torch.randn(11, 1, generator=g)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torchcde_b200.distributed import shard_batch, allreduce_gradients
    model = _model()
    x, y = _data()
    xs, ys = shard_batch([x, y])
    # sum-reduced loss so that shard gradients add up to the full-batch gradient
    loss = ((model(xs) - ys) ** 2).sum()
    loss.backward()
    nbytes = allreduce_gradients(model.parameters())
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in model.parameters()], "nbytes": nbytes,
                    "shard": tuple(xs.shape)}, out)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "online-neural-cdes_b200"))
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    model = _model()
    x, y = _data()
    ((model(x) - y) ** 2).sum().backward()
    for g, p in zip(got["grads"], model.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-6)
    assert got["shard"] == (6, 6)                       # 11 series -> shards of 6 and 5
    assert got["nbytes"] == sum(p.numel() for p in model.parameters()) * 4   # one flat fp32 buffer
