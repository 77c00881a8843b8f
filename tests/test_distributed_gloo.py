"""N > 1 host logic on CPU (gloo, world size 2): the autograd all-gather, the reduce-scatter fallback, the flat-buffer
gradient all-reduce.  The kernels themselves need a GPU (tests/test_gpu_multirank.py)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, ws, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import torch.distributed.nn
    from contrastors_b200.distributed import all_gather_rows, gather, gather_with_grad, reduce_scatter_rows
    from contrastors_b200.parallel import allreduce_gradients
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 7, requires_grad=True)
    w = torch.randn(ws * 5, 7, generator=torch.Generator().manual_seed(7)) * (rank + 1)
    # ours
    g = gather_with_grad(x)
    (g * w).sum().backward()
    mine_fwd, mine_bwd = g.detach().clone(), x.grad.clone()
    # the reference's implementation (distributed.py:5-12)
    x2 = x.detach().clone().requires_grad_()
    g2 = torch.cat(torch.distributed.nn.all_gather(x2), dim=0)
    (g2 * w).sum().backward()
    assert torch.equal(mine_fwd, g2.detach())
    assert torch.allclose(mine_bwd, x2.grad, rtol=1e-6, atol=1e-6)
    # plain helpers
    full = all_gather_rows(x.detach())
    assert torch.equal(full, mine_fwd)
    rs = reduce_scatter_rows(w.clone())
    tot = w.clone()
    dist.all_reduce(tot)
    assert torch.allclose(rs, tot[rank * 5:(rank + 1) * 5])
    gg = gather(x.detach())
    assert torch.equal(gg, mine_fwd)
    # 0-dim tensors are unsqueezed (distributed.py:9-10)
    s = gather_with_grad(torch.tensor(float(rank)))
    assert s.shape == (ws,) and s.tolist() == [float(r) for r in range(ws)]

    class FakeTrunk:
        def __init__(self):
            self._g = torch.full((11,), float(rank + 1))

        def flat_grad(self):
            return self._g

    t = FakeTrunk()
    allreduce_gradients(t, t)  # the same tower twice (tower1 is tower2) must be reduced once
    assert torch.allclose(t._g, torch.full((11,), sum(range(1, ws + 1)) / ws))
    # bucketed reduction (parallel.GradientBucketReducer): layer buckets in backward order + the remainder = one all-reduce
    from contrastors_b200.parallel import GradientBucketReducer

    class FakeLayered(FakeTrunk):
        _n_total = 64

        def __init__(self):
            self._g = torch.arange(64, dtype=torch.float32) * (rank + 1)

        def layer_grad_slices(self):
            return [(8, 24), (24, 40)]  # two "layers"; [0, 8) = embeddings, [40, 64) = 1-D parameters

    lt = FakeLayered()
    red = GradientBucketReducer(lt)
    red.arm()
    assert lt._bucket_reducer is red
    red.layer_done(1)
    red.layer_done(0)
    red.finish()
    assert lt._bucket_reducer is None and red.wait() == 1.0 / ws
    assert torch.equal(lt._g, torch.arange(64, dtype=torch.float32) * sum(range(1, ws + 1)))
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


def test_gloo_world_size_2(tmp_path):
    mp.spawn(_worker, args=(2, 29611, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(2))


def test_label_layout_matches_reference_rule():
    # loss.py:108-117 through the oracle helper the kernels are tested against
    from oracle.infonce import labels_for
    assert labels_for(4, 16, 1, 2).tolist() == [8, 10, 12, 14]  # (arange(4) + 1*4) * (16 // (4*2))
    assert labels_for(3, 3, 0, 1).tolist() == [0, 1, 2]
