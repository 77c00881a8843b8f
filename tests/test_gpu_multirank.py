"""world-size-2 NCCL run of the fused loss (gather_enabled=True) and the GradCache step against the reference goldens.
Needs two GPUs (gpurun --gpus 2); skipped otherwise."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden
from oracle import infonce as O
from oracle.cases import INFONCE_CASES, make_infonce_inputs

pytestmark = pytest.mark.gpu


def _worker(rank, ws, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws)
    from contrastors_b200 import LogitScale, clip_loss
    case = INFONCE_CASES[name]
    qs, ds = make_infonce_inputs(case)
    q = torch.tensor(O.bf16_round(qs[rank]), device="cuda", requires_grad=True)
    d = torch.tensor(O.bf16_round(ds[rank]), device="cuda", requires_grad=True)
    ls = LogitScale(logit_scale=case["scale"], trainable_logit_scale=True).cuda()
    loss = clip_loss(q, d, ls, gather_enabled=True)
    loss.backward()
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=loss.item(), dq=q.grad.cpu().numpy(), dd=d.grad.cpu().numpy(),
             dlogit=ls.logit_scale.grad.item())
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["ws2_square", "ws2_hardneg2"])
def test_clip_loss_two_ranks(name, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = INFONCE_CASES[name]
    mp.spawn(_worker, args=(2, 29621, name, str(tmp_path)), nprocs=2, join=True)
    qs, ds = make_infonce_inputs(case)
    outs = O.clip_loss_multirank([O.bf16_round(x) for x in qs], [O.bf16_round(x) for x in ds], case["scale"])
    z = golden(f"infonce_{name}.npz")
    for r in range(2):
        got = np.load(tmp_path / f"r{r}.npz")
        o = outs[r]
        assert abs(got["loss"] - o["loss"]) <= 1e-3 * max(abs(o["loss"]), 1e-2)
        assert np.abs(got["dq"] - o["dq"]).max() <= 1e-3 * np.abs(o["dq"]).max()
        assert np.abs(got["dd"] - o["dd_local"]).max() <= 1e-3 * np.abs(o["dd_local"]).max()
        assert abs(got["dlogit"] - o["dlogit"]) <= 1e-3 * o["dlogit_abs"] * 2
        assert abs(got["loss"] - float(z[f"r{r}_loss"])) <= 5e-2 * max(float(z[f"r{r}_loss"]), 0.05)
