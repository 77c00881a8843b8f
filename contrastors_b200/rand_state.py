"""RNG snapshot/replay for the GradCache second pass (reference: /root/reference/src/contrastors/rand_state.py:6-21).

Same contract as the reference's ``RandContext``: constructing it records the CPU generator state and the generator
state of every CUDA device the given tensors live on; entering it forks the RNG and restores those states so the
re-forward draws the same dropout masks; leaving it restores whatever was current.
"""
from __future__ import annotations

import torch


class RandContext:
    def __init__(self, tensors):
        if isinstance(tensors, dict):
            tensors = list(tensors.values())
        elif isinstance(tensors, torch.Tensor):
            tensors = [tensors]
        self.cpu_state = torch.get_rng_state()
        devs = sorted({t.get_device() for t in tensors if isinstance(t, torch.Tensor) and t.is_cuda})
        self.devices = devs
        self.cuda_states = [torch.cuda.get_rng_state(d) for d in devs]
        self._fork = None

    def __enter__(self):
        self._fork = torch.random.fork_rng(devices=self.devices, enabled=True)
        self._fork.__enter__()
        torch.set_rng_state(self.cpu_state)
        for d, s in zip(self.devices, self.cuda_states):
            torch.cuda.set_rng_state(s, d)
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        fork, self._fork = self._fork, None
        return fork.__exit__(exc_type, exc_val, exc_tb)
