"""mbarrier protocols of the attention kernels on the discrete-event model (tools/sim_attn_protocol.py): every kernel
terminates without deadlock, barrier over-arrival, parity aliasing or data hazard over random schedules, and the model does catch
a removed wait (negative controls).  CPU only; the two kernels modelled are the ones in the library (attn_fwd4_kernel,
attn_bwd4_kernel; the backward model was updated with the kernel when its schedule changed in round 2)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import sim_attn_protocol as sim  # noqa: E402


@pytest.mark.parametrize("name", sorted(sim.KERNELS))
def test_protocol_terminates_without_hazards(name):
    sim.sweep(name, sizes=range(1, 9), schedules=40)


@pytest.mark.parametrize("name,bug", [("attn_fwd4_kernel", "no_s_free"), ("attn_fwd4_kernel", "no_pv_done"),
                                      ("attn_bwd4_kernel", "no_dq_full_wait"), ("attn_bwd4_kernel", "early_lse_write")])
def test_model_catches_a_removed_wait(name, bug):
    caught = 0
    for n in range(2, 7):
        for seed in range(40):
            try:
                sim.KERNELS[name](n, seed * 31 + n, bug)
            except sim.ProtocolError:
                caught += 1
    assert caught > 0, f"{name}: removing the {bug} wait went unnoticed in 200 schedules"
