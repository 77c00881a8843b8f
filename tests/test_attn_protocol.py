"""mbarrier protocols of the attention kernels on the discrete-event model (tools/sim_attn_protocol.py): every kernel
generation terminates without deadlock, barrier over-arrival, parity aliasing or data hazard over random schedules, and the
model does catch a removed wait (negative controls).  CPU only; guards the protocol of the two kernels that have not run on
hardware yet (attn_fwd4_kernel, attn_bwd3_kernel) as well as the default ones."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import sim_attn_protocol as sim  # noqa: E402


@pytest.mark.parametrize("name", sorted(sim.KERNELS))
def test_protocol_terminates_without_hazards(name):
    sim.sweep(name, sizes=range(1, 9), schedules=40)


@pytest.mark.parametrize("name,bug", [("attn_fwd3_kernel", "no_s_free"), ("attn_fwd4_kernel", "no_s_free"),
                                      ("attn_fwd4_kernel", "no_pv_done"), ("attn_fwd5_kernel", "no_q_full"), ("attn_bwd2_kernel", "no_dq_full_wait"),
                                      ("attn_bwd3_kernel", "no_dq_full_wait")])
def test_model_catches_a_removed_wait(name, bug):
    caught = 0
    for n in range(2, 7):
        for seed in range(40):
            try:
                sim.KERNELS[name](n, seed * 31 + n, bug)
            except sim.ProtocolError:
                caught += 1
    assert caught > 0, f"{name}: removing the {bug} wait went unnoticed in 200 schedules"
