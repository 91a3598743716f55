"""K18: ContactSensor bookkeeping (history ring + air / contact-time state machine) against the restated IsaacLab rule
(oracle/il_contact_sensor.py, [IL] unpinned by necessity) -- bit-exact over a chain of steps with resets -- and the rule's own
invariants on the CPU."""
import pytest
import torch

from oracle import il_contact_sensor as CS
from tests import helpers as H


def make_state(n, b, h, gen):
    in_contact = torch.rand(n, b, generator=gen) < 0.5
    t = torch.rand(n, b, generator=gen) * 0.6
    return dict(net_forces_w=torch.zeros(n, b, 3), net_forces_w_history=torch.randn(n, h, b, 3, generator=gen),
                current_contact_time=torch.where(in_contact, t, torch.zeros(())), current_air_time=torch.where(in_contact, torch.zeros(()), t),
                last_air_time=torch.rand(n, b, generator=gen), last_contact_time=torch.rand(n, b, generator=gen))


def draw_forces(n, b, gen):
    f = torch.randn(n, b, 3, generator=gen) * 2.0
    f[torch.rand(n, b, generator=gen) < 0.5] = 0.0
    near = torch.rand(n, b, generator=gen) < 0.1   # around the 1 N threshold
    f[near] = torch.nn.functional.normalize(torch.randn(int(near.sum()), 3, generator=gen), dim=-1) * (1.0 + 1e-3 * torch.randn(int(near.sum()), 1, generator=gen))
    return f


def test_oracle_invariants():
    gen = torch.Generator().manual_seed(0)
    s = make_state(64, 5, 3, gen)
    for _ in range(30):
        prev = s
        f = draw_forces(64, 5, gen)
        s = CS.contact_sensor_update(prev, f, 0.02)
        contact = torch.norm(f, dim=-1) > 1.0
        assert ((s["current_air_time"] > 0) & (s["current_contact_time"] > 0)).sum() == 0, "accumulators are mutually exclusive"
        assert (s["current_contact_time"][~contact] == 0).all() and (s["current_air_time"][contact] == 0).all()
        landed = contact & (prev["current_air_time"] > 0)
        H.assert_equal(s["last_air_time"][landed], (prev["current_air_time"] + 0.02)[landed], "last_air_time latched at touch-down")
        H.assert_equal(s["last_air_time"][~landed], prev["last_air_time"][~landed], "last_air_time otherwise kept")
        H.assert_equal(s["net_forces_w_history"][:, 1:], prev["net_forces_w_history"][:, :-1], "history ring shifted")
        H.assert_equal(s["net_forces_w_history"][:, 0], f, "newest force in slot 0")
    s = CS.contact_sensor_update(s, draw_forces(64, 5, gen), 0.02, reset_env_ids=torch.tensor([1, 7]))
    assert all((v[[1, 7]] == 0).all() for v in s.values())


@pytest.mark.gpu
@pytest.mark.parametrize("n,b,h", [(1, 1, 1), (405, 17, 3), (4097, 13, 3), (64, 221, 0)])
def test_contact_sensor_kernel_matches_the_restated_rule(cuda, lt_lib, n, b, h):
    from locotouch_b200.sim.contact_sensor import ContactSensorState

    gen = torch.Generator().manual_seed(n + b)
    s = make_state(n, b, max(h, 1), gen)
    sensor = ContactSensorState(n, [f"body_{i}" for i in range(b)], history_length=h, device=cuda)
    d = sensor.data
    for k, v in s.items():
        getattr(d, k).copy_(v)
    for step in range(12):
        f = draw_forces(n, b, gen)
        reset = (torch.rand(n, generator=gen) < 0.05) if step % 3 == 2 else None
        s = CS.contact_sensor_update(s, f, 0.02, reset_env_ids=reset.nonzero().flatten() if reset is not None else None)
        sensor.update(f.to(cuda), 0.02, reset_mask=reset.to(torch.uint8).to(cuda) if reset is not None else None)
        for k in ("current_air_time", "last_air_time", "current_contact_time", "last_contact_time", "net_forces_w"):
            H.assert_equal(getattr(d, k), s[k], f"{k} after step {step}")
        if h > 0:
            H.assert_equal(d.net_forces_w_history, s["net_forces_w_history"], f"history after step {step}")
    assert sensor.compute_first_contact(0.02).shape == (n, b)
