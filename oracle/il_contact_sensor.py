"""TEST INFRASTRUCTURE (oracle): [IL] ContactSensor air / contact-time bookkeeping and force history, restated.

IsaacLab 2.2.1 `isaaclab/sensors/contact_sensor/contact_sensor.py` (`_update_buffers_impl`, `reset`) is NOT under /root/reference and
cannot be installed here; the reference configures it in locotouch/config/locotouch/locomotion_base_env_cfg.py:35-39,358-359
(history_length 3, track_air_time True, force_threshold default 1.0) and reads its outputs in locotouch/mdp/rewards.py:116-156,
596-604 and observations.py:60-66.  Restated from SURVEY.md App. B: parity for this file is UNPINNED beyond the invariants checked
in tests/test_contact_sensor.py (mutually exclusive accumulators, latching on transition, reset).  Only tests may import this."""
import torch


def contact_sensor_update(state: dict, forces: torch.Tensor, dt, force_threshold: float = 1.0, reset_env_ids=None) -> dict:
    """state: net_forces_w [N,B,3], net_forces_w_history [N,H,B,3], current_air_time / last_air_time / current_contact_time /
    last_contact_time [N,B].  Returns the new state (inputs untouched).  `dt` scalar or [N]."""
    s = {k: v.clone() for k, v in state.items()}
    s["net_forces_w"] = forces.clone()
    h = s["net_forces_w_history"]
    h[:, 1:] = state["net_forces_w_history"][:, :-1].clone()
    h[:, 0] = forces
    elapsed = torch.as_tensor(dt, dtype=torch.float32).expand(forces.shape[0]) if not torch.is_tensor(dt) or dt.dim() == 0 else dt
    is_contact = torch.norm(forces, dim=-1) > force_threshold
    is_first_contact = (state["current_air_time"] > 0) * is_contact
    is_first_detached = (state["current_contact_time"] > 0) * ~is_contact
    e = elapsed.unsqueeze(-1)
    s["last_air_time"] = torch.where(is_first_contact, state["current_air_time"] + e, state["last_air_time"])
    s["current_air_time"] = torch.where(~is_contact, state["current_air_time"] + e, torch.zeros(()))
    s["last_contact_time"] = torch.where(is_first_detached, state["current_contact_time"] + e, state["last_contact_time"])
    s["current_contact_time"] = torch.where(is_contact, state["current_contact_time"] + e, torch.zeros(()))
    if reset_env_ids is not None:  # ContactSensor.reset(env_ids)
        for k in s:
            s[k][reset_env_ids] = 0.0
    return s
