"""K1 parity: the fused MDP step (terminations, rewards incl. the stateful gait term, observations) vs the oracle on the
same seeded inputs over 36 chained steps with resets, and vs the reference golden vectors."""
import pytest
import torch

from tests import helpers as H
from tests import scenarios as S
from locotouch_b200.sim import synth

pytestmark = pytest.mark.gpu

MASK_TERMS = ("foot_dragging", "thigh_calf_collision", "object_dangerous_state", "object_z_contact")


def _device_env(env, cuda):
    return env.to(cuda)


@pytest.mark.parametrize("scenario", ["locomotion", "teacher"])
@pytest.mark.parametrize("fused", [True, False])
def test_mdp_step_matches_oracle_and_reference(cuda, lt_lib, scenario, fused):
    from locotouch_b200.mdp.fused import FusedMdp

    gold = H.load_golden(f"mdp_{scenario}.npz")
    gnames = [str(x) for x in gold["term_names"]]
    state = {}

    def on_step(step, env, out, pol, cri, u_obs, u_euler, oracle):
        denv = _device_env(env, cuda)
        if "mdp" not in state:
            state["mdp"] = FusedMdp(denv, oracle.spec)
        mdp = state["mdp"]
        mdp.env = denv
        mdp._bound_ptrs = None
        kw = dict(u_obs=u_obs.to(cuda), u_obj_euler=u_euler.to(cuda))
        if fused:
            mdp.step(True, True, **kw)
        else:  # IsaacLab order: terminations + rewards, (reset), observations
            mdp.compute_rewards()
            mdp.compute_observations(**kw)
        torch.cuda.synchronize()
        spec = oracle.spec
        # ---- masks: bit-exact
        H.assert_equal(mdp.terminated, out["terminated"], f"step {step} terminated")
        H.assert_equal(mdp.time_outs, out["time_outs"], f"step {step} time_outs")
        H.assert_equal(mdp.dones, out["done"], f"step {step} dones")
        H.assert_equal(mdp.dones, gold["done"][step], f"step {step} dones vs golden")
        for t in spec.terminations:
            H.assert_equal(mdp.termination(t.name), out["masks"][t.name], f"step {step} termination {t.name}")
        # ---- every reward term column vs oracle and vs the reference golden
        for i, t in enumerate(spec.rewards):
            if t.weight == 0.0:
                continue
            val = mdp.term(t.name)
            ref = out["raw"][t.name].float()
            if t.name in MASK_TERMS:
                H.assert_equal(val, ref, f"step {step} {t.name}")
            else:
                H.assert_close(val, ref, f"step {step} {t.name}")
            if t.name in gnames:
                g = gold["raw"][step, gnames.index(t.name)]
                if t.name in MASK_TERMS:
                    H.assert_equal(val, g, f"step {step} {t.name} vs golden")
                else:
                    H.assert_close(val, g, f"step {step} {t.name} vs golden")
        H.assert_close(mdp.reward_buf, out["reward"], f"step {step} reward")
        H.assert_close(mdp.step_reward, oracle.step_reward, f"step {step} step_reward", rtol=1e-5, atol=1e-5)
        H.assert_close(mdp.episode_sums, oracle.episode_sums, f"step {step} episode sums", rtol=1e-5, atol=1e-5)
        # ---- gait state carried across steps
        g = oracle.gait
        H.assert_close(mdp.valid_last_air_time, g.vla, f"step {step} valid_last_air_time")
        H.assert_equal(mdp.swinging_in_zero_cmd, g.sz, f"step {step} swinging_in_zero_cmd")
        H.assert_equal(mdp.valid_previous_contact, g.vpc, f"step {step} valid_previous_contact")
        H.assert_close(mdp.step_from_changing_cmd, g.steps, f"step {step} step_from_changing_cmd")
        # ---- observations (history of 6, policy noisy / critic clean)
        H.assert_close(mdp.policy_obs, pol, f"step {step} policy obs")
        H.assert_close(mdp.critic_obs, cri, f"step {step} critic obs")

    oracle, _ = S.run_oracle_mdp(scenario, on_step)
    mdp = state["mdp"]
    H.assert_close(mdp.valid_last_air_time, gold["gait_valid_last_air_time"], "final valid_last_air_time vs golden")
    H.assert_equal(mdp.swinging_in_zero_cmd, gold["gait_swinging_in_zero_cmd"], "final swinging_in_zero_cmd vs golden")
    H.assert_close(mdp.last_velocity_cmd, gold["gait_last_velocity_cmd"], "final last_velocity_cmd vs golden")
    # episode log: every reset contributed its sums
    assert float(mdp.episode_log_sums[-1]) == float(gold["done"].sum())


def test_any_nonzero_cmd_false_branch_and_reset(cuda, lt_lib):
    """All commands zero: the cross-env torch.any() at reference rewards.py:190 is False -> no landing update."""
    from locotouch_b200.mdp import task_spec as TS
    from locotouch_b200.mdp.fused import FusedMdp
    from oracle.mdp import MdpOracle

    env = synth.make_env(64, seed=5)
    spec = TS.locomotion_spec()
    oracle = MdpOracle(env, spec)
    mdp = None
    for step in range(6):
        if step >= 3:
            env.command_manager.get_command("base_velocity").zero_()
        out = oracle.step(env)
        denv = env.to(cuda)
        if mdp is None:
            mdp = FusedMdp(denv, spec)
        mdp.env, mdp._bound_ptrs = denv, None
        mdp.compute_rewards()
        H.assert_close(mdp.term("gait"), out["raw"]["gait"], f"gait step {step}")
        H.assert_close(mdp.valid_last_air_time, oracle.gait.vla, f"vla step {step}")
        synth.advance(env, keep_cmd_prob=1.0)
    mdp.reset(torch.tensor([1, 5], device=cuda))
    oracle.gait.reset(torch.tensor([1, 5]))
    H.assert_close(mdp.last_step_current_air_time, oracle.gait.lsa, "reset(env_ids)")
    assert float(mdp.episode_sums[:, [1, 5]].abs().sum()) == 0.0


@pytest.mark.parametrize("n", [1, 3, 4097])
def test_ragged_env_counts_and_rng_mode(cuda, lt_lib, n):
    from locotouch_b200.mdp import task_spec as TS
    from locotouch_b200.mdp.fused import FusedMdp
    from oracle.mdp import MdpOracle

    env = synth.make_env(n, seed=100 + n, with_object=True)
    spec = TS.teacher_spec()
    oracle = MdpOracle(env, spec)
    out = oracle.step(env)
    mdp = FusedMdp(env.to(cuda), spec, seed=1)
    mdp.step(True, True)  # in-kernel Philox noise
    H.assert_close(mdp.reward_buf, out["reward"], "reward")
    _, cri = oracle.observe(env, u_noise=torch.zeros(n, spec.obs_dim_per_step), u_obj_euler=torch.zeros(n, 3))
    H.assert_close(mdp.critic_obs, cri, "critic obs (noise-free group)")
    # policy noise stays inside the configured bands: |policy - critic| <= max(|n_min|,|n_max|) * scale
    diff = (mdp.policy_obs - mdp.critic_obs).abs().cpu()
    col = 0
    for t in spec.obs_terms:
        w = t.dim * spec.history_length
        if t.name != "object_state":
            bound = (max(abs(t.noise[0]), abs(t.noise[1])) * t.scale + 1e-6) if t.noise else 0.0
            assert float(diff[:, col:col + w].max()) <= bound, t.name
        col += w
    if n > 1000:
        assert float(diff[:, 18:36].mean()) > 0.01  # noise is actually there (projected gravity columns)


def _last_action_columns(spec):
    col = 0
    for t in spec.obs_terms:
        w = t.dim * spec.history_length
        if t.name == "last_action":
            return col, w, t
        col += w
    raise AssertionError("no last_action term")


@pytest.mark.parametrize("scenario", ["locomotion", "teacher"])
def test_action_term_reset_inside_the_step(cuda, lt_lib, scenario):
    """``reset_action_term``: IsaacLab's step order (rewards -> ActionManager.reset(env_ids) -> observations; reference
    mdp/actions.py:46-52) inside ONE launch.  36 chained steps vs the oracle with the same order: action-term state bit-exact, both
    observation groups (explicit uniforms), and the post-reset rows carry a zero last_action in every history slot of the critic."""
    from locotouch_b200.mdp.fused import FusedMdp
    from oracle.mdp import MdpOracle

    spec, env, steps = H.make_mdp_env(scenario)
    oracle = MdpOracle(env, spec)
    mdp, seen = None, 0
    col, width, _ = _last_action_columns(spec)
    for step in range(steps):
        denv = env.to(cuda)  # pre-reset copy: the launch resets its own action term
        before = env.action_manager.get_term("joint_pos").raw_actions.clone()
        out = oracle.step(env, auto_reset=True, reset_action_term=True)
        u_obs, u_euler = H.mdp_noise(scenario, step, env.num_envs, spec.obs_dim_per_step)
        pol, cri = oracle.observe(env, u_noise=u_obs, u_obj_euler=u_euler)
        if mdp is None:
            mdp = FusedMdp(denv, spec)
        mdp.env, mdp._bound_ptrs = denv, None
        dterm = denv.action_manager.get_term("joint_pos")
        mdp.step(True, True, u_obs=u_obs.to(cuda), u_obj_euler=u_euler.to(cuda), reset_action_term=True,
                 action_term=dict(prev_prev_raw=dterm.prev_prev_raw_actions))
        torch.cuda.synchronize()
        H.assert_equal(mdp.dones, out["done"], f"step {step} dones")
        oterm = env.action_manager.get_term("joint_pos")
        for k in ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions"):
            H.assert_equal(getattr(dterm, k), getattr(oterm, k), f"step {step} {k}")
        H.assert_close(mdp.policy_obs, pol, f"step {step} policy obs")
        H.assert_close(mdp.critic_obs, cri, f"step {step} critic obs")
        ids = out["done"].nonzero().flatten()
        if len(ids):
            seen += int((before[ids].abs().sum(dim=1) > 0).sum())
            assert float(mdp.critic_obs[ids.to(cuda), col:col + width].abs().max()) == 0.0
            assert float(dterm.raw_actions[ids.to(cuda)].abs().max()) == 0.0
        H.advance_mdp_env(env, step)
    assert seen > 0, "no env with a non-zero action was reset: the scenario does not exercise the path"


def test_action_term_reset_fused_equals_split_phases(cuda, lt_lib):
    """Production mode (in-kernel Philox noise, ragged env count): the in-launch reset == compute_rewards -> term.reset(env_ids) ->
    compute_observations, bit for bit (the re-drawn last_action noise uses the same Philox words); and a rewards-only launch with the
    flag zeroes the action term for the observation pass that follows."""
    from locotouch_b200.mdp import task_spec as TS
    from locotouch_b200.mdp.fused import FusedMdp

    spec = TS.teacher_spec()
    env = synth.make_env(4097, seed=321, with_object=True)
    envs = [env.to(cuda) for _ in range(3)]
    mdps = [FusedMdp(e, spec, seed=9) for e in envs]
    terms = [e.action_manager.get_term("joint_pos") for e in envs]
    kw = [dict(reset_action_term=True, action_term=dict(prev_prev_raw=t.prev_prev_raw_actions)) for t in terms]
    mdps[0].step(True, True, **kw[0])
    mdps[1].compute_rewards()
    ids = mdps[1].dones.nonzero().flatten()
    assert 0 < len(ids) < 4097
    for name in ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions"):
        getattr(terms[1], name)[ids] = 0.0
    mdps[1].compute_observations()
    mdps[2].compute_rewards(**kw[2])
    mdps[2].compute_observations()
    torch.cuda.synchronize()
    for m, t, what in ((mdps[1], terms[1], "split phases + host reset"), (mdps[2], terms[2], "rewards-only launch with the flag")):
        H.assert_equal(mdps[0].dones, m.dones, f"dones ({what})")
        H.assert_equal(mdps[0].policy_obs, m.policy_obs, f"policy obs ({what})")
        H.assert_equal(mdps[0].critic_obs, m.critic_obs, f"critic obs ({what})")
        for name in ("raw_actions", "prev_raw_actions", "prev_prev_raw_actions"):
            H.assert_equal(getattr(terms[0], name), getattr(t, name), f"{name} ({what})")
    # without the flag the post-reset row still shows the pre-reset action (the documented default of the read-only launch)
    plain = FusedMdp(env.to(cuda), spec, seed=9).step(True, True)
    col, width, _ = _last_action_columns(spec)
    assert float(plain.critic_obs[ids, col:col + width].abs().max()) > 0.0
    assert float(mdps[0].critic_obs[ids, col:col + width].abs().max()) == 0.0
