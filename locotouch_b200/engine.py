"""Hot-path engine: one PPO iteration of the LocoTouch learning path on one GPU (one process per GPU).

Sequence of reference ``OnPolicyRunner.learn`` (loco_rl/loco_rl/runners/on_policy_runner.py:150-222), with IsaacLab's
PhysX stepping replaced by pre-generated synthetic state sets (BASELINE.json north_star):

    for t in range(num_steps_per_env):                       HOT LOOP A
        actions = alg.act(obs, critic_obs)                   K19 hidden layers + K3b heads / sample -> rollout slot t
                                                             (K2 binary taxels + packed delay line of this step run beside it)
        [ env.step(actions) ]                                synthetic state set t % K; ONE launch: K0 action term (+ its reset of done
                                                             envs), K1 fused MDP step (obs written straight into slot t+1) and the K3
                                                             scalar store (bootstrapped reward + done flag into rollout row t)
        alg.process_env_step(rewards, dones, infos)          finds everything in its row: nothing to launch (LT_FUSE_K3=0: K3 store)
    alg.compute_returns(critic_obs)                          K4 GAE + advantage normalisation
    alg.update()                                             HOT LOOP B: K5 gather, K12 fwd / dgrad, K16 heads + loss,
                                                             K15 weight gradients, K7 clip+Adam (K14 across GPUs)

``iteration()`` runs it through the public drop-in classes (PPO / RolloutStorage / ActorCritic / FusedMdp);
``capture()`` records the same calls into CUDA graphs so that ``replay()`` has no Python or launch overhead.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import ops
from .loco_rl import PPO, ActorCritic
from .mdp import task_spec as TS
from .mdp.fused import FusedMdp
from .sim import synth
from .sim.scene import ActionTermState
from .streams import SideStream, graph_capture

# reference locotouch/config/locotouch/agents/rsl_rl_ppo_cfg.py:6-30
PPO_CFG = dict(num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0, entropy_coef=0.01,
               learning_rate=1.0e-3, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01)
NUM_STEPS_PER_ENV = 24
HIDDEN = (512, 256, 128)
# reference locotouch/config/base/locomotion_base_env_cfg.py:127-135
ACTION_CLIP, ACTION_RAW_SCALE = 100.0, 0.25


# What travels host -> device per env step.  Not uploaded at all: state tensors the hot path never reads.  Uploaded ONCE (at
# construction): the articulation constants (IsaacLab keeps them as persistent buffers, PhysX never rewrites them).  Uploaded as
# compact row subsets and scattered into place on the device: [N, num_bodies, .] tensors of which the path reads a few rows -- the
# four feet of body_pos_w / body_lin_vel_w, the 221 taxel bodies of body_quat_w.  Everything else is uploaded whole.
UNREAD = ("robot_contact_senosr.last_contact_time", "robot_contact_senosr.net_forces_w", "object_contact_sensor.last_air_time")
STATIC = ("robot.default_joint_pos", "robot.default_joint_vel", "robot.soft_joint_pos_limits")
FEET_ROWS_ONLY = ("robot.body_pos_w", "robot.body_lin_vel_w")
FEET = ("a_FR_foot", "b_FL_foot", "c_RR_foot", "d_RL_foot")


def stream_seed(seed: int, rank: int, consumer: int) -> int:
    """Philox key of one random-number consumer (0 = exploration noise, 1 = observation noise, 2 = taxel dropout / addition) of one
    rank: pairwise distinct for every (seed, rank, consumer), so no two consumers ever draw from the same counter blocks
    (splitmix64 finaliser over the packed triple; 63 bits so that it stays a non-negative int64 across the C ABI)."""
    x = ((seed & 0xFFFFFFFF) << 24) ^ ((rank & 0xFFFF) << 8) ^ (consumer & 0xFF)
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (x ^ (x >> 31)) & 0x7FFFFFFFFFFFFFFF


def row_subsets(env) -> dict[str, torch.Tensor]:
    """name -> int64 row (body) indices the hot path reads of that [N, num_bodies, .] tensor."""
    feet = torch.tensor(env.scene["robot"].find_bodies(list(FEET))[0], dtype=torch.int64)
    rows = {name: feet for name in FEET_ROWS_ONLY}
    quat = getattr(env.scene["robot"].data, "body_quat_w", None)
    if quat is None:
        return rows
    nb = quat.shape[1]
    if nb > synth.NUM_ROBOT_BODIES:  # taxel bodies follow the robot bodies (reference observations.py:154-159 reads only those)
        rows["robot.body_quat_w"] = torch.arange(synth.NUM_ROBOT_BODIES, nb, dtype=torch.int64)
    else:
        rows["robot.body_quat_w"] = torch.zeros(0, dtype=torch.int64)  # no tactile sensor: nothing reads body_quat_w
    return rows


def pack_host(env, pin: bool = False):
    """Lays every state tensor of a SynthEnv out in ONE contiguous host buffer (256-byte aligned slices) so that a single
    H2D copy of the first ``upload_bytes`` refreshes everything the hot path reads.  Returns (layout, compact, upload_bytes,
    flat_host_buffer); ``compact`` = [(full_name, offset, nbytes, dtype, shape)] of the row-subset tensors."""
    tensors = env.named_tensors()
    skip = {k for k in tensors if k.startswith("action.") or k in ("terminated", "time_outs")}
    subsets = {k: v for k, v in row_subsets(env).items() if k in tensors}
    layout, compact, off = [], [], 0

    def place(name, t, into):
        nonlocal off
        nbytes = t.numel() * t.element_size()
        into.append((name, off, nbytes, t.dtype, tuple(t.shape)))
        off += (nbytes + 255) // 256 * 256

    for name, t in tensors.items():  # uploaded whole, every step
        if name not in skip and name not in UNREAD and name not in STATIC and name not in subsets:
            place(name, t, layout)
    rows = {name: tensors[name][:, idx].contiguous() for name, idx in subsets.items() if idx.numel()}
    for name, t in rows.items():
        place(name, t, compact)
    upload_bytes = off
    for name, t in tensors.items():  # device-resident: copied once with the rest of the buffer at construction
        if name in UNREAD or name in STATIC or name in subsets:
            place(name, t, layout)
    host = torch.empty(off, dtype=torch.uint8, pin_memory=pin)
    for name, o, nbytes, dtype, shape in layout:
        host[o:o + nbytes].view(dtype).view(shape).copy_(tensors[name])
    for name, o, nbytes, dtype, shape in compact:
        host[o:o + nbytes].view(dtype).view(shape).copy_(rows[name])
    return layout, compact, upload_bytes, host


def device_set(env, layout, compact, host, device):
    """A device copy of the packed state set: (device_env whose tensors are views of the flat buffer, flat_device_buffer,
    [(full tensor, row indices, compact rows)] to scatter after an upload)."""
    flat = host.to(device)
    denv = env.to(device)
    views = {name: flat[o:o + nbytes].view(dtype).view(shape) for name, o, nbytes, dtype, shape in layout}
    denv.load_named_tensors(views)
    subsets = row_subsets(env)
    scatter = [(views[name], subsets[name].to(device), flat[o:o + nbytes].view(dtype).view(shape)) for name, o, nbytes, dtype, shape in compact]
    return denv, flat, scatter


def bind_host_thread_to_gpu(device: torch.device) -> str:
    """Pins the calling thread to the CPUs NVML names as local to ``device`` (same socket / PCIe root), so that the pinned host
    buffers allocated next land on the GPU's NUMA node and the upload threads of several ranks do not all sit on node 0.  Best effort:
    returns what happened; any failure (no NVML, restricted cpuset, ...) leaves the affinity untouched."""
    if os.environ.get("LT_NUMA_BIND", "1") == "0":
        return "off (LT_NUMA_BIND=0)"
    try:
        import pynvml

        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:  # noqa: BLE001
                handle = None
        if handle is None:  # no uuid in this torch build: NVML index == CUDA index unless CUDA_VISIBLE_DEVICES reorders
            if os.environ.get("CUDA_VISIBLE_DEVICES"):
                return "unchanged (cannot map the CUDA index to NVML under CUDA_VISIBLE_DEVICES)"
            handle = pynvml.nvmlDeviceGetHandleByIndex(device.index or 0)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = len(os.sched_getaffinity(0))
        return f"gpu-local cpus ({after} of {before})"
    except Exception as exc:  # noqa: BLE001
        return f"unchanged ({type(exc).__name__})"


class HotPathEngine:
    def __init__(self, num_envs: int = 4096, task: str = "teacher", tactile: bool = True, device="cuda:0", seed: int = 0,
                 num_state_sets: int = 6, num_steps: int = NUM_STEPS_PER_ENV, hidden=HIDDEN, ppo_cfg: dict | None = None,
                 pin_host: bool = False, tf32: bool = True, prefetch: bool = False, data_rank: int | None = None):
        self.device = torch.device(device)
        if self.device.type == "cuda":  # every kernel launch goes to the current device's stream (see _C.ptr)
            torch.cuda.set_device(self.device)
        self.N, self.T, self.K = num_envs, num_steps, num_state_sets
        self.spec = TS.SPECS[task]()
        self.tactile = tactile
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # which shard's synthetic data / random streams this rank uses: its own; diagnostics give every rank shard 0 so that a
        # W-rank run must reproduce the single-GPU run (bench.py --diagnose-replicas)
        drank = self.rank if data_rank is None else data_rank
        if tf32:  # reference locotouch/scripts/train.py:66-69
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.allow_tf32 = True
        # several ranks uploading from pinned host memory: keep every rank's thread and buffers on its GPU's NUMA node
        self.host_affinity = bind_host_thread_to_gpu(self.device) if (pin_host and self.world > 1 and self.device.type == "cuda") else "unchanged"
        # ---- synthetic state sets (stand-in for PhysX), one packed buffer each
        self.envs, self.dev_flat, self.host_flat = [], [], []
        base = synth.make_env(num_envs, seed=seed * 1000 + drank, with_object=self.spec.with_object, with_tactile=tactile,
                              max_episode_length=self.spec.max_episode_length)
        self.action_term = ActionTermState(num_envs, synth.NUM_JOINTS, self.device)
        # ``prefetch``: two banks of T device-resident sets, so that the H2D upload of the NEXT iteration's T state sets runs on
        # a copy stream while this iteration computes (the host keeps ``num_state_sets`` distinct pinned sets, cycled)
        self.prefetch = prefetch
        self.banks = 2 if prefetch else 1
        num_device_sets = 2 * num_steps if prefetch else num_state_sets
        layout = compact = None
        self.scatter = []
        self.feet_idx = torch.tensor(base.scene["robot"].find_bodies(list(FEET))[0], dtype=torch.int64, device=self.device)
        for k in range(num_device_sets):
            if k < num_state_sets:
                if k > 0:
                    synth.advance(base, keep_cmd_prob=0.9)
                layout, compact, self.upload_bytes, host = pack_host(base, pin=pin_host)
                self.host_flat.append(host)
            denv, flat, scatter = device_set(base, layout, compact, self.host_flat[k % num_state_sets], self.device)
            self.scatter.append(scatter)
            denv.action_manager._terms["joint_pos"] = self.action_term  # actions come from the policy, not from the set
            self.envs.append(denv)
            self.dev_flat.append(flat)
        self._iter = 0
        self._taxel_stream = SideStream(self.device)
        self._copy_stream = None
        self._upload_done = [None, None]
        self._rollout_done = [None, None]
        self.state_bytes = int(self.dev_flat[0].numel())
        self.default_joint_pos = self.envs[0].scene["robot"].data.default_joint_pos.clone()
        # ---- learner
        D = self.spec.obs_dim
        torch.manual_seed(seed)  # every rank starts from the same parameters
        ac = ActorCritic(D, D, synth.NUM_JOINTS, list(hidden), list(hidden), "elu", 1.0)
        cfg = dict(PPO_CFG)
        cfg.update(ppo_cfg or {})
        self.alg = PPO(ac, device=str(self.device), **cfg)
        self.alg.init_storage(num_envs, num_steps, [D], [D], [synth.NUM_JOINTS])
        # several ranks on one NVLink domain: gradients are exchanged by peer loads inside the optimizer kernel (K14) when the
        # symmetric-memory mapping can be set up; otherwise the NCCL all-reduce stays
        self.peer_gradients = self.alg.enable_peer_gradients() if self.world > 1 else False
        ac.seed = stream_seed(seed, drank, 0)
        self.step_counter = torch.zeros(1, device=self.device, dtype=torch.int64)  # device-resident env-step index
        self.fuse_action_term = os.environ.get("LT_FUSE_K0", "1") != "0"
        self.fuse_store = os.environ.get("LT_FUSE_K3", "1") != "0"
        self.taxel_after_policy = os.environ.get("LT_TAXEL_AFTER_POLICY", "1") != "0"
        # ActionManager.reset(env_ids) of the envs a step resets, inside the MDP launch (IsaacLab's order: rewards -> reset -> observations)
        self.reset_action_term = os.environ.get("LT_ACTION_RESET", "1") != "0"
        # ---- fused MDP: one instance, re-bound to the state set of each step
        self.mdp = FusedMdp(self.envs[0], self.spec, seed=stream_seed(seed, drank, 1))
        self.taxel_seed = stream_seed(seed, drank, 2)
        self.mdp_args = []  # per state set: a frozen copy of the argument block (pointers never change afterwards)
        # ---- tactile
        if tactile:
            g = torch.Generator().manual_seed(seed + 17)
            # contact_threshold 0.05 + U(-0.01, 0.01) per (env, taxel), sampled once (reference observations.py:121-126)
            self.taxel_thr = (0.05 + (torch.rand(num_envs, 221, generator=g) * 0.02 - 0.01)).to(self.device)
            self.taxel_ring = torch.zeros(num_envs, 2, 7, device=self.device, dtype=torch.int32)
            self.taxel_first = torch.ones(num_envs, device=self.device, dtype=torch.uint8)
            self.taxel_delay = torch.ones(num_envs, device=self.device, dtype=torch.int64)  # randint(1, 2) == 1
            self.tactile_obs = torch.zeros(num_envs, 442, device=self.device)
            self.taxel_packed = torch.zeros(num_envs, 7, device=self.device, dtype=torch.int32)
        self.perm = torch.zeros(num_envs * num_steps // cfg["num_mini_batches"] * cfg["num_mini_batches"], device=self.device, dtype=torch.int64)
        self.results = torch.zeros(8, device=self.device)  # mean reward, losses ... read back by the caller
        self._graphs = None
        self._initial_observation()

    # ------------------------------------------------------------------------------------------------------- pieces
    def _bind(self, k: int):
        self.mdp.env = self.envs[k]
        self.mdp._bound_ptrs = None

    def _initial_observation(self):
        st = self.alg.storage
        self._bind(0)
        self.mdp.compute_observations(policy_in=st._obs_buf[0], critic_in=st._priv_buf[0], policy_out=st._obs_buf[0], critic_out=st._priv_buf[0])

    def set_index(self, t: int, bank: int = 0) -> int:
        """Device state set read by env step t (of the iteration that uses ``bank``)."""
        return bank * self.T + t if self.prefetch else t % self.K

    def upload_state(self, k: int):
        """H2D refresh of device state set k from (pinned) host memory: what an env living on the host would have to do."""
        n = self.upload_bytes  # the per-step tensors the path reads; row subsets travel compact (see pack_host)
        self.dev_flat[k][:n].copy_(self.host_flat[k % self.K][:n], non_blocking=True)
        for full, idx, rows in self.scatter[k]:
            full.index_copy_(1, idx, rows)

    def prefetch_bank(self, bank: int):
        """Enqueues, on the copy stream, the H2D upload of the T state sets of ``bank``; it starts once the rollout that last
        read the bank has finished and overlaps whatever the compute stream is doing."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        if self._rollout_done[bank] is not None:
            cs.wait_event(self._rollout_done[bank])
        with torch.cuda.stream(cs):
            for t in range(self.T):
                self.upload_state(self.set_index(t, bank))
            ev = torch.cuda.Event()
            ev.record(cs)
        self._upload_done[bank] = ev

    def taxel_step(self, t: int, bank: int = 0, after=None):
        """K2 of env step ``t`` on the taxel stream.  It reads nothing the policy or the MDP step of this env step writes (sensor state
        of step t, the dones of step t - 1), so ``rollout_steps`` makes it a parallel branch of the POLICY step (``after`` = a mark taken
        before the policy; the launch itself is enqueued behind the policy kernels): K19 occupies 64 of the 148 SMs at 4096 envs and the
        taxel kernel fills the rest, instead of competing with the MDP step (one 1024-thread block per SM, no room for a taxel block)
        afterwards.  Enqueued AHEAD of K19 the same branch gained nothing: the first wave of taxel blocks covers every SM and K19's
        one-per-SM CTAs (all of an SM's shared memory) wait for them (rollout graph 1.53 ms; behind K19: 1.30 ms)."""
        k = self.set_index(t, bank)
        st = self.alg.storage
        if self.tactile:
            with self._taxel_stream.forked(after=after):
                env = self.envs[k]
                ops.taxel_synth(env.scene["robot"].data.body_quat_w, env.scene.sensors["tactile_contact_sensor"].data.net_forces_w,
                                self.taxel_thr, quat_body_offset=synth.NUM_ROBOT_BODIES, p_drop=0.005, p_add=0.005, seed=self.taxel_seed,
                                offset=t, offset_base=self.step_counter, signal=None, want_signal=False, packed=self.taxel_packed,
                                delay_ring=self.taxel_ring, delay_first=self.taxel_first, delay_steps=self.taxel_delay,
                                delayed_signal=self.tactile_obs,
                                # envs that were reset start a fresh delay line (reference replay_buffer.py:61 -> tactile_recorder.py:18-22):
                                # the dones of the previous env step, read from their RolloutStorage row (no flag |= dones launch)
                                delay_reset=st.dones[(t - 1) % self.T].view(-1))

    def env_step(self, t: int, actions: torch.Tensor, bank: int = 0, taxels_launched: bool = False):
        """Stand-in for ``env.step(actions)``: everything IsaacLab's managers would compute around PhysX."""
        k = self.set_index(t, bank)
        st = self.alg.storage
        a = self.action_term
        if not taxels_launched:
            self.taxel_step(t, bank)
        self._bind(k)
        # K3 inside K1: the time-out bootstrap + scalar rollout store of process_env_step written by the MDP launch into row t
        store = None
        if self.fuse_store:
            store = dict(rewards=st.rewards[t].view(-1), dones=st.dones[t].view(-1), values=st.values[t].view(-1), gamma=self.alg.gamma)
        if self.fuse_action_term:  # K0 inside K1: the action term's process_actions runs in the MDP launch, before any term reads it
            self.mdp.step(True, True, policy_in=st._obs_buf[t], critic_in=st._priv_buf[t], policy_out=st._obs_buf[t + 1], store=store,
                          critic_out=st._priv_buf[t + 1], step_offset=t, offset_base=self.step_counter, actions=actions,
                          action_term=dict(prev_prev_raw=a.prev_prev_raw_actions, processed=a.processed_actions, offset=self.default_joint_pos,
                                           clip=ACTION_CLIP, raw_scale=ACTION_RAW_SCALE, scale=1.0),
                          reset_action_term=self.reset_action_term)
        else:
            ops.process_actions(actions, a.raw_actions, a.prev_raw_actions, a.prev_prev_raw_actions, a.processed_actions,
                                clip=ACTION_CLIP, raw_scale=ACTION_RAW_SCALE, scale=1.0, offset=self.default_joint_pos)
            self.mdp.step(True, True, policy_in=st._obs_buf[t], critic_in=st._priv_buf[t], policy_out=st._obs_buf[t + 1],
                          critic_out=st._priv_buf[t + 1], step_offset=t, offset_base=self.step_counter, store=store,
                          action_term=dict(prev_prev_raw=a.prev_prev_raw_actions), reset_action_term=self.reset_action_term)
        if self.tactile:
            self._taxel_stream.join()
        if store is not None:  # rewards (bootstrapped) and dones already sit in their RolloutStorage row: process_env_step has nothing to launch
            return st._obs_buf[t + 1], store["rewards"], store["dones"], {"observations": {"critic": st._priv_buf[t + 1]}}
        return st._obs_buf[t + 1], self.mdp.reward_buf, self.mdp.dones, {"time_outs": self.mdp.time_outs, "observations": {"critic": st._priv_buf[t + 1]}}

    def rollout_steps(self, upload: bool = False, bank: int = 0):
        """HOT LOOP A: the T env steps (no collective inside: capturable for any world size)."""
        alg, st = self.alg, self.alg.storage
        ac = alg.actor_critic
        ac._offset_base = self.step_counter
        for t in range(self.T):
            if upload:
                self.upload_state(self.set_index(t, bank))
            ac._graph_slot = t
            if self.taxel_after_policy and self.tactile:
                # K2 depends on nothing of this env step, but it is ENQUEUED behind the policy kernels: the block scheduler serves grids in
                # launch order, so K19's 64 one-per-SM CTAs are resident before the 1024 taxel blocks arrive and the taxel kernel fills the
                # remaining SMs, instead of K19's CTAs waiting for a first wave of taxel blocks to leave the SMs they need whole
                mark = self._taxel_stream.mark()
                actions = alg.act(st._obs_buf[t], st._priv_buf[t])
                self.taxel_step(t, bank, after=mark)
            else:
                self.taxel_step(t, bank)
                actions = alg.act(st._obs_buf[t], st._priv_buf[t])
            obs, rewards, dones, infos = self.env_step(t, actions, bank, taxels_launched=True)
            alg.process_env_step(rewards, dones, infos)
        ops.counter_add(self.step_counter, self.T)

    def rollout_finish(self):
        """GAE (+ the 3-double advantage-statistics all-reduce when several ranks train together)."""
        st = self.alg.storage
        self.alg.compute_returns(st._priv_buf[self.T])
        torch.sum(st.rewards, dim=(0, 1, 2), out=self.results[0])

    def rollout(self, upload: bool = False):
        self.rollout_steps(upload)
        self.rollout_finish()

    def finish_iteration(self):
        """Observation after the last transition becomes the first observation of the next rollout."""
        st = self.alg.storage
        st._obs_buf[0].copy_(st._obs_buf[self.T])
        st._priv_buf[0].copy_(st._priv_buf[self.T])

    def draw_permutation(self):
        torch.randperm(self.perm.numel(), device=self.device, out=self.perm)

    # --------------------------------------------------------------------------------------------------- public API
    def iteration(self, upload: bool = False):
        """One PPO iteration through the drop-in classes (eager).  ``upload``: refresh every step's state set from host
        memory inside the step (the end-to-end measurement)."""
        self.rollout(upload=upload)
        self.draw_permutation()
        losses = self.alg.update(indices=self.perm)
        self.finish_iteration()
        return losses

    def capture(self, split: bool | None = None):
        """Records the iteration into CUDA graphs.  One process: rollout + GAE in one graph, the whole update in another.
        Several processes (``split``): NCCL collectives stay OUTSIDE the graphs -- graphs are the T env steps, the gather,
        and one forward/loss/backward graph per mini-batch slice; GAE, the all-reduces and clip+Adam run eagerly between."""
        alg = self.alg
        # With the peer-memory gradient exchange (K14) the update holds no NCCL call -- cross-GPU barrier and peer-sum kernel are
        # plain launches -- so several processes capture it as ONE graph like a single process does; only the advantage-statistics
        # all-reduce of the rollout's tail stays eager.
        whole_update = self.world > 1 and self.peer_gradients and split is None and os.environ.get("LT_PEER_GRAPH", "1") != "0"
        # ... and with the advantage statistics in symmetric memory too (barrier + peer loads instead of an NCCL all-reduce) the
        # rollout's tail is capturable as well: the whole iteration is two graphs per rank, exactly like a single process
        tail_in_graph = whole_update and self.alg._peer.get("adv_stats") is not None and os.environ.get("LT_PEER_STATS_GRAPH", "1") != "0"
        split = (self.world > 1) if split is None else split
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):  # warm-up on the side stream (cuBLAS workspaces, autograd buffers)
                self.rollout()
                self.draw_permutation()
                alg.update_body(self.perm)
                self.finish_iteration()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g_roll = []
        for bank in range(self.banks):
            alg.storage.clear()  # host-side slot counter: every bank records the same T slots
            g = torch.cuda.CUDAGraph()
            with graph_capture(g):
                self.rollout_steps(bank=bank)
                if not split or tail_in_graph:
                    self.rollout_finish()
            g_roll.append(g)
        if not split or whole_update:
            g_upd = torch.cuda.CUDAGraph()
            with graph_capture(g_upd):
                alg.update_body(self.perm)
                self.finish_iteration()
            self._graphs = dict(split=False, roll=g_roll, update=g_upd, finish_eager=split and not tail_in_graph)
        else:
            g_begin = torch.cuda.CUDAGraph()
            with graph_capture(g_begin):
                alg.update_begin(self.perm)
            g_mb = []
            for i in range(alg.num_mini_batches):
                g = torch.cuda.CUDAGraph()
                with graph_capture(g):
                    alg.minibatch_grads(i)
                g_mb.append(g)
            g_tail = torch.cuda.CUDAGraph()
            with graph_capture(g_tail):
                alg.step_after_reduce()
            self._graphs = dict(split=True, roll=g_roll, begin=g_begin, mb=g_mb, tail=g_tail)
        alg.storage.clear()
        torch.cuda.synchronize()
        return self

    def replay(self, upload: bool = False):
        """One iteration from the captured graphs.  ``upload`` (prefetch engines): this iteration's T state sets come from
        pinned host memory -- they were enqueued on the copy stream during the previous iteration -- and the next
        iteration's upload is enqueued before this iteration's kernels, so copies and compute overlap."""
        g = self._graphs
        alg = self.alg
        bank = self._iter % self.banks
        self._iter += 1
        if upload:
            if not self.prefetch:
                raise RuntimeError("replay(upload=True) needs an engine built with prefetch=True")
            if self._upload_done[bank] is None:  # first call: nothing in flight yet
                self.prefetch_bank(bank)
            torch.cuda.current_stream().wait_event(self._upload_done[bank])
            self._upload_done[bank] = None
            self.prefetch_bank(1 - bank)
        g["roll"][bank].replay()
        if self.prefetch:
            ev = torch.cuda.Event()
            ev.record()
            self._rollout_done[bank] = ev
        if not g["split"]:
            if g.get("finish_eager"):
                self.rollout_finish()
            self.draw_permutation()
            g["update"].replay()
        else:
            self.rollout_finish()
            self.draw_permutation()
            g["begin"].replay()
            for _epoch in range(alg.num_learning_epochs):
                for i in range(alg.num_mini_batches):
                    g["mb"][i].replay()
                    alg.allreduce_grads()  # NCCL all-reduce, or (K14) the barrier in front of the peer reads of the tail graph
                    g["tail"].replay()
                    alg.after_step_barrier()
            self.finish_iteration()
        alg.storage.clear()

    def read_results(self):
        """Device -> host read of the iteration's metrics (mean step reward, mean losses, learning rate)."""
        losses = self.alg.update_epilogue()
        return dict(mean_reward=float(self.results[0].item()) / (self.T * self.N), value_loss=losses[0], surrogate_loss=losses[1], entropy=losses[2],
                    learning_rate=self.alg.learning_rate)
