#!/usr/bin/env python
"""Benchmark of the LocoTouch hot path (BASELINE.json metric: env-steps/s, MDP + tactile + GAE + PPO, 4096 envs/GPU).

    python bench.py --gpus 1 --steps 10 --warmup 3                       # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1       # the reference algorithm on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one PPO iteration at 4096 envs per GPU: 24 x [actor-critic act, action pre-processing, fused MDP step of
the RandCylinderTransportTeacher task (23 active rewards incl. the gait term, 6 terminations, 2 x 348-D observations),
binary taxel synthesis + delay line, transition store] + GAE + 5 epochs x 4 mini-batches of PPO update.  Synthetic state
tensors stand in for PhysX (6 distinct pinned host sets per rank, 48 device-resident sets ~2.7 GB, cycled: larger than L2).
value = (N_gpus * 4096 * 24) / seconds per step, device-timed (CUDA events, max over ranks), inputs resident in HBM,
the whole step replayed from CUDA graphs.  e2e = the same iteration through HotPathEngine.replay(upload=True): every env
step's state set is copied host -> device from pinned memory inside the timed region (double-buffered on a copy stream) and the
iteration's metrics are read back; the eager per-step-upload variant is reported next to it.  Several GPUs: envs are sharded, the
flat PPO gradient is exchanged by peer loads inside the optimizer kernel (K14; NCCL all-reduce as fallback).
Prints ONE JSON line (rank 0) on stdout.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
T_STEPS = 24
METRIC = "env-steps/s (MDP+tactile+GAE+PPO) at 4096 envs/GPU"
WORKLOAD = ("PPO iteration, Isaac-RandCylinderTransportTeacher-LocoTouch-v1 terms (23 rewards incl. gait, 6 terminations, 2x348 obs) "
            "+ binary taxels (221) + GAE + PPO update (5 epochs x 4 mini-batches, MLP [512,256,128]), 4096 envs x 24 steps per GPU")

# algorithmic bytes per unit (SURVEY.md 8d / DESIGN.md section 4)
ALG_BYTES = {"mdp_step[teacher]": 6919, "mdp_step[locomotion]": 5579, "taxel_synth": 8840}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p["hbm_gbs"], p.get("bf16_tflops_sustained", 1363.9), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  The sampler process is started ahead of the
    region (nvidia-smi needs ~100 ms to come up) and polls every 20 ms; only rows read between mark_start() and mark_end()
    are summarised."""

    def __init__(self, index: int = 0):
        self.index, self.rows, self.proc, self.t0, self.t1 = index, [], None, None, None

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def wait_ready(self, timeout: float = 4.0):
        """Blocks until nvidia-smi has delivered its first row (it needs 0.1-1 s to come up on a fresh box)."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.01)

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *exc):
        if self.t1 is None:
            self.t1 = time.perf_counter()
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        t0 = self.t0 if self.t0 is not None else 0.0
        rows = [r for t, r in self.rows if t0 <= t <= self.t1 + 0.03]
        if not rows:  # a region shorter than the polling period: the rows nearest to it (the GPU is under the same load right around it)
            rows = [r for t, r in self.rows if t0 - 0.1 <= t <= self.t1 + 0.1]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------ CPU baseline
def bench_config(envs: int = ENVS_PER_GPU):
    """The workload description BOTH arms print (identical dicts: the driver compares them)."""
    return {"workload": WORKLOAD, "envs_per_gpu": envs, "steps_per_env": T_STEPS, "ppo": "5 epochs x 4 mini-batches, adaptive lr, clip 0.2, MLP [512,256,128] x 2",
            "inputs": "synthetic Go1 + cylinder state sets (seeded), larger than L2 in aggregate"}


class CpuReference:
    """The reference algorithm for the whole step (oracle port of the reference's torch code, torch CPU fp32, all host threads):
    one call of ``iteration()`` = 24 x [actor-critic act, MDP terms + observations, binary taxels + delay line, time-out bootstrap,
    store] + GAE + PPO update of 5 epochs x 4 mini-batches -- a WHOLE iteration, nothing extrapolated."""

    def __init__(self, envs: int = ENVS_PER_GPU, threads: int | None = None):
        from oracle import ppo as OP
        from oracle import tactile as OT
        from oracle import mdp as OM
        from oracle.mdp import MdpOracle
        from locotouch_b200.mdp import task_spec as TS
        from locotouch_b200.sim import synth

        self.OP, self.OT, self.OM = OP, OT, OM
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.envs = envs
        self.spec = TS.teacher_spec()
        self.env = synth.make_env(envs, seed=0, with_object=True, with_tactile=True)
        self.oracle = MdpOracle(self.env, self.spec)
        D, A = self.spec.obs_dim, 12
        self.D, self.A = D, A
        self.g = torch.Generator().manual_seed(0)
        self.shapes = OP.actor_critic_shapes(D, D, A, [512, 256, 128], [512, 256, 128])
        n_params = sum(int(torch.tensor(s).prod()) for _, s in self.shapes)
        self.flat = torch.randn(n_params, generator=self.g) * 0.05
        self.flat[:A] = 1.0
        self.thr = 0.05 + (torch.rand(envs, 221, generator=self.g) * 0.02 - 0.01)
        self.delay = OT.TactileDelayOracle(envs, 442, 1, 2)
        self.obs = torch.zeros(envs, D)
        self.cobs = torch.zeros(envs, D)
        self.lr = 1e-3

    def iteration(self):
        OP, OT, g, env, spec, N, A, D = self.OP, self.OT, self.g, self.env, self.spec, self.envs, self.A, self.D
        params = OP.unflatten(self.flat, self.shapes)
        aw, ab, cw, cb = OP._split(params)
        T = T_STEPS
        st = dict(obs=torch.empty(T, N, D), critic_obs=torch.empty(T, N, D), actions=torch.empty(T, N, A), values=torch.empty(T, N, 1),
                  logp=torch.empty(T, N, 1), mu=torch.empty(T, N, A), sigma=torch.empty(T, N, A), rewards=torch.empty(T, N, 1), dones=torch.empty(T, N, 1, dtype=torch.uint8))
        with torch.no_grad():
            for t in range(T):
                mu = OP.mlp_forward(self.obs, aw, ab)
                actions, logp = OP.act_sample(mu, params["std"], torch.randn(N, A, generator=g))
                values = OP.mlp_forward(self.cobs, cw, cb)
                st["obs"][t], st["critic_obs"][t], st["actions"][t], st["values"][t] = self.obs, self.cobs, actions, values
                st["logp"][t, :, 0], st["mu"][t], st["sigma"][t] = logp, mu, params["std"].expand_as(mu)
                # the action term of the env step (reference mdp/actions.py:30-52: process_actions ahead of the managers, reset(env_ids) of the
                # done envs between the reward and the observation pass) -- the work K0 / act_reset_on_done do inside K1 on the GPU arm
                term = env.action_manager.get_term("joint_pos")
                self.OM.action_term_process(term, actions, 100.0, 0.25, 1.0, env.scene["robot"].data.default_joint_pos)
                out = self.oracle.step(env, auto_reset=True, reset_action_term=True)
                self.obs, self.cobs = self.oracle.observe(env, u_noise=torch.rand(N, spec.obs_dim_per_step, generator=g), u_obj_euler=torch.rand(N, 3, generator=g))
                tac = OT.binary_taxels(env.scene["robot"].data.body_quat_w[:, 17:], env.scene.sensors["tactile_contact_sensor"].data.net_forces_w, self.thr,
                                       torch.rand(N, 221, generator=g), torch.rand(N, 221, generator=g))
                self.delay.record(tac["signal"])
                self.delay.get()
                st["rewards"][t, :, 0] = OP.bootstrap_rewards(out["reward"], values, out["time_outs"], 0.99)
                st["dones"][t, :, 0] = out["done"].to(torch.uint8)
            last_values = OP.mlp_forward(self.cobs, cw, cb)
            returns, adv = OP.gae_returns(st["rewards"], st["values"], st["dones"], last_values, 0.99, 0.95, True)
        flat_st = {k: st[k].flatten(0, 1) for k in ("obs", "critic_obs", "actions", "values", "logp", "mu", "sigma")}
        flat_st["returns"], flat_st["advantages"] = returns.flatten(0, 1), adv.flatten(0, 1)
        perm = torch.randperm(T * N, generator=g)
        _, _, _, lrs, flat = OP.ppo_update(self.flat, self.shapes, flat_st, perm, num_learning_epochs=5, num_mini_batches=4, clip_param=0.2, value_loss_coef=1.0,
                                           entropy_coef=0.01, learning_rate=self.lr, max_grad_norm=1.0, desired_kl=0.01)
        self.flat, self.lr = flat, lrs[-1]


def time_cpu_reference(steps: int, warmup: int, envs: int = ENVS_PER_GPU):
    """(env-steps/s, seconds per iteration list, cores) over ``steps`` whole iterations after ``warmup`` untimed ones."""
    ref = CpuReference(envs)
    for _ in range(warmup):
        ref.iteration()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ref.iteration()
        times.append(time.perf_counter() - t0)
    per = statistics.median(times)
    return envs * T_STEPS / per, times, ref.threads


def run_reference(args, rank: int):
    if rank != 0:
        return
    t0 = time.perf_counter()
    value, times, cores = time_cpu_reference(args.steps, args.warmup, args.envs_per_gpu)
    wall = time.perf_counter() - t0
    sample = f"{args.steps} whole PPO iterations ({T_STEPS} env steps + GAE + 20 mini-batches at {args.envs_per_gpu} envs) after {args.warmup} warm-up, median"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * args.envs_per_gpu * T_STEPS / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(args.envs_per_gpu),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample, "iteration_s": times},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "sample_wall_s": wall,
        "note": "the reference is torch-CPU Python that needs IsaacLab to import as a package; its algorithm for this path is timed through the oracle port "
                "(oracle/*.py, pinned to the unmodified reference by tests/golden); one rank's workload whatever --gpus says (the reference is single-process)",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ kernel roofline
def ncu_traffic_bytes(kernel_substr: str, largest: int = 0):
    """DRAM bytes per launch (read + write) of the committed `ncu --set full` capture, or None.  ``largest`` > 0: mean over the
    launches with the most traffic only (a capture that holds the same kernel at two problem sizes: the larger one)."""
    path = next((q for q in (os.path.join(ROOT, "profiles", f) for f in ("r2x_ncu_full_summary.json", "r2f_ncu_full_summary.json", "r1c_k1_k2_ncu_full_summary.json", "r1_k1_k2_ncu_full_summary.json"))
                 if os.path.exists(q)), None)  # newest capture first
    if path is None:
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for rec in json.load(open(path)):
        if kernel_substr in rec.get("kernel", ""):
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = rec[key].split()
                tot += float(v) * unit[u]
            vals.append(tot)
    if largest > 0:
        vals = sorted(vals)[-largest:]
    return sum(vals) / len(vals) if vals else None


def kernel_rooflines(engine, reps: int = 96):
    """Per-launch duration of the two streaming kernels of the rollout, measured live: `reps` launches captured in one CUDA
    graph on the timing stream, cycling through the engine's state sets (6 x ~33 MB > L2), CUDA events around the replay."""
    from locotouch_b200 import ops
    from locotouch_b200.sim import synth
    from locotouch_b200.streams import graph_capture

    st = engine.alg.storage
    stream = torch.cuda.current_stream()
    out = {}

    def timed(launch, name, bytes_per_launch):
        nonlocal reps
        for k in range(engine.K):
            launch(k)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with graph_capture(g):
            for r in range(reps):
                launch(r % engine.K)
        for _ in range(3):
            g.replay()
        best = 1e30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            e0.record()
            g.replay()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        out[name] = (best, bytes_per_launch)

    def mdp(k):
        engine._bind(k)
        t = k % engine.T
        engine.mdp.step(True, True, policy_in=st._obs_buf[t], critic_in=st._priv_buf[t], policy_out=st._obs_buf[t + 1], critic_out=st._priv_buf[t + 1],
                        step_offset=k, offset_base=engine.step_counter)

    timed(mdp, "mdp_step[teacher]", ALG_BYTES["mdp_step[teacher]"] * engine.N)
    if engine.tactile:
        def tax(k):
            env = engine.envs[k]
            ops.taxel_synth(env.scene["robot"].data.body_quat_w, env.scene.sensors["tactile_contact_sensor"].data.net_forces_w, engine.taxel_thr,
                            quat_body_offset=synth.NUM_ROBOT_BODIES, seed=1, offset=k, offset_base=engine.step_counter, signal=None, want_signal=False,
                            packed=engine.taxel_packed, delay_ring=engine.taxel_ring, delay_first=engine.taxel_first, delay_steps=engine.taxel_delay,
                            delayed_signal=engine.tactile_obs)
        timed(tax, "taxel_synth", ALG_BYTES["taxel_synth"] * engine.N)

    # the whole env step as the rollout graph runs it: K0 (action term) -> K1 (fused MDP step) with K2 (taxels + delay line) on its
    # parallel branch -> K3 (store into the RolloutStorage slot); algorithmic bytes = K1 + K2 (+ the [N, 12] action tensors and scalars)
    alg = engine.alg
    actions = st.actions[0]

    def env_step(k):
        t = k % engine.T
        st.step = t
        _, rew, dn, inf = engine.env_step(t, actions, 0)
        alg.transition.observations, alg.transition.critic_observations = st._obs_buf[t], st._priv_buf[t]
        alg.transition.actions, alg.transition.values = st.actions[t], st.values[t]
        alg.transition.actions_log_prob, alg.transition.action_mean, alg.transition.action_sigma = st.actions_log_prob[t], st.mu[t], st.sigma[t]
        alg.process_env_step(rew, dn, inf)

    step_bytes = (ALG_BYTES["mdp_step[teacher]"] + (ALG_BYTES["taxel_synth"] if engine.tactile else 0) + 4 * 12 * 6 + 32) * engine.N
    reps, saved_reps = 24, reps
    timed(env_step, "env step: K0 + K1 || K2 (+ delay line) + K3 store, as in the rollout graph", step_bytes)
    reps = saved_reps
    st.step = 0
    return out


def gemm_rows(engine, peak_tflops: float):
    """The tensor-core side of one PPO mini-batch (BASELINE.md "Tensor-core side"): forward, heads + loss and explicit backward of both
    MLPs at the mini-batch size exactly as PPO.minibatch_grads runs them, timed as one CUDA graph; FLOPs = 6 x B x (weights of both
    MLPs) (forward 2, dgrad 2, wgrad 2 per multiply-accumulate; the first layer has no dgrad)."""
    from locotouch_b200.streams import graph_capture

    ac = engine.alg.actor_critic
    alg = engine.alg
    B = engine.N * T_STEPS // 4
    saved, saved_lr, saved_acc = alg.optimizer.flat.clone(), alg.optimizer.lr_t.clone(), alg._loss_accum.clone()
    if getattr(alg, "_mb", None) is None:
        alg.update_begin(engine.perm)

    def run(i):  # the production mini-batch pass: K19 (forward of both MLPs), K16 (heads + loss + head dgrad), K12 dgrad x 4, K15 pair x 3 + 2 heads
        alg.minibatch_grads(i % alg.num_mini_batches)

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    reps = 12
    with graph_capture(g):
        for i in range(reps):
            run(i)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    alg.optimizer.flat.copy_(saved)
    alg.optimizer.lr_t.copy_(saved_lr)   # the loss kernel takes the adaptive learning-rate decision on the device
    alg._loss_accum.copy_(saved_acc)
    import torch.nn as nn

    macs, first = 0, 0
    for net in (ac.actor, ac.critic):
        lin = [m for m in net if isinstance(m, nn.Linear)]
        macs += sum(m.in_features * m.out_features for m in lin)
        first += lin[0].in_features * lin[0].out_features
    flops = 2.0 * B * (3 * macs - first)
    tfs = flops / best / 1e6
    tf32_peak = peak_tflops / 2.0  # TF32 dense = half the bf16 rate on B200 (2.25 vs 1.1 PFLOP/s nominal); measured bf16 peak / 2
    return [{"kernel": "actor-critic MLPs fwd + heads/loss + bwd of one mini-batch as PPO.minibatch_grads runs it (K19 + K16 + K12 dgrad + K15, TF32 tcgen05)", "bound": "tensor", "achieved": tfs, "peak": tf32_peak,
             "unit": "TFLOP/s", "frac": tfs / tf32_peak, "us_per_launch": best, "flops_per_launch": flops,
             "peak_note": "measured sustained bf16 cuBLAS peak / 2 (TF32 runs at half the bf16 rate)"}]


def replica_check(engine, world: int):
    """All ranks must hold bit-identical parameters, Adam moments and learning rate (they take the same summed gradient in the same
    order): all-reduced min == max of an integer checksum of the fp32 bit patterns."""
    import torch.distributed as dist

    opt = engine.alg.optimizer
    sums = torch.stack([opt.flat.view(torch.int32).to(torch.int64).sum(), opt.exp_avg_sq.view(torch.int32).to(torch.int64).sum(),
                        opt.lr_t.view(torch.int32).to(torch.int64).sum()])
    if world == 1:
        return {"identical": True, "world": 1}
    lo, hi = sums.clone(), sums.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = (lo == hi).tolist()
    return {"identical": all(same), "params": same[0], "adam_state": same[1], "learning_rate": same[2], "world": world}


def diagnose_replicas(args, device, rank, world):
    """Eager PPO iterations with the trace the adaptive schedule saw: per mini-batch (local KL, learning rate after the decision,
    gradient norm).  --same-data: every rank holds rank 0's envs and random streams and normalises advantages per rank, so the summed
    gradient / W and the mean KL are EXACTLY the single-GPU values (x + x is exact, x / 2 is exact): the W-rank trace must equal the
    W = 1 trace bit for bit.  Without it: the ordinary env-sharded job (different envs per rank, global advantage statistics)."""
    from locotouch_b200.engine import HotPathEngine

    eng = HotPathEngine(num_envs=args.envs_per_gpu, task="teacher", tactile=True, device=device, seed=0, num_state_sets=6,
                        data_rank=0 if args.same_data else None)
    if args.same_data:
        eng.alg.global_advantage_normalization = False
    eng.alg.trace = []
    out = []
    for it in range(args.diag_iters):
        torch.manual_seed(1234 + it)  # the same permutation on every rank and every world size
        eng.iteration()
        r = eng.read_results()
        out.append({"iteration": it, "lr_end": r["learning_rate"], "surrogate": r["surrogate_loss"], "value_loss": r["value_loss"], "entropy": r["entropy"],
                    "kl_local": [t[0] for t in eng.alg.trace], "lr_after": [t[1] for t in eng.alg.trace], "grad_norm": [t[2] for t in eng.alg.trace]})
        eng.alg.trace = []
    chk = replica_check(eng, world)
    csum = int(eng.alg.optimizer.flat.view(torch.int32).to(torch.int64).sum())
    if rank == 0:
        print(json.dumps({"diagnose_replicas": True, "world": world, "same_data": args.same_data, "peer_gradients": bool(eng.peer_gradients),
                          "replica_check": chk, "param_checksum": csum, "iterations": out}), flush=True)


# ------------------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-lite", action="store_true", help="for ncu launch lists: skip e2e / roofline / cpu legs")
    ap.add_argument("--diagnose-replicas", action="store_true",
                    help="eager iterations with the per-mini-batch KL / learning-rate trace; with --same-data every rank gets rank 0's data "
                         "and a W-rank run must reproduce the single-GPU trace")
    ap.add_argument("--same-data", action="store_true")
    ap.add_argument("--diag-iters", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if (args.impl == "ours" and not args.profile_lite) else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the locotouch_b200 product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator comes up: send it to stderr so that the JSON line is the
        # only thing rank 0 writes to stdout
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend="nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from locotouch_b200 import _C
    from locotouch_b200.engine import HotPathEngine

    N = args.envs_per_gpu
    if args.diagnose_replicas:
        diagnose_replicas(args, device, rank, world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    engine = HotPathEngine(num_envs=N, task="teacher", tactile=True, device=device, seed=0, num_state_sets=6, pin_host=True, prefetch=True)
    launches_before = _C.launch_count
    engine.capture()
    # ABI launches recorded into the graphs = launches of one iteration (2 warm-up iterations + 1 captured)
    launches_per_step = (_C.launch_count - launches_before) // 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            engine.replay()
        clocks.wait_ready()
        for _ in range(2):  # keep the GPU under load while the sampler's first rows arrive
            engine.replay()
        barrier()
        clocks.mark_start()
        start.record()
        for _ in range(args.steps):
            engine.replay()
        end.record()
        barrier()
        clocks.mark_end()
    ms = torch.tensor([start.elapsed_time(end)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = world * N * T_STEPS / (ms_per_step * 1e-3)
    metrics = engine.read_results()
    replicas = replica_check(engine, world)

    if args.profile_lite:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "ms_per_step": ms_per_step, "profile_lite": True}), flush=True)
        return
    # ---- end to end through the public drop-in classes, host buffers, copies inside the timed region
    # (a) graph replay + double-buffered prefetch: iteration i+1's T state sets upload from pinned host memory on a copy stream
    #     while iteration i computes; every timed iteration enqueues one full upload and reads its metrics back
    e2e_steps = max(1, args.e2e_steps)
    engine.replay(upload=True)
    engine.read_results()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        engine.replay(upload=True)
        engine.read_results()
    torch.cuda.synchronize()  # includes the copy stream: e2e_steps uploads are inside the timed region
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=device)
    barrier()
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * N * T_STEPS / float(e2e_s.item())
    # (b) the raw copy: T state sets, nothing else running
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    engine.prefetch_bank(0)
    torch.cuda.synchronize()
    h2d_gbs = T_STEPS * engine.upload_bytes / (time.perf_counter() - t0) / 1e9
    # (c) the same iteration through eager calls of the drop-in classes, one state-set upload per env step on the compute stream
    engine.iteration(upload=True)
    engine.read_results()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        engine.iteration(upload=True)
        engine.read_results()
    torch.cuda.synchronize()
    eager_value = N * T_STEPS / ((time.perf_counter() - t0) / 2)
    h2d = T_STEPS * engine.upload_bytes
    d2h = 4 * 6

    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (TF32 tensor-core GEMMs, as the reference)",
        "data": "synthetic",
        "config": bench_config(N),
        "setup": {"parallelism": (f"env-sharded dp{world}, " + ("gradients summed by peer loads over NVLink inside the optimizer kernel (K14)" if engine.peer_gradients
                                                                   else "NCCL all-reduce of flat PPO gradients")) if world > 1 else "single GPU",
                  "l2": "inputs larger than L2 (48 state sets ~2.7 GB + 290 MB rollout storage per rank)", "timing": "CUDA events on the launch stream around K graph replays, max over ranks",
                  "gemm_kernels": _C.gemm_backend()},
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "how": "HotPathEngine.replay(upload=True): the drop-in classes' calls replayed from CUDA graphs; the T state sets of iteration i+1 are "
                       "copied pinned-host -> device on a copy stream while iteration i computes (two device banks); metrics read back every iteration",
                "h2d_gbs_alone": h2d_gbs, "host_affinity": engine.host_affinity, "eager_per_rank": {"value": eager_value, "unit": "env-steps/s",
                                                         "how": "eager PPO/RolloutStorage/FusedMdp calls, one state-set upload per env step on the compute stream"}},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks.summary(),
        "iteration_metrics": metrics,
        "replica_check": replicas["identical"], "replica_detail": replicas,
    }
    if rank == 0:
        hbm, tf, kind = peaks()
        rl = kernel_rooflines(engine)
        rows = []
        for name, (us, nbytes) in rl.items():
            gbs = nbytes / us / 1e3
            rows.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "us_per_launch": us,
                         "alg_bytes_per_launch": nbytes})
        single = [r for r in rows if not r["kernel"].startswith("env step")]  # the headline fraction is one kernel's; the combined step rides along
        top = max(single, key=lambda r: r["us_per_launch"])
        line["roofline"] = {"bound": "hbm", "achieved": top["achieved"], "peak": hbm, "unit": "GB/s", "frac": top["frac"], "traffic": ncu_traffic_bytes("mdp_step_kernel"),
                            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the newest profiles/*_ncu_full_summary.json (one ncu --set "
                                            "full capture of this kernel, not re-measured in this run; the kernel's 12.6 MB of writes are still in L2 when it "
                                            "ends, so ncu counts few bytes written)",
                            "kernel": top["kernel"], "us_per_launch": top["us_per_launch"], "peak_kind": kind,
                            "how": "96 launches in one CUDA graph cycling 6 state sets (> L2), CUDA events on the launch stream"}
        step = [r for r in rows if r["kernel"].startswith("env step")]
        if step:  # K0 + K1 || K2 + K3 of one env step together (round-1 verdict: fewer, fatter launches): 65 MB per step
            line["roofline"]["env_step"] = {"achieved": step[0]["achieved"], "frac": step[0]["frac"], "us_per_step": step[0]["us_per_launch"],
                                            "alg_bytes_per_step": step[0]["alg_bytes_per_launch"]}
        try:  # the small kernels of the update + GAE, timed the same way (tools/kbench.py: graph of launches over > L2 of inputs)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import kbench as KB

            extra = {}
            extra.update(KB.bench_gae(N, 100))
            extra.update(KB.bench_ppo_loss(N * T_STEPS // 4, 100))
            extra.update(KB.bench_adam(engine.alg.optimizer.flat.numel(), 100))
            extra.update(KB.bench_k9(N * T_STEPS // 4, 80))
            extra.update(KB.bench_gather(N * T_STEPS, N * T_STEPS // 4, 20, obs_dim=engine.spec.obs_dim))
            extra.update(KB.bench_heads(N * T_STEPS // 4, 100))
            for name, (us, nbytes) in extra.items():
                gbs = nbytes / us / 1e3
                rows.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "us_per_launch": us,
                             "alg_bytes_per_launch": nbytes})
            # one row per GEMM of a mini-batch: fp32 operands make them HBM / L2 bound, so the row carries both views
            for name, (us, nbytes, flops) in KB.bench_gemms(N * T_STEPS // 4, 30, obs_dim=engine.spec.obs_dim).items():
                gbs = nbytes / us / 1e3
                rows.append({"kernel": name, "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "us_per_launch": us,
                             "alg_bytes_per_launch": nbytes, "tflops": flops / us / 1e6, "tf32_peak_tflops": tf / 2.0, "tensor_frac": flops / us / 1e6 / (tf / 2.0)})
            # K19 (both MLPs' hidden layers in one persistent kernel: rollout step and mini-batch forward) and the K15 pair launches
            for name, (us, nbytes, flops) in KB.bench_mlp3(N, N * T_STEPS // 4, 30, obs_dim=engine.spec.obs_dim).items():
                gbs = nbytes / us / 1e3
                rows.append({"kernel": name, "bound": "tensor" if name.startswith("K19") else "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                             "us_per_launch": us, "alg_bytes_per_launch": nbytes, "tflops": flops / us / 1e6, "tf32_peak_tflops": tf / 2.0,
                             "tensor_frac": flops / us / 1e6 / (tf / 2.0)})
            rows.extend(gemm_rows(engine, tf))
            # config C4 (distillation): the student's tactile pre-encoder (K17, fp32 FMA on CUDA cores: 432 kFLOP per frame) at the
            # distillation env count, the BASELINE size and 16 384 frames; "frac" = fraction of the HBM peak its 2 KB per frame amount to
            for name, (us, nbytes) in KB.bench_student_cnn(30).items():
                m = nbytes / ((442 + 64) * 4.0)
                rows.append({"kernel": name, "bound": "fp32 fma", "achieved": 2 * 216.0e3 * m / us / 1e6, "peak": None, "unit": "TFLOP/s", "frac": None,
                             "us_per_launch": us, "alg_bytes_per_launch": nbytes, "frames_per_s": m / us * 1e6})
            # the kernel with the largest share of the iteration (K19: ~35 % of the summed kernel time, profiles/r2m_launches_summary.md) next
            # to the streaming kernel the headline fraction is quoted on; without K19 (other widths) the largest K12 forward layer
            k19 = [r for r in rows if r["kernel"].startswith("K19") and "mini-batch" in r["kernel"]]
            k12 = k19 or [r for r in rows if r["kernel"].startswith("K12 forward")]
            if k12:
                big = max(k12, key=lambda r: r["us_per_launch"])
                line["roofline"]["dominant_by_share"] = {"kernel": big["kernel"], "bound": "tensor" if k19 else "hbm", "achieved": big["tflops"] if k19 else big["achieved"],
                                                         "peak": tf / 2.0 if k19 else hbm, "unit": "TFLOP/s" if k19 else "GB/s",
                                                         "frac": big["tensor_frac"] if k19 else big["frac"], "us_per_launch": big["us_per_launch"], "tflops": big["tflops"],
                                                         "tensor_frac_of_tf32_peak": big["tensor_frac"], "hbm_gbs": big["achieved"], "hbm_frac": big["frac"],
                                                         "traffic": ncu_traffic_bytes("mlp3_forward", largest=3) if k19 else ncu_traffic_bytes("device_kernel"),
                                                         "note": ("persistent fused MLP: activations stay in TMEM, so the algorithmic HBM bytes (x once, h1 / h2 / h3 once) are 0.36 of "
                                                                  "what three GEMM launches move and the kernel is bound by the TF32 tensor pipe + the shared-memory fill of the "
                                                                  "weight tiles; peak = measured bf16 cuBLAS peak / 2") if k19 else
                                                                 "fp32 operands: 105 FLOP/B against a ridge of ~100 (TF32) -- HBM / L2 bound, so the fraction is of the HBM peak"}
        except Exception as exc:  # noqa: BLE001 -- the table is diagnostics; the headline numbers above do not depend on it
            line["kernels_error"] = repr(exc)
        line["kernels"] = rows
        if world == 1 and not args.no_cpu_baseline:
            v, times, cores = time_cpu_reference(steps=2, warmup=1, envs=N)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"2 whole PPO iterations ({T_STEPS} env steps + GAE + 20 mini-batches at {N} envs) after 1 warm-up, median; nothing extrapolated",
                                    "iteration_s": times}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
