"""Per-kernel micro-benchmark: achieved algorithmic GB/s of every hand-written kernel at BASELINE sizes.

    python tools/kbench.py [--envs 4096] [--reps 200] [--json out.json]

Each kernel is timed with CUDA events over ``reps`` launches captured in a CUDA graph (so that Python / ctypes launch
overhead is not what is measured), cycling through ``copies`` independent input sets whose total footprint exceeds the
126 MB L2 ("inputs larger than L2").  Algorithmic bytes per unit are the SURVEY.md section 8(d) figures.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from locotouch_b200 import ops  # noqa: E402
from locotouch_b200.mdp import task_spec as TS  # noqa: E402
from locotouch_b200.mdp.fused import FusedMdp  # noqa: E402
from locotouch_b200.sim import synth  # noqa: E402
from locotouch_b200.streams import graph_capture  # noqa: E402

L2_BYTES = 126e6


def peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def time_graph(launch, copies: int, reps: int, warmup: int = 3):
    """``launch(i)`` enqueues one kernel on input set i.  Returns mean microseconds per launch."""
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for i in range(copies):
            launch(i)
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with graph_capture(graph, stream=stream):
            for r in range(reps):
                launch(r % copies)
        for _ in range(warmup):
            graph.replay()
        stream.synchronize()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(5):
            start.record(stream)
            graph.replay()
            end.record(stream)
            end.synchronize()
            best = min(best, start.elapsed_time(end) * 1e3 / reps)
    return best


def copies_for(bytes_per_set: float) -> int:
    return max(2, int(L2_BYTES * 1.5 / max(bytes_per_set, 1)) + 1)


def bench_mdp(task: str, n: int, reps: int, phases=("fused",)):
    spec = TS.SPECS[task]()
    env_kw = dict(with_object=spec.with_object)
    per_env = {"locomotion": 5579, "teacher": 6919}[task]
    copies = min(copies_for(per_env * n), 16)
    mdps = []
    for i in range(copies):
        env = synth.make_env(n, seed=i, **env_kw).to("cuda")
        m = FusedMdp(env, spec, seed=i)
        m.exact_any_nonzero_cmd = True
        m.step(True, True)  # warm: fills history, publishes the any() flag
        mdps.append(m)
    torch.cuda.synchronize()
    out = {}

    def fused(i):
        mdps[i].step(True, True)

    def rew(i):
        mdps[i].compute_rewards()

    def obs(i):
        mdps[i].compute_observations()

    out[f"mdp_step[{task}]"] = (time_graph(fused, copies, reps), per_env * n)
    rew_bytes = {"locomotion": 1619, "teacher": 2015}[task] * n
    obs_bytes = per_env * n - rew_bytes + (276 if task == "locomotion" else 420) * n
    out[f"mdp_rewards[{task}]"] = (time_graph(rew, copies, reps), rew_bytes)
    out[f"mdp_obs[{task}]"] = (time_graph(obs, copies, reps), obs_bytes)
    return out


def bench_taxel(n: int, reps: int):
    per_env = 8840
    copies = min(copies_for(per_env * n), 12)
    sets = []
    for i in range(copies):
        g = torch.Generator().manual_seed(i)
        q = torch.randn(n, 238, 4, generator=g)
        q = (q / q.norm(dim=-1, keepdim=True)).cuda()
        f = (torch.randn(n, 221, 3, generator=g) * 0.1).cuda()
        thr = (0.05 + (torch.rand(n, 221, generator=g) - 0.5) * 0.02).cuda()
        sets.append((q, f, thr, torch.empty(n, 442, device="cuda"), torch.empty(n, 7, device="cuda", dtype=torch.int32),
                     torch.zeros(n, 2, 7, device="cuda", dtype=torch.int32), torch.ones(n, device="cuda", dtype=torch.uint8),
                     torch.ones(n, device="cuda", dtype=torch.int64), torch.empty(n, 442, device="cuda")))

    def plain(i):
        q, f, thr, sig, packed, *_ = sets[i]
        ops.taxel_synth(q, f, thr, quat_body_offset=17, seed=1, offset=i, signal=sig, packed=packed)

    def delayed(i):
        q, f, thr, sig, packed, ring, first, delay, dsig = sets[i]
        ops.taxel_synth(q, f, thr, quat_body_offset=17, seed=1, offset=i, signal=None, packed=packed, want_signal=False, delay_ring=ring,
                        delay_first=first, delay_steps=delay, delayed_signal=dsig)

    fsets = [torch.empty(n, 4, 221, device="cuda") for _ in range(copies)]

    def forces(i):
        q, f, thr, *_ = sets[i]
        ops.taxel_forces(q, f, thr, fsets[i], ("contact", "normalized", "minmax", "discretized"), quat_body_offset=17, p_drop=0.005, p_add=0.005,
                         add_force_noise=True, force_n_prop_min=-0.1, force_n_prop_max=0.1, maximal_force=3.0, total_levels=5,
                         add_level_noise=True, level_n_min=-1, level_n_max=1, seed=1, offset=i)

    extra = {"taxel_forces (4 channels)": (time_graph(forces, copies, reps), (221 * 32 + 4 * 221 * 4) * n)}
    return {**extra, "taxel_synth": (time_graph(plain, copies, reps), per_env * n),
            "taxel_synth+delay": (time_graph(delayed, copies, reps), (per_env + 28 * 3) * n)}


def bench_gae(n: int, reps: int, T: int = 24):
    per = 17 * T * n + 4 * n
    copies = min(copies_for(per), 64)
    sets = []
    for i in range(copies):
        g = torch.Generator().manual_seed(i)
        sets.append((torch.randn(T, n, generator=g).cuda(), torch.randn(T, n, generator=g).cuda(),
                     (torch.rand(T, n, generator=g) < 0.02).byte().cuda(), torch.randn(n, generator=g).cuda(),
                     torch.empty(T, n, device="cuda"), torch.empty(T, n, device="cuda")))

    def run(i):
        r, v, d, lv, ret, adv = sets[i]
        ops.gae(r, v, d, lv, 0.99, 0.95, True, ret, adv)

    return {"gae+normalize (1 launch at T = 24)": (time_graph(run, copies, reps), per + 8 * T * n)}


def bench_ppo_loss(b: int, reps: int, A: int = 12):
    per = 264 * b
    copies = min(copies_for(per), 32)
    sets = []
    for i in range(copies):
        g = torch.Generator().manual_seed(i)
        rn = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
        sets.append(dict(mu=rn(b, A), sigma=(0.5 + torch.rand(A, generator=g)).cuda(), value=rn(b), actions=rn(b, A), old_logp=rn(b),
                         old_mu=rn(b, A), old_sigma=(0.5 + torch.rand(b, A, generator=g)).cuda(), advantages=rn(b), returns=rn(b), old_values=rn(b)))
    bufs = [ops.PpoLossBuffers(b, A, "cuda") for _ in range(copies)]
    lr = torch.tensor([1e-3], device="cuda")

    def run(i):
        ops.ppo_loss(**sets[i], entropy_coef=0.01, desired_kl=0.01, lr=lr, buffers=bufs[i])

    return {"ppo_loss fwd+bwd": (time_graph(run, copies, reps), per)}


def bench_adam(n: int, reps: int):
    per = 32 * n
    copies = min(copies_for(per), 16)
    sets = [tuple(torch.randn(n, device="cuda") for _ in range(2)) + (torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")) for _ in range(copies)]
    lr, step = torch.tensor([1e-3], device="cuda"), torch.zeros(1, device="cuda")

    def run(i):
        p, g, m, v = sets[i]
        ops.clip_adam(p, g, m, v, lr, step)

    return {"clip+adam (1 launch)": (time_graph(run, copies, reps), per)}


def bench_k9(b: int, reps: int):
    out = {}
    for n in (512, 256, 128):
        per = 12 * b * n
        copies = min(copies_for(per), 8)
        sets = [(torch.randn(b, n, device="cuda"), torch.randn(b, n, device="cuda"), torch.empty(n, device="cuda")) for _ in range(copies)]

        def run(i, sets=sets):
            g, h, bg = sets[i]
            ops.bias_act_bwd(g, h, bg)

        out[f"bias_act_bwd [{b}x{n}]"] = (time_graph(run, copies, max(20, reps // 4)), per)
    return out


def bench_split_pad(n: int, reps: int, T: int = 24, D: int = 348):
    """K10 at the teacher observation size: reads the valid rows once, writes [T, M, D] (zero padding included)."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(T, n, D, generator=g).cuda()
    dones = (torch.rand(T, n, 1, generator=g) < 0.02).cuda()
    index = ops.TrajectoryIndex(dones)
    per = T * n * D * 4 + T * index.M * D * 4

    def run(i):
        index.split_and_pad(x, want_masks=True)

    return {f"split_and_pad [24x{n}x{D}] (M={index.M})": (time_graph(run, 2, max(10, reps // 10)), per)}


def bench_gather(n_rows: int, count: int, reps: int, obs_dim: int = 270):
    dims = (obs_dim, obs_dim, 12, 1, 1, 1, 1, 12, 12)
    per = sum(dims) * 4 * 2 * count
    src = [torch.randn(n_rows, d, device="cuda") for d in dims]
    copies = 4
    outs = [[torch.empty(count, d, device="cuda") for d in dims] for _ in range(copies)]
    idx = [torch.randperm(n_rows, device="cuda")[:count].contiguous() for _ in range(copies)]

    def run(i):
        ops.gather_rows(src, idx[i], outs[i])

    return {"gather_rows (9 tensors)": (time_graph(run, copies, reps), per)}


def bench_command(n: int, reps: int):
    """K13: [IL] CommandTerm.reset (all envs marked) + compute per env step, and the single-block curriculum decision."""
    from types import SimpleNamespace

    from locotouch_b200.mdp.commands import UniformVelocityCommandGaitLoggingMultiSampling
    from locotouch_b200.mdp.curriculums import ModifyVelCommandsRangeBasedonReward
    from locotouch_b200.sim.scene import SceneEntityCfg

    env = synth.make_env(n, seed=0).to("cuda")
    env.episode_length_buf[:] = 400
    vla = torch.rand(n, 4, device="cuda")
    sums = {"track_lin_vel_xy": torch.rand(n, device="cuda") * 20, "track_ang_vel_z": torch.rand(n, device="cuda") * 10}
    env.reward_manager._episode_sums = sums
    env.reward_manager.set_term_cfg("track_lin_vel_xy", SimpleNamespace(weight=1.0, params={"sigma": 0.25}))
    env.reward_manager.set_term_cfg("track_ang_vel_z", SimpleNamespace(weight=0.5, params={"sigma": 0.25}))
    env.reward_manager.set_term_cfg("gait", SimpleNamespace(func=SimpleNamespace(valid_last_air_time=vla)))
    cfg = SimpleNamespace(asset_name="robot", resampling_time_range=(8.0, 8.0), heading_command=False, ranges=SimpleNamespace(
        lin_vel_x=(-0.2, 0.2), lin_vel_y=(-0.1, 0.1), ang_vel_z=(-0.3, 0.3)), new_command_probs=0.15, rel_standing_envs=0.1,
        final_rel_standing_envs=0.05, initial_zero_command_steps=0, final_initial_zero_command_steps=50, binary_maximal_command=False,
        sensor_cfg=SceneEntityCfg("robot_contact_senosr", body_names=".*foot"))
    term = UniformVelocityCommandGaitLoggingMultiSampling(cfg, env)
    env.command_manager._terms["base_velocity"] = term
    params = dict(command_name="base_velocity", command_maximum_ranges=[0.5, 0.25, 0.785], curriculum_bins=[20, 20, 20], reset_envs_episode_length=0.98,
                  reward_name_lin="track_lin_vel_xy", reward_name_ang="track_ang_vel_z", error_threshold_lin=0.08, error_threshold_ang=0.1,
                  repeat_times_lin=1, repeat_times_ang=1, max_distance_bins=4)
    cur = ModifyVelCommandsRangeBasedonReward(SimpleNamespace(params=params), env)
    mask = torch.rand(n, device="cuda") < 0.02
    term._set_mask(mask)
    a_reset, a_comp = term._args(1, 0.0, None), term._args(2, 0.02, None)
    import ctypes as C

    from locotouch_b200 import _C

    def reset(i):
        _C.check(_C.lib().lt_command_step(C.byref(a_reset), _C.current_stream()))

    def compute(i):
        _C.check(_C.lib().lt_command_step(C.byref(a_comp), _C.current_stream()))

    def curriculum(i):
        cur(env, mask)

    per_reset = n * (14 * 4 * 0.02 + 1 + 8 + 12 + 12)  # mask, episode length, command rows (+ the metric rows of the 2 % reset envs)
    per_comp = n * (12 + 24 + 16 + 16 + 4 + 8 + 12 + 1 + 3 * 4 + 4 + 12 + 11 * 4)
    return {"command reset (2 % of envs)": (time_graph(reset, 1, reps), per_reset), "command compute (2 launches)": (time_graph(compute, 1, reps), per_comp),
            "velocity curriculum (1 block)": (time_graph(curriculum, 1, reps), n * (1 + 2 * (1 + 4 + 4)))}


def bench_copy(sizes_mb, reps: int):
    """Launch-size calibration: a plain device-to-device copy (cudaMemcpyAsync through torch) that moves the same number of bytes
    (half read, half written) as a kernel's algorithmic traffic, timed the same way.  What fraction of the 4 GB copy peak a launch
    of this size can reach at all (launch latency + ramp) -- the ceiling the small-launch fractions above should be read against."""
    out = {}
    for mb in sizes_mb:
        half = int(mb * 1e6 / 2) // 16 * 16
        copies = min(copies_for(2 * half), 64)
        sets = [(torch.empty(half, dtype=torch.uint8, device="cuda"), torch.empty(half, dtype=torch.uint8, device="cuda")) for _ in range(copies)]

        def run(i, sets=sets):
            sets[i][1].copy_(sets[i][0])

        out[f"d2d copy calibration [{mb:g} MB]"] = (time_graph(run, copies, reps), 2 * half)
    return out


def bench_heads(b: int, reps: int, A: int = 12, H: int = 128):
    """K16: heads + loss + head dgrad.  Algorithmic bytes per sample: both hidden rows read, both gradient rows written, the 264 B
    of rollout row, dL/dmu and dL/dV written for the head weight gradients."""
    per = (4 * H * 4 + 264 + (A + 1) * 4) * b
    copies = min(copies_for(per), 8)
    sets = []
    for i in range(copies):
        g = torch.Generator().manual_seed(i)
        rn = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
        sets.append((rn(b, H), rn(b, H), rn(A, H) * 0.1, rn(A), rn(1, H) * 0.1, rn(1), (0.5 + torch.rand(A, generator=g)).cuda(), rn(b, A), rn(b), rn(b, A),
                     (0.5 + torch.rand(b, A, generator=g)).cuda(), rn(b), rn(b), rn(b), torch.empty(b, H, device="cuda"), torch.empty(b, H, device="cuda")))
    bufs = [ops.PpoLossBuffers(b, A, "cuda") for _ in range(copies)]
    lr = torch.tensor([1e-3], device="cuda")

    def run(i):
        ops.ppo_heads_loss(*sets[i], entropy_coef=0.01, desired_kl=0.01, lr=lr, buffers=bufs[i])

    return {f"heads+loss+dgrad K16 [{b}x{H}]": (time_graph(run, copies, reps), per)}


def bench_gemms(b: int, reps: int, obs_dim: int = 348, hidden=(512, 256, 128), A: int = 12):
    """Per-GEMM rows of one mini-batch: K12 forward (bias + ELU fused), K12 dgrad (ELU backward fused), K15 weight + bias gradient.
    Each row: (us, algorithmic bytes = operands read once + result written once, flops)."""
    out = {}
    dims = [obs_dim] + list(hidden)
    torch.backends.cuda.matmul.allow_tf32 = True
    for k, n in zip(dims[:-1], dims[1:]):
        copies = 4
        xs = [torch.randn(b, k, device="cuda") for _ in range(copies)]
        w, bias = torch.randn(n, k, device="cuda") / k ** 0.5, torch.randn(n, device="cuda")
        hs = [torch.empty(b, n, device="cuda") for _ in range(copies)]
        out[f"K12 forward {b}x{k}->{n}"] = (time_graph(lambda i: ops.linear_bias_act(xs[i], w, bias, out=hs[i], elu=True), copies, reps),
                                            4.0 * (b * k + n * k + b * n), 2.0 * b * k * n)
        gs = [torch.randn(b, n, device="cuda") for _ in range(copies)]
        dw, db = torch.zeros(n, k, device="cuda"), torch.zeros(n, device="cuda")
        out[f"K15 wgrad+bias {n}x{k} over {b}"] = (time_graph(lambda i: ops.wgrad(gs[i], xs[i], dw, db, zero_first=False), copies, reps),
                                                   4.0 * (b * (n + k) + n * k), 2.0 * b * k * n)
        if k != obs_dim:  # the input layer needs no dgrad
            gin = [torch.empty(b, k, device="cuda") for _ in range(copies)]
            out[f"K12 dgrad {b}x{n}->{k}"] = (time_graph(lambda i: ops.dgrad_act_bwd(gs[i], w, xs[i], out=gin[i]), copies, reps),
                                              4.0 * (b * n + n * k + 2 * b * k), 2.0 * b * k * n)
    for n in (A, 1):
        k = hidden[-1]
        copies = 4
        xs = [torch.randn(b, k, device="cuda") for _ in range(copies)]
        gs = [torch.randn(b, n, device="cuda") for _ in range(copies)]
        dw, db = torch.zeros(n, k, device="cuda"), torch.zeros(n, device="cuda")
        out[f"K15 head wgrad+bias {n}x{k} over {b}"] = (time_graph(lambda i: ops.wgrad(gs[i], xs[i], dw, db, zero_first=False), copies, reps),
                                                        4.0 * (b * (n + k) + n * k), 2.0 * b * k * n)
    return out


def bench_mlp3(n_envs: int, b: int, reps: int, obs_dim: int = 348):
    """K19: the three hidden layers of actor AND critic in one persistent tcgen05 kernel, at the rollout batch (one env step, h3 only)
    and at the mini-batch size (h1 / h2 stored for the backward); next to it the K15 pair launches (layer i of both networks).
    Each row: (us, algorithmic bytes = x read once + weights + stored activations, flops)."""
    out = {}
    if obs_dim & 3 or obs_dim > 352:
        return out
    macs = obs_dim * 512 + 512 * 256 + 256 * 128
    wbytes = 4.0 * (macs + 512 + 256 + 128)
    for rows_, keep, tag in ((n_envs, False, "rollout step"), (b, True, "mini-batch forward")):
        copies = 4
        sets = []
        for _ in range(copies):
            pair = []
            for _ in range(2):
                x = torch.randn(rows_, obs_dim, device="cuda")
                ps = []
                for (n, k) in ((512, obs_dim), (256, 512), (128, 256)):
                    ps += [torch.randn(n, k, device="cuda") / k ** 0.5, torch.randn(n, device="cuda")]
                hs = (torch.empty(rows_, 512, device="cuda") if keep else None, torch.empty(rows_, 256, device="cuda") if keep else None,
                      torch.empty(rows_, 128, device="cuda"))
                pair.append((x, tuple(ps), hs))
            sets.append(pair)
        nbytes = 2 * (4.0 * rows_ * (obs_dim + 128 + ((512 + 256) if keep else 0)) + wbytes)
        out[f"K19 fused MLP x2 [{rows_}x{obs_dim}->512->256->128] {tag}"] = (time_graph(lambda i: ops.mlp3_forward(sets[i]), copies, reps), nbytes, 2.0 * 2 * rows_ * macs)
    for k, n in ((obs_dim, 512), (512, 256), (256, 128)):
        copies = 4
        xs = [[torch.randn(b, k, device="cuda") for _ in range(2)] for _ in range(copies)]
        gs = [[torch.randn(b, n, device="cuda") for _ in range(2)] for _ in range(copies)]
        dws, dbs = [torch.zeros(n, k, device="cuda") for _ in range(2)], [torch.zeros(n, device="cuda") for _ in range(2)]
        out[f"K15 pair wgrad+bias 2 x {n}x{k} over {b}"] = (
            time_graph(lambda i: ops.wgrad_pair(gs[i][0], xs[i][0], dws[0], dbs[0], gs[i][1], xs[i][1], dws[1], dbs[1]), copies, reps),
            2 * 4.0 * (b * (n + k) + n * k), 2 * 2.0 * b * k * n)
    return out


def bench_student_cnn(reps: int):
    out = {}
    w = tuple(torch.randn(*s, device="cuda") * 0.1 for s in ((24, 2, 4, 4), (24,), (24, 24, 3, 3), (24,), (24, 24, 2, 2), (24,), (64, 192), (64,)))
    for m in (405, 4096, 16384):
        x = [(torch.rand(m, 442, device="cuda") < 0.1).float() for _ in range(4)]
        o = torch.empty(m, 64, device="cuda")
        out[f"student pre-encoder K17 [{m} frames]"] = (time_graph(lambda i: ops.student_cnn_forward(w, image=x[i], out=o), 4, reps), (442 + 64) * 4.0 * m)
    return out


def bench_contact_sensor(n: int, reps: int, bodies: int = 17, H: int = 3):
    copies = 8
    f = [torch.randn(n, bodies, 3, device="cuda") for _ in range(copies)]
    hist = [torch.zeros(n, H, bodies, 3, device="cuda") for _ in range(copies)]
    tm = [[torch.zeros(n, bodies, device="cuda") for _ in range(4)] for _ in range(copies)]

    def run(i):
        ops.contact_sensor_update(f[i], history=hist[i], current_air_time=tm[i][0], last_air_time=tm[i][1], current_contact_time=tm[i][2],
                                  last_contact_time=tm[i][3], dt=0.02)

    return {f"contact sensor K18 [{n}x{bodies}]": (time_graph(run, copies, reps), (12.0 * (H + 1) + 16 + 12 * H + 16) * n * bodies)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--json", type=str, default=None)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    peak, kind = peak_gbs()
    n = args.envs
    res = {}
    only = set(args.only.split(",")) if args.only else None
    want = lambda k: only is None or k in only  # noqa: E731
    if want("mdp"):
        for task in ("locomotion", "teacher"):
            res.update(bench_mdp(task, n, args.reps))
    if want("taxel"):
        res.update(bench_taxel(n, args.reps))
    if want("gae"):
        res.update(bench_gae(n, args.reps))
    if want("ppo"):
        res.update(bench_ppo_loss(n * 24 // 4, args.reps))
    if want("adam"):
        res.update(bench_adam(607641 // 4 * 4, args.reps))
    if want("k9"):
        res.update(bench_k9(n * 24 // 4, args.reps))
    if want("splitpad"):
        res.update(bench_split_pad(n, args.reps))
    if want("gather"):
        res.update(bench_gather(n * 24, n * 24 // 4, max(20, args.reps // 10)))
    if want("command"):
        res.update(bench_command(n, args.reps))
    if want("heads"):
        res.update(bench_heads(n * 24 // 4, args.reps))
    if want("gemm"):
        res.update({k: v[:2] for k, v in bench_gemms(n * 24 // 4, max(20, args.reps // 5)).items()})
    if want("mlp3"):
        res.update({k: v[:2] for k, v in bench_mlp3(n, n * 24 // 4, max(20, args.reps // 5)).items()})
    if want("student"):
        res.update(bench_student_cnn(max(20, args.reps // 5)))
    if want("contact"):
        res.update(bench_contact_sensor(n, args.reps))
    if want("copy"):
        res.update(bench_copy([2.47 * n / 4096, 6.49 * n / 4096, 22.85 * n / 4096, 28.34 * n / 4096, 36.21 * n / 4096], args.reps))
    rows = []
    print(f"envs={n}  peak={peak} GB/s ({kind})")
    print(f"{'kernel':42s} {'us/launch':>10s} {'alg MB':>9s} {'GB/s':>9s} {'frac':>6s}")
    for k, (us, nbytes) in res.items():
        gbs = nbytes / us / 1e3
        rows.append(dict(kernel=k, us=us, bytes=nbytes, gbs=gbs, frac=gbs / peak))
        print(f"{k:42s} {us:10.2f} {nbytes / 1e6:9.2f} {gbs:9.1f} {gbs / peak:6.3f}")
    if args.json:
        json.dump(dict(envs=n, peak_gbs=peak, peak_kind=kind, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
