"""Small invocations of every hand-written kernel for `compute-sanitizer` ("Sanitizers" row).

    compute-sanitizer --tool memcheck  --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r2_memcheck.log 2>&1
    compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py --race > gpurun_out/r2_racecheck.log 2>&1

ONE tool per gpurun call (B200_PROFILING.md).  Sizes N in {1, 405, 4096, 4097}: a single env, the distillation env count, the
BASELINE size and one past it (ragged last block).  Results are not compared here (the -m gpu tests do that); the point is that
the sanitizer sees every kernel at every size and reports nothing.  --race keeps the cases short (racecheck is ~100x slower)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from locotouch_b200 import ops  # noqa: E402
from locotouch_b200.mdp import task_spec as TS  # noqa: E402
from locotouch_b200.mdp.fused import FusedMdp  # noqa: E402
from locotouch_b200.sim import synth  # noqa: E402


def case_mdp(n, steps):
    for task in ("teacher", "locomotion"):
        spec = TS.SPECS[task]()
        env = synth.make_env(n, seed=3, with_object=spec.with_object).to("cuda")
        m = FusedMdp(env, spec, seed=1)
        for _ in range(steps):
            m.step(True, True)
        m.compute_rewards()
        m.compute_observations()
    torch.cuda.synchronize()


def case_taxel(n):
    g = torch.Generator().manual_seed(n)
    q = torch.randn(n, 238, 4, generator=g)
    q = (q / q.norm(dim=-1, keepdim=True)).cuda()
    f = (torch.randn(n, 221, 3, generator=g) * 0.1).cuda()
    thr = (0.05 + (torch.rand(n, 221, generator=g) - 0.5) * 0.02).cuda()
    ring = torch.zeros(n, 2, 7, device="cuda", dtype=torch.int32)
    first = torch.ones(n, device="cuda", dtype=torch.uint8)
    delay = torch.ones(n, device="cuda", dtype=torch.int64)
    for k in range(2):
        ops.taxel_synth(q, f, thr, quat_body_offset=17, seed=1, offset=k, delay_ring=ring, delay_first=first, delay_steps=delay,
                        delayed_signal=torch.empty(n, 442, device="cuda"))
    out = torch.empty(n, 4, 221, device="cuda")
    ops.taxel_forces(q, f, thr, out, ("contact", "normalized", "minmax", "discretized"), quat_body_offset=17, p_drop=0.005, p_add=0.005,
                     add_force_noise=True, force_n_prop_min=-0.1, force_n_prop_max=0.1, maximal_force=3.0, total_levels=5, add_level_noise=True,
                     level_n_min=-1, level_n_max=1, seed=1, offset=0)
    torch.cuda.synchronize()


def case_gae(n):
    for T in (24, 7):  # one-launch kernel (T = 24) and the two-launch path
        g = torch.Generator().manual_seed(n + T)
        r, v = torch.randn(T, n, generator=g).cuda(), torch.randn(T, n, generator=g).cuda()
        d, lv = (torch.rand(T, n, generator=g) < 0.05).byte().cuda(), torch.randn(n, generator=g).cuda()
        if T * n > 1:
            for _ in range(2):  # twice: the barrier counters must re-arm
                ops.gae(r, v, d, lv, 0.99, 0.95, True)
        ops.gae(r, v, d, lv, 0.99, 0.95, False)
    torch.cuda.synchronize()


def case_update(b):
    """K6 + K9 + K7 (+ K14 over two local buffers) at mini-batch size b."""
    A = 12
    g = torch.Generator().manual_seed(b)
    rn = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
    lr = torch.tensor([1e-3], device="cuda")
    for _ in range(2):
        ops.ppo_loss(rn(b, A), (0.5 + torch.rand(A, generator=g)).cuda(), rn(b), rn(b, A), rn(b), rn(b, A), (0.5 + torch.rand(b, A, generator=g)).cuda(), rn(b),
                     rn(b), rn(b), entropy_coef=0.01, desired_kl=0.01, lr=lr)
    for n in (128, 12, 1):
        gr, h, bg = rn(b, n), rn(b, n), torch.empty(n, device="cuda")
        ops.bias_act_bwd(gr, h, bg)
        ops.bias_act_bwd(gr, None, bg)
    for n in (607644, 1000, 4):
        p, gr, m, v = rn(n), rn(n), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
        step = torch.zeros(1, device="cuda")
        for _ in range(2):
            ops.clip_adam(p, gr, m, v, lr, step)
        bufs = [torch.randn(n + 4, device="cuda") for _ in range(2)]
        gsum = torch.zeros(n + 4, device="cuda")
        for _ in range(2):
            ops.peer_sum_clip_adam(p, [x.data_ptr() for x in bufs], gsum, 4, m, v, lr, step, desired_kl=0.01, kl_scale=0.5, grad_scale=0.5)
    torch.cuda.synchronize()


def case_rollout(n):
    A, D = 12, 348
    mu, sigma = torch.randn(n, A, device="cuda"), torch.ones(A, device="cuda")
    ops.act_sample(mu, sigma, seed=1, offset=0)
    raw, prev, pp, proc = (torch.zeros(n, A, device="cuda") for _ in range(4))
    ops.process_actions(torch.randn(n, A, device="cuda"), raw, prev, pp, proc, clip=100.0, raw_scale=0.25, scale=1.0, offset=torch.zeros(n, A, device="cuda"))
    srcs = [torch.randn(n * 4, d, device="cuda") for d in (D, D, A, 1, 1, 1, 1, A, A)]
    ops.gather_rows(srcs, torch.randperm(n * 4, device="cuda")[: max(1, n)].contiguous())
    torch.cuda.synchronize()


def case_round2(n):
    """The kernels added in round 2: K15 (tcgen05 split-K weight + bias gradient; mbarrier pipeline, TMEM, bulk tensor reductions),
    K16 (heads + loss + head dgrad), K12 (fused linear / dgrad), K17 (student pre-encoder), K18 (contact sensor), K10 (trajectories)."""
    g = torch.Generator().manual_seed(n)
    rn = lambda *s: torch.randn(*s, generator=g).cuda()  # noqa: E731
    b = max(8, n)
    for (no, k) in ((128, 256), (512, 348), (12, 128), (1, 128), (20, 12)):
        ops.wgrad(rn(b, no), rn(b, k), torch.zeros(no, k, device="cuda"), torch.zeros(no, device="cuda"))
    if torch.backends.cuda.matmul.allow_tf32 or True:
        ops.linear_bias_act(rn(b, 348), rn(512, 348), rn(512), elu=True)
        ops.dgrad_act_bwd(rn(b, 256), rn(256, 512), rn(b, 512))
    A, H = 12, 128
    bufs = ops.PpoLossBuffers(b, A, "cuda")
    lr = torch.tensor([1e-3], device="cuda")
    for _ in range(2):
        ops.ppo_heads_loss(rn(b, H), rn(b, H), rn(A, H), rn(A), rn(1, H), rn(1), (0.5 + torch.rand(A, generator=g)).cuda(), rn(b, A), rn(b), rn(b, A),
                           (0.5 + torch.rand(b, A, generator=g)).cuda(), rn(b), rn(b), rn(b), torch.empty(b, H, device="cuda"), torch.empty(b, H, device="cuda"),
                           entropy_coef=0.01, desired_kl=0.01, lr=lr, buffers=bufs)
    w = (rn(24, 2, 4, 4), rn(24), rn(24, 24, 3, 3), rn(24), rn(24, 24, 2, 2), rn(24), rn(64, 192), rn(64))
    ops.student_cnn_forward(w, image=(torch.rand(n, 442, generator=g) < 0.1).float().cuda())
    ops.student_cnn_forward(w, packed=torch.randint(0, 2 ** 31 - 1, (n, 7), generator=g, dtype=torch.int32).cuda())
    hist, tm = torch.zeros(n, 3, 17, 3, device="cuda"), [torch.zeros(n, 17, device="cuda") for _ in range(4)]
    for step in range(2):
        ops.contact_sensor_update(rn(n, 17, 3), net_forces_w=torch.empty(n, 17, 3, device="cuda"), history=hist, current_air_time=tm[0], last_air_time=tm[1],
                                  current_contact_time=tm[2], last_contact_time=tm[3], dt=0.02,
                                  reset_mask=(torch.rand(n, generator=g) < 0.1).byte().cuda() if step else None)
    idx = ops.TrajectoryIndex((torch.rand(24, n, 1, generator=g) < 0.1).cuda())
    padded, masks = idx.split_and_pad(rn(24, n, 30))
    idx.unpad(padded)
    torch.cuda.synchronize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--race", action="store_true")
    args = ap.parse_args()
    sizes = (1, 405, 4096, 4097) if not args.race else (1, 405, 4097)
    for n in sizes:
        case_mdp(n, 1 if args.race else 2)
        case_taxel(n)
        case_gae(n)
        case_rollout(n)
        case_update(max(4, n if args.race else n * 6))
        case_round2(n)
        print(f"sanitize_cases: N = {n} done", flush=True)
    print("sanitize_cases: all kernels ran", flush=True)


if __name__ == "__main__":
    main()
