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
        print(f"sanitize_cases: N = {n} done", flush=True)
    print("sanitize_cases: all kernels ran", flush=True)


if __name__ == "__main__":
    main()
