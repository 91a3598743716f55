"""Config C4 end to end (BASELINE.json configs[3]): one DAgger iteration of the distillation loop on the GPU -- data collection
through the device-side ReplayBuffer (teacher policy, K0/K1/K2 env stand-in, TactileRecorder, K11 bookkeeping), then
behaviour-cloning epochs of the CNN-RNN student on padded batches (K8 padding / masked loss, K7 AdamW).

    python tools/distill_bench.py [--envs 405] [--steps 20000] [--epochs 3]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200.distill import DistillationRandCylinderCNNRNNMonCfg, ReplayBuffer, Student, TactileRecorder  # noqa: E402
from locotouch_b200.sim.transport_env import SyntheticTransportEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=405)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--episode", type=int, default=100, help="time-out length of the synthetic episodes (env steps)")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    dev = torch.device("cuda:0")
    env = SyntheticTransportEnv(args.envs, dev, seed=0, max_episode_length=args.episode)
    torch.manual_seed(0)
    teacher = torch.nn.Sequential(torch.nn.Linear(348, 512), torch.nn.ELU(), torch.nn.Linear(512, 256), torch.nn.ELU(), torch.nn.Linear(256, 128),
                                  torch.nn.ELU(), torch.nn.Linear(128, 12)).to(dev)
    cfg = DistillationRandCylinderCNNRNNMonCfg(device=str(dev))
    cfg.batch_steps = 20000
    student = Student(cfg, 270, 442, 12, teacher_policy_inference=teacher)
    rb = ReplayBuffer(env, TactileRecorder(dev, args.envs, 442, 1, 2), 270)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rewards, lengths = rb.collect_data(teacher, None, args.steps)
    torch.cuda.synchronize()
    t_collect = time.perf_counter() - t0
    env_steps = env.t * args.envs
    print(f"collect_data: {rb.num_steps} recorded steps in {rb.num_trajs} trajectories ({len(lengths)} episodes, {env.t} env steps x {args.envs} envs) "
          f"in {t_collect * 1e3:.1f} ms -> {env_steps / t_collect:,.0f} env-steps/s, {rb.num_steps / t_collect:,.0f} recorded steps/s")
    student.train()
    batch_trajs = int(cfg.batch_steps / (rb.num_steps / rb.num_trajs)) + 1
    n_batches = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.epochs):
        for b in rb.to_recurrent_generator(batch_trajs):
            loss = student.train_on_batch(b)
            n_batches += 1
    torch.cuda.synchronize()
    t_train = time.perf_counter() - t0
    print(f"train: {args.epochs} epochs x {n_batches // args.epochs} batches of <= {batch_trajs} trajectories in {t_train * 1e3:.1f} ms -> "
          f"{args.epochs * rb.num_steps / t_train:,.0f} trained steps/s   last loss {float(loss[0] if loss.ndim else loss):.5f}")


if __name__ == "__main__":
    main()
