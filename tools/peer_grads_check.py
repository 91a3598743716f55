"""K14 check on N GPUs of one node: gradients summed by peer loads inside the optimizer kernel == NCCL all-reduce + clip + Adam.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/peer_grads_check.py

Every rank builds two identical PPO learners, fills both gradient buffers with the same rank-dependent values (and a rank-dependent
KL statistic in the tail), then runs learner A through NCCL (`allreduce_grads` + `step_after_reduce`) and learner B through K14
(barrier + `lt_peer_sum_clip_adam` + barrier) for a few steps.  Checks: parameters, Adam state and learning rate agree between A and
B (bit-exact for 2 ranks: a + b is commutative; 1e-6 beyond, where NCCL's association order differs), and B's parameters are
bit-identical on every rank.  Prints the per-step device time of both exchanges."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200.loco_rl import PPO, ActorCritic  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def learner():
    torch.manual_seed(0)
    ac = ActorCritic(348, 348, 12, [512, 256, 128], [512, 256, 128], "elu", 1.0)
    return PPO(ac, device=str(dev), num_learning_epochs=1, num_mini_batches=1, desired_kl=0.01, schedule="adaptive", learning_rate=1e-3, max_grad_norm=1.0)


os.environ["LT_PEER_GRADS"] = "0"
a = learner()
assert not a.enable_peer_gradients()
os.environ["LT_PEER_GRADS"] = "1"
b = learner()
ok = b.enable_peer_gradients()
if not ok:
    print(f"[rank {rank}] symmetric memory unavailable: {getattr(b, '_peer_error', '?')}", flush=True)
    dist.destroy_process_group()
    sys.exit(3)
os.environ["LT_PEER_TWO_SHOT"] = "1"  # third learner: the two-shot exchange (reduce-scatter by peer loads, barrier, gather) forced on
c = learner()
assert c.enable_peer_gradients() and c._peer["two_shot"]
os.environ["LT_PEER_TWO_SHOT"] = "auto"
n = a.actor_critic.flat_grads_ext.numel()
for step in range(4):
    g = torch.Generator(device="cpu").manual_seed(100 * step + rank)
    vals = (torch.randn(n, generator=g) * (0.05 if step else 5.0)).to(dev)  # step 0: the clip is active
    vals[-4] = 0.03 if step < 2 else 0.001  # KL statistic: first "too large" (lr / 1.5), then "too small" (lr * 1.5)
    vals[-3:] = 0
    for alg in (a, b, c):
        alg.select_gradient_buffer()  # K14 alternates between two symmetric buffers, one per mini-batch
        alg.actor_critic.flat_grads_ext.copy_(vals)
        alg.reduce_and_step()
    torch.cuda.synchronize()
    pa, pb = a.optimizer.flat, b.optimizer.flat
    tol = 0.0 if world == 2 else 1e-6
    for name, x, y in (("params", pa, pb), ("exp_avg", a.optimizer.exp_avg, b.optimizer.exp_avg), ("exp_avg_sq", a.optimizer.exp_avg_sq, b.optimizer.exp_avg_sq),
                       ("lr", a.optimizer.lr_t, b.optimizer.lr_t), ("grad_norm", a.optimizer.grad_norm, b.optimizer.grad_norm)):
        err = (x - y).abs().max().item()
        assert err <= tol * max(1.0, y.abs().max().item()), f"step {step} {name}: NCCL vs peer-sum differ by {err}"
    for name, x, y in (("params", c.optimizer.flat, pb), ("exp_avg_sq", c.optimizer.exp_avg_sq, b.optimizer.exp_avg_sq), ("lr", c.optimizer.lr_t, b.optimizer.lr_t),
                       ("grad_norm", c.optimizer.grad_norm, b.optimizer.grad_norm)):
        assert torch.equal(x, y), f"step {step} {name}: two-shot exchange differs from the one-shot exchange"
    lo, hi = pb.clone(), pb.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "replicas diverged"
    if rank == 0:
        print(f"step {step}: lr {float(b.optimizer.lr_t):.6g} grad_norm {float(b.optimizer.grad_norm):.6g} -- NCCL and peer-sum agree, replicas identical", flush=True)


def timed(alg, reps=50):
    step = lambda: (alg.select_gradient_buffer(), alg.reduce_and_step())  # noqa: E731
    for _ in range(5):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


# ---- the same exchange replayed from the captured update graph (barriers and peer loads are graph nodes): after a few iterations the
# replicas must still be bit-identical -- a barrier that let a rank read a buffer early or late would break that at once
from locotouch_b200.engine import HotPathEngine  # noqa: E402

eng = HotPathEngine(num_envs=512, task="teacher", tactile=True, device=dev, seed=3, num_state_sets=3, hidden=(128, 64))
assert eng.peer_gradients, "engine did not enable the peer exchange"
eng.capture()
assert not eng._graphs["split"] and not eng._graphs.get("finish_eager"), "two graphs per iteration expected (update and rollout incl. its tail)"
p0 = eng.alg.optimizer.flat.clone()
for _ in range(4):
    eng.replay()
torch.cuda.synchronize()
pe = eng.alg.optimizer.flat
assert not torch.equal(pe, p0), "parameters did not move"
assert bool(torch.isfinite(pe).all())
lo, hi = pe.clone(), pe.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert torch.equal(lo, hi), "replicas diverged under graph replay"
lr = eng.alg.optimizer.lr_t.clone()
lr_lo, lr_hi = lr.clone(), lr.clone()
dist.all_reduce(lr_lo, op=dist.ReduceOp.MIN)
dist.all_reduce(lr_hi, op=dist.ReduceOp.MAX)
assert torch.equal(lr_lo, lr_hi)
if rank == 0:
    print(f"graph replay: 4 iterations x 20 mini-batch exchanges, replicas bit-identical (lr {float(lr):.6g})", flush=True)

ta, tb, tc = timed(a), timed(b), timed(c)
if rank == 0:
    print(f"world={world}: NCCL all-reduce + clip + Adam {ta:.1f} us per step; barrier + peer-sum + clip + Adam {tb:.1f} us per step; "
          f"two-shot (barrier + reduce-scatter + barrier + gather + clip + Adam) {tc:.1f} us per step", flush=True)
dist.barrier()
dist.destroy_process_group()
