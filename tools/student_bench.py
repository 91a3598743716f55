"""Config C4: one behaviour-cloning step of the CNN-RNN student on a padded batch [L, B, ...] -- GPU (drop-in
Student: cuDNN conv / GRU, TF32 GEMMs, fused masked loss K8 + fused AdamW K7) next to the same architecture and loss in plain
torch on the host cores (what the reference's Student.train_on_data does per batch, reference student.py:121-151).

    python tools/student_bench.py [--L 500] [--B 41] [--cpu-reps 1]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200.distill import DistillationRandCylinderCNNRNNMonCfg, Student  # noqa: E402


def make_batch(L, B, device, seed=0):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(min(100, L), L + 1, (B,), generator=g)
    lengths[0] = L
    masks = torch.arange(L).unsqueeze(1) < lengths.unsqueeze(0)
    tac = (torch.rand(L, B, 442, generator=g) < 0.1).float()
    batch = dict(proprioceptions=torch.randn(L, B, 270, generator=g), teacher_encoder_obses=torch.randn(L, B, 78, generator=g),
                 tactile_signals=tac, masks=masks)
    return {k: v.to(device) for k, v in batch.items()}, int(masks.sum())


def build(device, seed=0):
    torch.manual_seed(seed)
    cfg = DistillationRandCylinderCNNRNNMonCfg(device=str(device))
    w = (torch.randn(12, 348) * 0.05).to(device)
    return Student(cfg, 270, 442, 12, teacher_policy_inference=lambda x: torch.nn.functional.linear(x, w))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=int, default=500)
    ap.add_argument("--B", type=int, default=41)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu-reps", type=int, default=1)
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    dev = torch.device("cuda:0")
    student = build(dev)
    student.train()
    batch, steps = make_batch(args.L, args.B, dev)
    for _ in range(3):
        student.train_on_batch(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        loss = student.train_on_batch(batch)
    e1.record()
    e1.synchronize()
    gpu_ms = e0.elapsed_time(e1) / args.reps
    print(f"GPU  student step [L={args.L}, B={args.B}] ({steps} valid steps): {gpu_ms:8.2f} ms  -> {steps / gpu_ms * 1e3:,.0f} steps/s   loss {float(loss[0] if loss.ndim else loss):.5f}")
    # ---- does cuDNN take the GRU weights where they lie (inside the flat AdamW buffer) or re-compact them per call?
    import warnings

    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        student.student_encoder.memory.rnn(torch.randn(4, 3, student.student_encoder.memory.rnn.input_size, device=dev))
        torch.cuda.synchronize()
    compact = [str(w.message)[:60] for w in rec if "contiguous chunk" in str(w.message)]
    print(f"GRU weights inside the flat buffer: {'re-compacted per call (' + compact[0] + '...)' if compact else 'used in place (no cuDNN weight copy)'}")
    # ---- inference: the student acting for N envs (one env step of the DAgger collection): K17 + GRU cell + MLPs vs the torch modules
    student.eval()
    for n_envs in (405, 4096, 16384):
        obs = {"policy": torch.randn(n_envs, 270, device=dev), "tactile": (torch.rand(n_envs, 442, device=dev) < 0.1).float()}
        res = {}
        for name, ctx in (("fused pre-encoder (K17)", torch.no_grad), ("torch modules (cuDNN)", torch.enable_grad)):
            student.reset()
            with ctx():
                for _ in range(3):
                    student.extract_input_and_forward(obs)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(20):
                    out = student.extract_input_and_forward(obs)
                e1.record()
                e1.synchronize()
            res[name] = e0.elapsed_time(e1) / 20 * 1e3
            del out
        a, b = res["fused pre-encoder (K17)"], res["torch modules (cuDNN)"]
        print(f"student act, {n_envs:6d} envs: {a:8.1f} us with K17   {b:8.1f} us through the torch modules   ({n_envs / a:,.1f} M env-steps/s)")
        with torch.no_grad():
            e0.record()
            for _ in range(20):
                student.pre_encoder(obs["tactile"].reshape(-1, 2, 17, 13))
            e1.record()
            e1.synchronize()
        k17 = e0.elapsed_time(e1) / 20 * 1e3
        with torch.enable_grad():
            e0.record()
            for _ in range(20):
                student.pre_encoder(obs["tactile"].reshape(-1, 2, 17, 13))
            e1.record()
            e1.synchronize()
        ref = e0.elapsed_time(e1) / 20 * 1e3
        print(f"   pre-encoder alone: K17 {k17:7.1f} us   torch {ref:7.1f} us   ({2 * 216.0e3 * n_envs / k17 / 1e6:.2f} TFLOP/s fp32 FMA)")
    student.train()
    # ---- the same architecture / loss in plain torch on the host cores
    torch.set_num_threads(os.cpu_count() or 1)
    cpu = build(torch.device("cpu"))
    cpu.train()
    cb, _ = make_batch(args.L, args.B, "cpu")
    opt = torch.optim.AdamW(cpu.parameters(), lr=5e-4)
    crit = torch.nn.MSELoss(reduction="none")

    def cpu_step():
        opt.zero_grad()
        a = cpu(cb["proprioceptions"], cb["tactile_signals"])
        with torch.no_grad():
            t = cpu.teacher_policy_inference(torch.cat((cb["proprioceptions"], cb["teacher_encoder_obses"]), dim=-1))
        loss = (crit(a, t).mean(dim=-1) * cb["masks"]).sum() / cb["masks"].sum()
        loss.backward()
        opt.step()

    cpu_step()
    t0 = time.perf_counter()
    for _ in range(args.cpu_reps):
        cpu_step()
    cpu_ms = (time.perf_counter() - t0) / args.cpu_reps * 1e3
    print(f"CPU  same step, torch fp32, {torch.get_num_threads()} threads: {cpu_ms:8.1f} ms  -> {steps / cpu_ms * 1e3:,.0f} steps/s   (GPU/CPU = {cpu_ms / gpu_ms:.0f}x)")


if __name__ == "__main__":
    main()
