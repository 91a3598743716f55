"""CUDA-event timing of the captured rollout / update graphs: python tools/graphs.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from locotouch_b200.engine import HotPathEngine  # noqa: E402


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


eng = HotPathEngine(num_envs=4096, task="teacher", tactile=True, device="cuda:0")
eng.capture()
print(f"[{os.environ.get('LT_TAG', '')}] rollout graph {timed(eng._graphs['roll'][0].replay):7.3f} ms   update graph {timed(eng._graphs['update'].replay):7.3f} ms")
