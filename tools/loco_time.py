import torch, sys
sys.path.insert(0, ".")
from locotouch_b200.engine import HotPathEngine
eng = HotPathEngine(num_envs=4096, task="locomotion", tactile=False, device=torch.device("cuda:0"), seed=0, num_state_sets=6)
eng.capture()
for _ in range(3): eng.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.replay()
e1.record(); e1.synchronize()
print("locomotion iteration ms:", e0.elapsed_time(e1) / 10)
