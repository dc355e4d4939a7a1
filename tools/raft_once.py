"""ncu target: one RAFT forward (31 pairs of 224x224) without CUDA-graph replay."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotgb_b200 import synthetic as syn
from videotgb_b200.raft import RAFT

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 31
m = RAFT().cuda()
m.load_state_dict(syn.raft_state_dict())
m.eval()
m.graph_max_rows = 0
fr = torch.rand(pairs + 1, 3, 224, 224, device="cuda") * 255
for _ in range(2):
    m(fr[:-1], fr[1:])
torch.cuda.synchronize()
