"""Where the stage-1 harness step spends its GPU time (torch.profiler, kernel table)."""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tvq_b200 as tvq

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cl = len(sys.argv) > 2 and sys.argv[2] == "cl"
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = True
torch.set_float32_matmul_precision("high")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0); np.random.seed(0)
model = tvq.Stage1(200, 4, tvq.stage1.default_config()).to(dev)
if cl:
    model = model.to(memory_format=torch.channels_last)
tr = tvq.Stage1Trainer(model, (B, 4, 200), use_graph=False)
x = torch.rand(B, 4, 200, device=dev) * 2 - 1
for _ in range(5):
    tr.step(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
