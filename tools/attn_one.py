import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unite_b200 import ops
n_seq, S, H = 256, 197, 12
qkv = (torch.randn(n_seq * S, 3 * H * 64, device="cuda") * 0.7).bfloat16()
o = torch.empty(n_seq * S, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_fwd(qkv, o, None, n_seq, S, H, 0.125)
torch.cuda.synchronize()
print("ok", o.float().abs().mean().item())
