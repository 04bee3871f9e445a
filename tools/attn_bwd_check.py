"""GPU check of the attention kernels on one shape: python tools/attn_bwd_check.py n_seq S H [time]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import kernels_check as kc

n_seq, S, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kc.check_attention(n_seq, S, H, time_it=len(sys.argv) > 4)
torch.cuda.synchronize()
print("ALL OK" if kc.OK else "FAILED", flush=True)
sys.exit(0 if kc.OK else 1)
