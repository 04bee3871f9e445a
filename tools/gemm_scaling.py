"""Fixed cost vs per-tile cost of ub_gemm_bf16: time(M) at fixed N, K (python tools/gemm_scaling.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemm_check as g
for (N, K) in ((2304, 768), (768, 3072)):
    for M in (256, 2368, 4736, 9472, 18944, 37888):     # 1, 9.25, 18.5, 37, 74, 148 m-tiles of 256
        g.run(M, N, K, bias=True, time_it=True)
