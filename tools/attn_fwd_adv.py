"""GPU check of the attention forward (LSE) kernel's slow path: the second key chunk holds far larger scores than the first,
so the reference maximum has to move and O / l in TMEM are rescaled.  python tools/attn_fwd_adv.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unite_b200 import ops

dev = "cuda"
ok = True
for (n_seq, S, H, mult) in [(3, 320, 4, 40.0), (2, 300, 3, 25.0), (2, 320, 2, 3.0)]:
    g = torch.Generator(device=dev).manual_seed(S + H)
    qkv = (torch.randn(n_seq * S, 3 * H * 64, device=dev, generator=g) * 0.7)
    v = qkv.view(n_seq, S, 3, H, 64)
    v[:, 160:, 1] *= mult                       # keys of the second chunk
    v[:, 170:200, 1] *= 0.01                    # ... but not all of them
    qkv = qkv.bfloat16()
    o = torch.empty(n_seq * S, H * 64, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(n_seq, H, S, device=dev)
    ops.attn_fwd(qkv, o, lse, n_seq, S, H, 0.125)
    q, k, vv = qkv.float().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref = s.softmax(-1) @ vv
    got = o.float().view(n_seq, S, H, 64).permute(0, 2, 1, 3)
    rel = ((got - ref).norm() / ref.norm()).item()
    lrel = ((lse - torch.logsumexp(s, -1)).abs().max() / torch.logsumexp(s, -1).abs().max()).item()
    good = rel < 8e-3 and lrel < 1e-3 and torch.isfinite(got).all().item()
    ok &= good
    print(f"adversarial S={S} H={H} x{mult}: o rel={rel:.3e} lse rel={lrel:.3e} max score {s.max().item():.0f} {'OK' if good else 'FAIL'}")
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
