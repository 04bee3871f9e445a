"""Per-op CUDA-event breakdown of one eager step of a bench workload (stage2 | stage3 | vitl), like bench.py's op_breakdown_ms.
    python tools/stage_profile.py stage2 [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from unite_b200 import ops  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "stage2"
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(1000)
B = 8 if wl == "stage3" else 32
if wl == "stage2":
    from unite_b200.registry import create_model
    from unite_b200 import modeling_finetune  # noqa: F401
    from unite_b200.engine_for_finetuning import finetune_step
    from unite_b200.engine import FusedAdamW
    torch.manual_seed(0)
    model = create_model("vit_base_patch16_224", pretrained=False, num_classes=12, all_frames=8, tubelet_size=1, drop_path_rate=0.1,
                         use_mean_pooling=True, init_scale=0.001).to(dev).train()
    opt = FusedAdamW(model.core().arena, lr=1e-3)
    v, y = torch.randn(B, 3, 8, 224, 224, generator=g).to(dev), torch.randint(0, 12, (B,), generator=g).to(dev)
    loss = torch.zeros(1, device=dev)

    def step():
        opt.zero_grad(); loss.zero_()
        finetune_step(model, v, y, loss)
        opt.step()
elif wl == "stage3":
    from unite_b200.engine_stage3 import Stage3Engine
    student, teacher = bench.build_models(seed=0)
    eng = Stage3Engine(student.to(dev).train(), teacher.to(dev).eval(), torch.randn(12, 768, generator=g) * 0.5, torch.zeros(12),
                       torch.randn(12, 512, generator=g), mask_ratio=0.8, k=2)
    mk = lambda: torch.randn(B, 3, 8, 224, 224, generator=g)
    vt = mk()
    b = (mk().to(dev), torch.randint(0, 12, (B,), generator=g).to(dev), vt.to(dev), (vt + 0.1 * mk()).to(dev))

    def step():
        eng.step(*b)
else:
    from unite_b200.engine import Stage1Engine
    student, teacher = bench.build_models(seed=0, large=True)
    eng = Stage1Engine(student.to(dev).train(), teacher.to(dev).eval(), mask_ratio=0.8, use_graph=False)
    b = [(v.to(dev), q.to(dev)) for v, q in bench.host_batches(B, 0, frames=16, tokens_per_frame=2)][0]

    def step():
        eng.step(*b)
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record(); torch.cuda.synchronize()
eager_ms = e0.elapsed_time(e1) / 3
ops.PROFILE = []
torch.cuda._sleep(int(0.3 * 1.9e9))
step()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
by = {}
for name, info, a, b_ in prof:
    d = a.elapsed_time(b_)
    key = name if name != "gemm" else "gemm"
    by.setdefault(key, [0, 0.0]); by[key][0] += 1; by[key][1] += d
tot = sum(v[1] for v in by.values())
out = dict(workload=wl, eager_ms_per_step=round(eager_ms, 3), sum_of_op_ms=round(tot, 3), launches=sum(v[0] for v in by.values()),
           ops={k: dict(launches=v[0], ms=round(v[1], 3), share=round(v[1] / tot, 4)) for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])})
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
