"""Run-to-run variation of one stage-3 step from identical state (tiny configuration): which tensor moves, and by how much.
    python tools/stage3_determinism.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests.util import build_student, build_teacher, load_golden, oracle_cfgs, seeded_states  # noqa: E402
from unite_b200.engine_stage3 import Stage3Engine  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
fix = load_golden("tiny_stage12.pt")
scfg, tcfg = oracle_cfgs(fix)
ssd, tsd, _ = seeded_states(fix)
C, D = 12, scfg.embed_dim
g = torch.Generator().manual_seed(31)
cls_w, cls_b, text = torch.randn(C, D, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1, torch.randn(C, tcfg.output_dim, generator=g)
shape = (3, scfg.num_frames, scfg.img_size, scfg.img_size)
vt = torch.randn(4, *shape, generator=g)
batch = (torch.randn(4, *shape, generator=g).cuda(), torch.randint(0, C, (4,), generator=g).cuda(), vt.cuda(),
         (vt + 0.1 * torch.randn(4, *shape, generator=g)).cuda())
student, teacher = build_student(scfg, drop_path_rate=0.2), build_teacher(tcfg)
student.load_state_dict(ssd, strict=True)
teacher.load_state_dict(tsd, strict=True)
eng = Stage3Engine(student.cuda().train(), teacher.cuda().eval(), cls_w, cls_b, text, mask_ratio=0.75, k=2, lr=1e-3)
p0 = eng.core.arena.params.clone()
ref = None
worst = {}
for r in range(reps):
    eng.core.arena.params.copy_(p0)
    eng.core.sync_shadow(force=True)
    eng.core.drop_path.step.zero_()
    eng.optimizer.zero_grad()
    eng.forward_backward(*batch)
    torch.cuda.synchronize()
    cur = {k: v.clone().float() for k, v in eng.last.items()}
    cur["grads"] = eng.core.arena.grads.clone()
    cur["loss3"] = torch.cat([eng.loss, eng.loss_s, eng.loss_t]).clone()
    if ref is None:
        ref = cur
        continue
    for k in cur:
        d = (cur[k] - ref[k]).abs().max().item() / max(ref[k].abs().max().item(), 1e-30)
        worst[k] = max(worst.get(k, 0.0), d)
    bad = {k: f"{(cur[k] - ref[k]).abs().max().item():.2e}" for k in cur if not torch.equal(cur[k], ref[k])}
    print(f"rep {r}: differs from rep 0 in {bad}", flush=True)
print("worst relative-to-max deviations:", {k: f"{v:.2e}" for k, v in worst.items()})
# where in the gradient arena
names = eng.core.arena.names
