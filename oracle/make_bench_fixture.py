"""Writes tests/golden/bench_b32_check.json — the scalar fixture bench.py checks its first eager step against at the
BENCHMARKED shapes (per-GPU batch 32, BASELINE.json configs[1]) before it prints anything — and tests/golden/vitb16_shapes.json.

    python oracle/make_bench_fixture.py        (build container, ~2 min of CPU; needs no GPU and no /root/reference)

The fixture is the fp32 oracle (oracle/unite_oracle.stage1_step, pinned to the reference modules by oracle/make_golden.py) run on
exactly what `bench.py --gpus 1` feeds its first resident batch: the seed-0 initial weights of bench.build_models (the reference's
initialisers) and host_batches(32, rank 0)[0], DropPath off.  Stored: the loss, the packed visibility mask, the visible-token
count and a digest of the weights (bench.weights_digest) so that a changed initialisation is reported as a stale fixture, not as
a kernel bug.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import unite_oracle as O  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    student, teacher = bench.build_models(seed=0)
    ssd = {k: v.detach().clone() for k, v in student.state_dict().items()}
    tsd = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    videos, q = bench.host_batches(32, 0)[0]
    ref = O.stage1_step(ssd, tsd, videos, q, O.StudentCfg(), O.TeacherCfg(), mask_ratio=0.8, with_grads=False)
    mask = ref["mask"].numpy().astype(np.uint8).reshape(-1)
    out = dict(loss=float(ref["loss"]), visible_tokens=int((~ref["mask"]).sum()), mask_hex=np.packbits(mask).tobytes().hex(),
               weights_sha16=bench.weights_digest(ssd),
               inputs_sha16=bench.weights_digest({"videos": videos[:1, :, :1, :64, :64].contiguous(), "q": q[:8]}), batch=32, seed_weights=0, seed_inputs=1000,
               made_by="oracle/make_bench_fixture.py", torch=torch.__version__)
    gold = os.path.join(ROOT, "tests", "golden")
    json.dump(out, open(os.path.join(gold, "bench_b32_check.json"), "w"))
    json.dump(dict(student={k: list(v.shape) for k, v in ssd.items()}, teacher={k: list(v.shape) for k, v in tsd.items()}),
              open(os.path.join(gold, "vitb16_shapes.json"), "w"))
    print(f"loss {out['loss']:.6f}, visible {out['visible_tokens']}, weights {out['weights_sha16']}")


if __name__ == "__main__":
    main()
