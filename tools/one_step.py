"""One eager stage-1 step at the benchmarked shapes (B=32) between cudaProfilerStart/Stop, for ncu captures:

    python tools/one_step.py                          # plain run (must exit 0 before the ncu run)
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/step_dram.csv python tools/one_step.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from unite_b200.engine import Stage1Engine  # noqa: E402

B = int(os.environ.get("UB_ONE_STEP_B", "32"))
dev = torch.device("cuda", 0)
student, teacher = bench.build_models(seed=0)
eng = Stage1Engine(student.to(dev).train(), teacher.to(dev).eval(), mask_ratio=0.8, lr=1.5e-4 * B / 256, use_graph=False)
batches = [(v.to(dev), q.to(dev)) for v, q in bench.host_batches(B, 0)]
for i in range(3):
    eng.step(*batches[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = eng.step(*batches[1])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", loss.item())
