"""Experiment: does capturing the whole stage-1 step in a CUDA graph remove measurable launch gaps?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_models
from unite_b200.engine import Stage1Engine

dev = torch.device("cuda")
student, teacher = build_models(0)
eng = Stage1Engine(student.to(dev).train(), teacher.to(dev).eval(), mask_ratio=0.8, lr=1e-4)
g = torch.Generator(device=dev).manual_seed(1)
v = torch.randn(32, 3, 8, 224, 224, device=dev, generator=g)
q = torch.empty(256, 196, device=dev).exponential_(1, generator=g)
for _ in range(3):
    eng.step(v, q)
torch.cuda.synchronize()

def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print("eager ms/step", timeit(lambda: eng.step(v, q)))
t0 = time.perf_counter()
for _ in range(5): eng.step(v, q)
t1 = time.perf_counter(); torch.cuda.synchronize()
print("host time to enqueue one step (ms)", (t1 - t0) / 5 * 1e3)
graph = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(graph):
        eng.step(v, q)
    torch.cuda.synchronize()
    print("graph ms/step", timeit(graph.replay))
except Exception as e:
    print("capture failed:", repr(e)[:500])
