#!/usr/bin/env python
"""Where the time of DefaultPredictor.__call__ goes (development aid): host timers with a synchronise after every stage."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ape_b200.engine import DefaultPredictor, ResizeShortestEdge  # noqa: E402
from ape_b200.modeling import build_model  # noqa: E402

dev = torch.device("cuda:0")
model = bench.bench_weights(build_model(bench.bench_spec(), num_text=1203)).to(dev)
model.engine_dtype = torch.float16
model.use_cuda_graphs = True
pred = DefaultPredictor(model, ResizeShortestEdge(1024, 1024), "RGB")
u8 = [np.random.default_rng(i).integers(0, 256, (1024, 1024, 3), dtype=np.uint8) for i in range(4)]
host = [torch.randint(0, 256, (3, 1024, 1024)).float().pin_memory() for _ in range(4)]
for i in range(4):
    pred(u8[i]); model([{"image": host[i], "height": 1024, "width": 1024}])
torch.cuda.synchronize()
def T(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("model(host float)      ", round(T(lambda i: model([{"image": host[i % 4], "height": 1024, "width": 1024}])), 2), "ms")
print("predictor(u8)          ", round(T(lambda i: pred(u8[i % 4])), 2), "ms")
print("  upload only          ", round(T(lambda i: pred._upload(u8[i % 4], 0)), 2), "ms")
print("  upload only (again)  ", round(T(lambda i: pred._upload(u8[i % 4], 0)), 2), "ms")
print("predictor(u8) again    ", round(T(lambda i: pred(u8[i % 4])), 2), "ms")
torch.set_num_threads(1)
print("predictor(u8), 1 thread", round(T(lambda i: pred(u8[i % 4])), 2), "ms")
print("model(host), 1 thread  ", round(T(lambda i: model([{"image": host[i % 4], "height": 1024, "width": 1024}])), 2), "ms")
print("  preprocess only      ", round(T(lambda i: pred.preprocess(u8[i % 4])), 2), "ms")
inp = pred.preprocess(u8[0])
print("  model(device float)  ", round(T(lambda i: model([inp])), 2), "ms")
pred.predict_batch(u8)  # the packed-forward graph is captured on the first call
print("  predict_batch(16)/img", round(T(lambda i: pred.predict_batch(u8 * 4), 3) / 16, 2), "ms")
