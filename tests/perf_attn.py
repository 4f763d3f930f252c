#!/usr/bin/env python
"""ViT attention shapes: ape_attn_fwd vs library SDPA (development aid; CUDA events, 20 iterations)."""
import os, sys, json
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ape_b200
from ape_b200 import ops
DEV = "cuda:0"
for (nb, n) in ((4, 1024), (1, 4096)):
    heads, hd = 16, 64
    qkv = torch.randn(nb * n, 3 * heads * hd, device=DEV, dtype=torch.float16)
    q5 = qkv.view(nb, n, 3, heads, hd)
    def own(variant):
        def f():
            ape_b200._lib.lib.ape_attn_variant(variant)
            return ops.attention_qkv(qkv, nb, n, heads, hd, 0.125)
        return f

    fns = {"own_smemP": own(0), "own_tmemP": own(1),
           "sdpa": lambda: F.scaled_dot_product_attention(q5[:, :, 0].transpose(1, 2), q5[:, :, 1].transpose(1, 2), q5[:, :, 2].transpose(1, 2), scale=0.125)}
    for name, fn in fns.items():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): fn()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        fl = 4 * nb * heads * n * n * hd
        print(json.dumps({"shape": [nb, n], "impl": name, "us": round(us, 1), "tflops": round(fl / us / 1e6, 1)}), flush=True)
