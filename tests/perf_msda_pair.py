"""GPU sweep of the pair-layout MSDA kernel's CTA mappings at the APE-L_D encoder shapes (development aid):

    python tests/perf_msda_pair.py > gpurun_out/msda_pair_sweep.txt

Encoder-like inputs (reference points = pixel centres, offsets = the sampling_offsets bias grid + noise, as the model
produces them), fp16, L2 flushed before every timed launch, median of 9."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ape_b200  # noqa: E402,F401
from ape_b200 import ops  # noqa: E402

DEV = "cuda:0"
H, D, P = 8, 32, 4


def case(shapes, dtype, noise=1.0, seed=0):
    L = len(shapes)
    g = torch.Generator().manual_seed(seed)
    ss = torch.tensor(shapes)
    areas = ss[:, 0] * ss[:, 1]
    st = torch.cat([areas.new_zeros(1), areas.cumsum(0)[:-1]])
    S = int(areas.sum())
    value = torch.randn(1, S, H * D, generator=g).to(DEV, dtype)
    th = torch.arange(H, dtype=torch.float32) * (2 * math.pi / H)
    gr = torch.stack([th.cos(), th.sin()], -1)
    gr = gr / gr.abs().max(-1, keepdim=True)[0]
    off = gr[:, None, None, :] * torch.arange(1, P + 1, dtype=torch.float32)[None, None, :, None]  # [H,1,P,2]
    offs = off.expand(H, L, P, 2)[None, None] + noise * torch.randn(1, S, H, L, P, 2, generator=g)
    qo = torch.cat([offs.reshape(1, S, -1), torch.randn(1, S, H * L * P, generator=g)], -1).to(DEV, dtype)
    pts = []
    for (h, w) in shapes:
        ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
        pts.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    ref = torch.cat(pts, 0)[None, :, None, :].expand(1, S, L, 2).contiguous().to(DEV)
    return value, ss.to(DEV), st.to(DEV), qo, ref, S


def med(fn, flush, iters=9):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for name, shapes in (("1024", [(256, 256), (128, 128), (64, 64), (32, 32), (16, 16)]),
                         ("1536", [(384, 384), (192, 192), (96, 96), (48, 48), (24, 24)])):
        for dtype in (torch.float16, torch.bfloat16):
            value, ss, st, qo, ref, S = case(shapes, dtype)
            n_off = H * len(shapes) * P * 2
            v4 = value.view(1, S, H, D)
            t = med(lambda: ops.ms_deform_attn_fused_forward(v4, ss, st, qo[..., :n_off], qo[..., n_off:], ref, P), flush)
            print(f"{name} {str(dtype)[6:]:9s} generic fused kernel (fp32 blend)        {t:8.1f} us")
            t = med(lambda: ops.msda_pair_values(value, H), flush)
            print(f"{name} {str(dtype)[6:]:9s} pairing pass                              {t:8.1f} us")
            v2 = ops.msda_pair_values(value, H)
            base = ops.ms_deform_attn_pair_fused_forward(v2, ss, st, shapes, qo[..., :n_off], qo[..., n_off:], ref, P, tile_w=0)
            for hpc in (1, 2, 8):
                for tw in (0, 4, 8, 16, 32):
                    if tw > 32 // hpc:
                        continue
                    for hm in (0, 1):
                        fn = lambda: ops.ms_deform_attn_pair_fused_forward(v2, ss, st, shapes, qo[..., :n_off], qo[..., n_off:], ref, P,
                                                                           heads_per_cta=hpc, tile_w=tw, head_major=hm)
                        out = fn()
                        err = (out.float() - base.float()).abs().max().item()
                        print(f"{name} {str(dtype)[6:]:9s} pair heads/cta {hpc} tile_w {tw:2d} head_major {hm}   {med(fn, flush):8.1f} us   (max diff vs tile_w=0: {err:.1e})")
            if name == "1536" or dtype == torch.bfloat16:
                continue


if __name__ == "__main__":
    main()
