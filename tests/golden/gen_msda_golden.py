"""Generates tests/golden/msda_*.npz from the REFERENCE's own implementation.

Run in the build container only (needs /root/reference):  python tests/golden/gen_msda_golden.py
It imports ape/layers/multi_scale_deform_attn.py *unmodified* by path (with an empty `ape._C`
stand-in so the module keeps its real class, multi_scale_deform_attn.py:415-423) and records
`multi_scale_deformable_attn_pytorch` (:84-124) on seeded inputs, plus the full
`MultiScaleDeformableAttention.forward` (:215-358, pytorch_attn=True) for the fused path."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_module():
    pkg = types.ModuleType("ape")
    pkg.__path__ = [os.path.join(REF, "ape")]
    sys.modules.setdefault("ape", pkg)
    sys.modules.setdefault("ape._C", types.ModuleType("ape._C"))
    pkg._C = sys.modules["ape._C"]
    spec = importlib.util.spec_from_file_location(
        "ape_ref_msda", os.path.join(REF, "ape/layers/multi_scale_deform_attn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


CASES = {
    # name: (B, Q, H, D, shapes, P, border)
    "tiny": (1, 3, 2, 8, [(3, 4)], 2, False),
    "ragged": (2, 17, 8, 32, [(6, 10), (3, 5), (2, 3), (1, 2), (1, 1)], 4, False),
    "border": (2, 29, 8, 32, [(8, 8), (4, 4), (2, 2), (1, 1)], 4, True),
    "d16p8": (1, 11, 4, 16, [(9, 7), (5, 4)], 8, True),
    "onepix": (1, 5, 8, 32, [(1, 1), (1, 3)], 4, True),
}


def main():
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.msda import make_inputs

    mod = load_reference_module()
    torch.manual_seed(0)
    for name, (B, Q, H, D, shapes, P, border) in CASES.items():
        value, ss, st, loc, attn = make_inputs(B, Q, H, D, shapes, P, seed=3, border=border)
        out = mod.multi_scale_deformable_attn_pytorch(value, ss, loc, attn)
        np.savez_compressed(os.path.join(HERE, f"msda_{name}.npz"), value=value.numpy(), shapes=ss.numpy(),
                            starts=st.numpy(), loc=loc.numpy(), attn=attn.numpy(), out=out.numpy())
        print(name, tuple(out.shape), float(out.abs().mean()))

    # full module forward (reference points in both 2-d and 4-d form) for the fused entry point
    for tag, ref_dim in (("ref2", 2), ("ref4", 4)):
        shapes = [(10, 14), (5, 7), (3, 4)]
        L = len(shapes)
        torch.manual_seed(1)
        m = mod.MultiScaleDeformableAttention(embed_dim=64, num_heads=4, num_levels=L, num_points=4,
                                              dropout=0.0, batch_first=True, pytorch_attn=True).eval()
        with torch.no_grad():
            # zero-initialised in the reference (:194,208-209); redraw so the test is not degenerate
            m.sampling_offsets.weight.normal_(0, 0.02)
            m.attention_weights.weight.normal_(0, 0.5)
            m.attention_weights.bias.normal_(0, 0.5)
        ss = torch.tensor(shapes)
        S = int((ss[:, 0] * ss[:, 1]).sum())
        st = torch.cat([ss.new_zeros(1), (ss[:, 0] * ss[:, 1]).cumsum(0)[:-1]])
        g = torch.Generator().manual_seed(5)
        B, Q = 2, 23
        query = torch.randn(B, Q, 64, generator=g)
        value = torch.randn(B, S, 64, generator=g)
        qpos = torch.randn(B, Q, 64, generator=g)
        ref = torch.rand(B, Q, L, ref_dim, generator=g)
        if ref_dim == 4:
            ref[..., 2:] = ref[..., 2:] * 0.5 + 0.05
        mask = torch.zeros(B, S, dtype=torch.bool)
        mask[1, -7:] = True
        with torch.no_grad():
            out = m(query, value=value, identity=query, query_pos=qpos, key_padding_mask=mask,
                    reference_points=ref, spatial_shapes=ss, level_start_index=st)
        sd = {k: v.numpy() for k, v in m.state_dict().items()}
        np.savez_compressed(os.path.join(HERE, f"msda_module_{tag}.npz"), query=query.numpy(), value=value.numpy(),
                            query_pos=qpos.numpy(), ref=ref.numpy(), mask=mask.numpy(), shapes=ss.numpy(),
                            starts=st.numpy(), out=out.numpy(), **{"sd." + k: v for k, v in sd.items()})
        print("module", tag, tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
