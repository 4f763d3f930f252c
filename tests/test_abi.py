"""CPU: the C-ABI library loads and exports exactly what include/ape_b200.h declares; the
operator names of the reference's native library are registered; no CPU fallback exists."""
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ape_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ape_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    import ape_b200

    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    nm = subprocess.run(["nm", "-D", "--defined-only", ape_b200._lib.LIB_PATH], capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in nm.stdout.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared but not exported: {missing}"
    assert sorted(ape_b200._lib.EXPORTS) == declared
    for s in declared:
        getattr(ape_b200._lib.lib, s)


def test_abi_version(built):
    import ape_b200

    assert ape_b200._lib.lib.ape_abi_version() == 3
    assert ape_b200._lib.lib.ape_last_error() == b""


def test_reference_operator_names_registered(built):
    import ape_b200  # noqa: F401

    assert torch.ops.ape.ms_deform_attn_forward is not None
    assert torch.ops.ape.ms_deform_attn_backward is not None
    schema = str(torch.ops.ape.ms_deform_attn_forward.default._schema)
    assert "Tensor value, Tensor spatial_shapes, Tensor level_start_index, Tensor sampling_loc, Tensor attn_weight, int im2col_step" in schema


def test_cpu_tensors_raise_like_reference(built):
    import ape_b200  # noqa: F401

    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):  # ms_deform_attn.h:39
        torch.ops.ape.ms_deform_attn_forward(
            torch.zeros(1, 4, 1, 8), torch.tensor([[2, 2]]), torch.tensor([0]),
            torch.zeros(1, 1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1), 64)


def test_invalid_arguments_are_reported_without_gpu(built):
    import ape_b200

    lib = ape_b200._lib.lib
    # argument validation happens before any CUDA call, so it is testable on a CPU-only box
    rc = lib.ape_msda_fwd(None, None, None, None, None, None, 1, 4, 1, 8, 1, 1, 1, 7, None)
    assert rc == -1 and b"dtype" in lib.ape_last_error()
    rc = lib.ape_msda_fwd(None, None, None, None, None, None, 1, 4, 1, 8, 1, 1, 1, 0, None)
    assert rc == -3
    rc = lib.ape_msda_fwd(None, None, None, None, None, None, 0, 4, 1, 8, 1, 1, 1, 0, None)
    assert rc == 0  # empty batch: nothing to do, like a zero-size launch in the reference


def test_product_does_not_import_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ape_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "oracle/" in src:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product code references oracle/: {bad}"
