"""CPU: pins the oracle (oracle/msda_ref.c, oracle/msda.py) against golden vectors generated
from the reference's own multi_scale_deformable_attn_pytorch (tests/golden/gen_msda_golden.py)."""
import glob
import os

import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import msda as O

CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "msda_*.npz")) if "module" not in p)


def test_golden_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("case", CASES)
def test_c_oracle_matches_reference_golden(case, built):
    g = load_golden(case)
    for acc_double in (True, False):
        out = O.msda_c(g["value"], g["shapes"], g["starts"], g["loc"], g["attn"], acc_double=acc_double)
        # the reference's grid_sample path and its CUDA kernel differ only by fp32 summation order
        torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("case", CASES)
def test_torch_oracle_matches_reference_golden(case):
    g = load_golden(case)
    out = O.msda_torch(g["value"], g["shapes"], g["loc"], g["attn"])
    torch.testing.assert_close(out, g["out"], rtol=0, atol=0)  # same ops, same order: bit-exact


def test_level_start_index():
    ss = torch.tensor([[4, 5], [2, 3], [1, 1]])
    assert O.level_start_index(ss).tolist() == [0, 20, 26]


def test_c_oracle_threads_agree(built):
    v, ss, st, loc, attn = O.make_inputs(2, 50, 8, 32, [(9, 11), (5, 6), (3, 3)], 4, seed=7, border=True)
    a = O.msda_c(v, ss, st, loc, attn, nthreads=1)
    b = O.msda_c(v, ss, st, loc, attn, nthreads=4)
    assert torch.equal(a, b)


def test_out_of_range_samples_contribute_zero(built):
    # every location outside (-1, size): the reference skips the sample (…cuh:285-291)
    v, ss, st, loc, attn = O.make_inputs(1, 4, 2, 8, [(3, 3)], 2, seed=1)
    loc = loc * 0 + 2.0
    assert O.msda_c(v, ss, st, loc, attn).abs().max() == 0
    assert O.msda_torch(v, ss, loc, attn).abs().max() == 0
