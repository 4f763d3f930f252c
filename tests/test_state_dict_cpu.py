"""CPU: the engine's module tree exposes exactly the reference's `state_dict` (names and shapes), so
`DetectionCheckpointer.load` (ape/engine/defaults.py:193-194) fills it by name — checked against the reference model
built from its own files under the import shims (build container only)."""
import pytest

from ape_b200 import configs


@pytest.mark.parametrize("spec_name", ["MINI", "APE_TI", "APE_L_D"])
def test_state_dict_keys_and_shapes_equal_reference(spec_name):
    from oracle import ref_model, refshim

    if not refshim.available():
        pytest.skip("reference sources not present (GPU box)")
    from ape_b200.modeling import build_model

    spec = getattr(configs, spec_name)
    ref, _ = ref_model.build_reference_model(spec, num_text=16)
    eng = build_model(spec, num_text=16)
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in eng.state_dict().items()}
    assert sorted(a) == sorted(b), (sorted(set(a) - set(b))[:5], sorted(set(b) - set(a))[:5])
    assert a == b
