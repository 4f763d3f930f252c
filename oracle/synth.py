"""oracle/synth.py — TEST INFRASTRUCTURE: re-exports the deterministic synthetic weight / input
recipe (ape_b200/synthetic.py) under the name the tests and golden generators use."""
from ape_b200.synthetic import fill_state_dict, image, suppress_invalid_anchor_logits, tensor_for, text_features  # noqa: F401
