"""End-to-end flow: the oracle composition against the fixture produced by the reference's
own STLLMForCausalLM.forward(samples) (tests/golden/make_golden.py golden_flow).  CPU, fp32."""
import json
import os

import numpy as np
import pytest
import torch

from tests.flow_fixture import flow_samples, flow_state_dict, oracle_flow, weights_checksum
from tests.util import GOLDEN, assert_close


def reference_shapes():
    with open(os.path.join(GOLDEN, "stllm_flow_state_dict_keys.json")) as f:
        keys = json.load(f)
    return {k: tuple(v) for k, v in keys.items()
            if "position_ids" not in k and ".stllm_model.embed_tokens." not in k}


def test_oracle_flow_matches_reference():
    g = np.load(os.path.join(GOLDEN, "stllm_flow.npz"))
    sd = flow_state_dict(reference_shapes())
    if abs(weights_checksum(sd) - float(g["weights_checksum"])) > 1e-6 * float(g["weights_checksum"]):
        pytest.skip("torch CPU RNG produced different seeded weights than when the fixture was made")
    np.random.seed(1237)
    with torch.no_grad():
        out = oracle_flow(sd, flow_samples())
    assert np.array_equal(out["mask"].numpy(), g["mask"])                 # bit-exact index path
    assert out["img_len"] == int(g["img_len"]) and out["mask_img_len"] == int(g["mask_img_len"])
    assert tuple(out["logits"].shape) == tuple(g["logits_shape"])
    assert_close(out["logits"][:, :, :64], g["logits_head"], 2e-4, 1e-4, "flow logits")
    assert_close(out["loss"].reshape(1), g["loss"].reshape(1), 1e-4, 1e-5, "flow loss (CE + MVM)")
