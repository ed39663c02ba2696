"""End-to-end flow: the oracle composition against the fixture produced by the reference's
own STLLMForCausalLM.forward(samples) (tests/golden/make_golden.py golden_flow).  CPU, fp32."""
import json
import os
import re

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


FIXTURES = {True: "stllm_flow.npz", False: "stllm_flow_minigpt4.npz"}


@pytest.mark.parametrize("text", [True, False], ids=["instructblip_text", "minigpt4_bos_img_start8"])
def test_oracle_flow_matches_reference(text):
    g = np.load(os.path.join(GOLDEN, FIXTURES[text]))
    shapes = reference_shapes()
    if not text:   # the MiniGPT-4 variant strips the Q-Former text branch (st_llm.py:277-283)
        shapes = {k: v for k, v in shapes.items()
                  if not any(t in k for t in ("word_embeddings", "Qformer.bert.embeddings.position_embeddings"))
                  and not re.search(r"Qformer\.bert\.encoder\.layer\.\d+\.(output|intermediate)\.", k)}
    sd = flow_state_dict(shapes)
    # the seeded weights depend on torch's CPU RNG stream: a drift must FAIL here, not silently skip the comparison
    assert abs(weights_checksum(sd) - float(g["weights_checksum"])) <= 1e-6 * float(g["weights_checksum"]), \
        "torch CPU RNG produced different seeded weights than when the fixture was made: regenerate tests/golden"
    np.random.seed(1237)
    with torch.no_grad():
        out = oracle_flow(sd, flow_samples(text), text=text)
    assert np.array_equal(out["mask"].numpy(), g["mask"])                 # bit-exact index path
    assert out["img_len"] == int(g["img_len"]) and out["mask_img_len"] == int(g["mask_img_len"])
    assert np.array_equal(out["attention_mask"].numpy(), g["attention_mask"])
    assert np.array_equal(out["unmask_attention_mask"].numpy(), g["unmask_attention_mask"])
    assert np.array_equal(out["targets"].numpy(), g["targets"])
    assert_close(out["inputs_embeds"][:, :, :16], g["inputs_embeds_head"], 2e-4, 1e-4, "assembled inputs_embeds")
    assert_close(out["unmask_inputs_embeds"][:, :, :16], g["unmask_inputs_embeds_head"], 2e-4, 1e-4, "unmasked inputs")
    assert tuple(out["logits"].shape) == tuple(g["logits_shape"])
    assert_close(out["logits"][:, :, :64], g["logits_head"], 2e-4, 1e-4, "flow logits")
    assert_close(out["loss"].reshape(1), g["loss"].reshape(1), 1e-4, 1e-5, "flow loss (CE + MVM)")
