"""GPU: the whole training-branch hot path (STLLMForCausalLM.forward(samples): encode_img with Q-Former text input,
global-local residual aggregation, dynamic mask, prompt_wrap/concat/targets, masked + unmasked LLM passes, MVM + CE
loss) through the C ABI, against the fixture produced by the REFERENCE's own forward and against the oracle."""
import os

import numpy as np
import pytest
import torch

from tests.flow_fixture import build_product_flow, flow_samples, oracle_flow, weights_checksum
from tests.util import GOLDEN, assert_close

pytestmark = pytest.mark.gpu


def test_flow_vs_reference_golden_and_oracle():
    g = np.load(os.path.join(GOLDEN, "stllm_flow.npz"))
    model, sd = build_product_flow(torch.device("cuda", 0))
    samples = flow_samples()
    dev_samples = dict(samples, image=samples["image"].cuda())
    np.random.seed(1237)
    out = model(samples=dev_samples)
    sm = model.get_model().stllm_model
    # integer / index path: bit-exact
    assert np.array_equal(sm.mask.cpu().numpy(), g["mask"])
    assert sm.img_len == int(g["img_len"]) and sm.mask_img_len == int(g["mask_img_len"])
    assert np.array_equal(sm.residual_index.cpu().numpy(), g["residual_index"])
    assert tuple(out.logits.shape) == tuple(g["logits_shape"])
    same_weights = abs(weights_checksum(sd) - float(g["weights_checksum"])) <= 1e-6 * float(g["weights_checksum"])
    if same_weights:
        # fp16 kernels vs the reference's fp32 CPU forward: logits magnitude ~1, 2+2+1 layers deep
        assert_close(out.logits[:, :, :64], g["logits_head"], atol=3e-2, rtol=2e-2, what="flow logits vs reference")
        assert abs(out.loss.item() - float(g["loss"])) < 2e-2 * max(1.0, abs(float(g["loss"]))), (out.loss.item(), g["loss"])
    # and against the oracle on the very same weights (always possible)
    np.random.seed(1237)
    with torch.no_grad():
        ref = oracle_flow({k: v.float().cpu() for k, v in model.state_dict().items()}, samples)
    assert np.array_equal(sm.mask.cpu().numpy(), ref["mask"].numpy())
    assert_close(out.logits, ref["logits"], atol=3e-2, rtol=2e-2, what="flow logits vs oracle")
    rel = ((out.logits.float().cpu() - ref["logits"]).norm() / ref["logits"].norm()).item()
    assert rel < 5e-3, rel
    assert abs(out.loss.item() - ref["loss"].item()) < 2e-2 * max(1.0, abs(ref["loss"].item()))
