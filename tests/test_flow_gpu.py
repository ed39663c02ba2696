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


FIXTURES = {True: "stllm_flow.npz", False: "stllm_flow_minigpt4.npz"}


@pytest.mark.parametrize("text", [True, False], ids=["instructblip_text", "minigpt4_bos_img_start8"])
def test_flow_vs_reference_golden_and_oracle(text):
    """text=False is the MiniGPT-4 layout: BOS + 7 prefix tokens, video at position 8 = the hard-coded img_start of the MVM
    slices (st_llm.py:71,519-530; SURVEY a21')"""
    g = np.load(os.path.join(GOLDEN, FIXTURES[text]))
    model, sd = build_product_flow(torch.device("cuda", 0), text=text)
    samples = flow_samples(text)
    dev_samples = dict(samples, image=samples["image"].cuda())
    np.random.seed(1237)
    out = model(samples=dev_samples)
    sm = model.get_model().stllm_model
    # integer / index path: bit-exact
    assert np.array_equal(sm.mask.cpu().numpy(), g["mask"])
    assert sm.img_len == int(g["img_len"]) and sm.mask_img_len == int(g["mask_img_len"])
    assert np.array_equal(sm.residual_index.cpu().numpy(), g["residual_index"])
    assert tuple(out.logits.shape) == tuple(g["logits_shape"])
    np.random.seed(1237)
    emb, att, emb_u, att_u, targets = sm(dev_samples)
    assert np.array_equal(att.cpu().numpy(), g["attention_mask"])
    assert np.array_equal(att_u.cpu().numpy(), g["unmask_attention_mask"])
    assert np.array_equal(targets.cpu().numpy(), g["targets"])
    # the seeded weights must be the fixture's (a torch RNG drift fails loudly instead of weakening the test)
    assert abs(weights_checksum(sd) - float(g["weights_checksum"])) <= 1e-6 * float(g["weights_checksum"]), \
        "seeded weights differ from the fixture's: regenerate tests/golden with make_golden.py"
    # fp16 kernels vs the reference's fp32 CPU forward: logits magnitude ~1, 2+2+1 layers deep
    assert_close(emb[:, :, :16], g["inputs_embeds_head"], atol=4e-3, rtol=4e-3, what="assembled inputs vs reference")
    assert_close(out.logits[:, :, :64], g["logits_head"], atol=2e-2, rtol=1e-2, what="flow logits vs reference")
    ref_head = torch.from_numpy(g["logits_head"])
    rel = ((out.logits[:, :, :64].float().cpu() - ref_head).norm() / ref_head.norm()).item()
    assert rel < 4e-3, rel
    assert abs(out.loss.item() - float(g["loss"])) < 1e-2 * max(1.0, abs(float(g["loss"]))), (out.loss.item(), g["loss"])
    # and against the oracle on the very same weights
    np.random.seed(1237)
    with torch.no_grad():
        ref = oracle_flow({k: v.float().cpu() for k, v in model.state_dict().items()}, samples, text=text)
    assert np.array_equal(sm.mask.cpu().numpy(), ref["mask"].numpy())
    assert_close(out.logits, ref["logits"], atol=2e-2, rtol=1e-2, what="flow logits vs oracle")
    rel = ((out.logits.float().cpu() - ref["logits"]).norm() / ref["logits"].norm()).item()
    assert rel < 4e-3, rel
    assert abs(out.loss.item() - ref["loss"].item()) < 1e-2 * max(1.0, abs(ref["loss"].item()))
