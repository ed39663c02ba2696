"""The `stllm.models` loader API and the reference's OWN callers walking the product model on the CPU.

* load_model / load_preprocess / load_model_and_preprocess / model_zoo: signatures, defaults and error behaviour of
  reference stllm/models/__init__.py:27-158 and base_model.py:29-79 (offline: config dicts, injected tokenizers);
* where the reference tree exists (build container): `Chat.upload_video` + `Chat.answer` (conversation.py:181-340) and
  `mv_bench.infer_mvbench` (mv_bench.py:229-283), loaded per file with decord / mmengine stubbed, run UNMODIFIED against a
  product STLLMForCausalLM whose kernels are replaced by torch stand-ins -- i.e. "drops into demo.py / stllm/test
  unchanged": every attribute, call signature and return shape those callers rely on is exercised."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from tests.fake_tokenizer import FakeTokenizer


def _overrides(video_input="all", text=False):
    from stllm_b200.qformer import BertConfig
    V = 64
    return dict(llama_model="", llama_config=dict(vocab_size=V, hidden_size=256, intermediate_size=512, num_hidden_layers=1,
                                                  num_attention_heads=2, max_position_embeddings=2048, pad_token_id=0,
                                                  bos_token_id=1, eos_token_id=2),
                vit_depth=1, qformer_config=BertConfig(num_hidden_layers=2, vocab_size=97), llama_hidden=256,
                video_input=video_input, residual_size=2, qformer_text_input=text,
                tokenizer=FakeTokenizer(97), llama_tokenizer=FakeTokenizer(V, pad_token_id=0, bos_token_id=1, eos_token_id=2))


def test_load_model_family_and_model_zoo(tmp_path):
    from stllm_b200 import models_api as A
    from stllm_b200.st_llm import STLLMForCausalLM
    zoo = A.ModelZoo()
    assert dict(zoo)["st_llm_hf"] == ["instructblip_vicuna0", "instructblip_vicuna0_btadapter", "minigpt4_vicuna0",
                                      "minigpt4_vicuna0_btadapter"]
    assert len(zoo) == 4 and "st_llm_hf" in str(zoo)
    model = A.load_model("st_llm_hf", "instructblip_vicuna0", is_eval=True, device="cpu", **_overrides())
    assert isinstance(model, STLLMForCausalLM) and not model.training
    assert all(p.dtype == torch.float32 for p in model.parameters())              # device == "cpu" -> model.float()
    assert model.get_model().stllm_model.visual_encoder.__class__.__name__ == "VisionTransformer"
    bt = A.load_model("st_llm_hf", "minigpt4_vicuna0_btadapter", **dict(_overrides(), vit_depth=3))   # 3 adapter blocks
    assert "BTAdapter" in bt.get_model().stllm_model.visual_encoder.__class__.__name__     # vit_model from the type's defaults
    assert bt.training                                                             # is_eval defaults to False
    with pytest.raises(AssertionError, match="Unknown model type"):
        A.load_model("st_llm_hf", "no_such_type", **_overrides())
    with pytest.raises(KeyError):
        A.load_model("no_such_arch", "instructblip_vicuna0")
    # checkpoint= : base_model.load_checkpoint semantics ('model' wrapper, strict=False, invalid path -> RuntimeError)
    ck = {"model": {"model.stllm_model.query_tokens": torch.full((1, 32, 768), 0.25)}}
    f = tmp_path / "ft.pth"
    torch.save(ck, f)
    m2 = A.load_model("st_llm_hf", "instructblip_vicuna0", checkpoint=str(f), **_overrides())
    assert torch.equal(m2.state_dict()["model.stllm_model.query_tokens"], ck["model"]["model.stllm_model.query_tokens"])
    with pytest.raises(RuntimeError, match="checkpoint url or path is invalid"):
        A.load_model("st_llm_hf", "instructblip_vicuna0", checkpoint=str(tmp_path / "missing.pth"), **_overrides())


def test_load_preprocess_and_load_model_and_preprocess():
    from stllm_b200 import models_api as A
    vis, txt = A.load_preprocess({})                     # nothing named -> identity processors (models/__init__.py:76-80)
    assert set(vis) == set(txt) == {"train", "eval"} and vis["eval"]("x") == "x" and isinstance(txt["train"], A.BaseProcessor)
    with pytest.raises(KeyError, match="blip2_image_eval"):
        A.load_preprocess({"vis_processor": {"eval": {"name": "blip2_image_eval", "image_size": 224}}})

    @A.register_processor("blip2_image_train")
    @A.register_processor("blip2_image_eval")
    @A.register_processor("blip_caption")
    class Echo(A.BaseProcessor):
        @classmethod
        def from_config(cls, cfg=None):
            o = cls()
            o.cfg = dict(cfg or {})
            return o
    model, vis, txt = A.load_model_and_preprocess("st_llm_hf", "instructblip_vicuna0", is_eval=True, device="cpu",
                                                  **_overrides())
    assert not model.training and vis["eval"].cfg == {"name": "blip2_image_eval", "image_size": 224}
    assert txt["train"].cfg == {"name": "blip_caption"}


# ----------------------------------------------------------------------------------------------------------------------
# the reference's own callers
# ----------------------------------------------------------------------------------------------------------------------
def _load_reference_callers():
    import importlib.util
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present on this box")
    ref_shim.load_reference()                          # registers the stllm.* package shells and stubs

    def stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    def load(rel, name):
        spec = importlib.util.spec_from_file_location(name, f"{ref_shim.REF}/{rel}")
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m
    for pkg in ("stllm.test", "stllm.conversation", "stllm.test.mvbench"):
        if pkg not in sys.modules:
            stub(pkg).__path__ = [f"{ref_shim.REF}/" + pkg.replace(".", "/")]
    stub("decord", VideoReader=object, cpu=lambda *a: None)                 # video decoding: not on the path under test
    stub("mmengine"); stub("mmengine.fileio", FileClient=lambda *a, **k: None)
    stub("imageio"); stub("cv2")
    load("stllm/test/video_transforms.py", "stllm.test.video_transforms")
    stub("stllm.test.video_utils", load_video=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no decoding here")))
    conv = load("stllm/conversation/conversation.py", "stllm.conversation.conversation")
    mvc = load("stllm/conversation/mvbench_conversation.py", "stllm.conversation.mvbench_conversation")
    mvb = load("stllm/test/mvbench/mv_bench.py", "stllm.test.mvbench.mv_bench")
    return conv, mvc, mvb


class _FakeKernels:
    """torch stand-ins for the two kernel entry points the callers reach: encode_img and the LLM forward"""

    def __init__(self, model, monkeypatch):
        self.model = model
        sm = model.get_model().stllm_model
        self.calls = []
        D = model.config.hidden_size

        def encode_img(image, text=None):
            self.calls.append(("encode_img", tuple(image.shape), text))
            assert image.dim() == 4 and image.shape[1:] == (3, 224, 224)
            g = torch.Generator().manual_seed(0)
            return torch.randn(image.shape[0], 32, D, generator=g), torch.ones(image.shape[0], 32, dtype=torch.long), True
        monkeypatch.setattr(sm, "encode_img", encode_img)

        class Cache:
            def __init__(self):
                self.hist = [[]]

            def expand_batch(self, n):
                self.hist = [list(self.hist[0]) for _ in range(n)]

            def reorder(self, idx):
                self.hist = [list(self.hist[i]) for i in idx.tolist()]
        table = torch.randn(53, model.config.vocab_size, generator=torch.Generator().manual_seed(3)) * 2
        outer = self

        def forward(samples=None, inputs_embeds=None, input_ids=None, attention_mask=None, past_key_values=None,
                    use_cache=False, cache_max_len=None, **kw):
            O = type("O", (), {})()
            if past_key_values is None:
                outer.calls.append(("prefill", tuple(inputs_embeds.shape)))
                cache = Cache()
                rows = torch.stack([table[len(h) % 53] for h in cache.hist])
                O.logits = rows[:, None, :].expand(-1, inputs_embeds.shape[1], -1)
            else:
                cache = past_key_values
                for h, t in zip(cache.hist, input_ids.reshape(-1).tolist()):
                    h.append(t)
                O.logits = torch.stack([table[(len(h) * 7 + h[-1]) % 53] for h in cache.hist])[:, None, :]
            O.past_key_values = cache
            return O
        monkeypatch.setattr(model, "forward", forward)


@pytest.mark.parametrize("video_input", ["all", "mean", "residual"])
def test_reference_chat_upload_video_and_answer_walk_the_product_model(video_input, monkeypatch):
    conv_mod, _, _ = _load_reference_callers()
    from stllm_b200 import models_api as A
    model = A.load_model("st_llm_hf", "instructblip_vicuna0", is_eval=True, device="cpu", **_overrides(video_input))
    sm = model.get_model().stllm_model
    if video_input == "residual":
        sm.up_proj.weight.data.normal_(0, 0.02)
    fake = _FakeKernels(model, monkeypatch)
    chat = conv_mod.Chat(model, device="cpu")                       # conversation.py:181-203, UNMODIFIED
    assert chat.model is sm and chat.LLM is model
    state = conv_mod.CONV_instructblip_Vicuna0.copy() if hasattr(conv_mod, "CONV_instructblip_Vicuna0") else \
        conv_mod.CONV_VISION.copy()
    from PIL import Image
    frames = [Image.fromarray(np.random.default_rng(i).integers(0, 255, (240, 320, 3), dtype=np.uint8)) for i in range(4)]
    img_list = []
    msg = chat.upload_video(frames, state, img_list, num_frame=4, text="what happens ?")   # conversation.py:274-299
    assert msg == "Received." and fake.calls[0][0] == "encode_img" and fake.calls[0][1] == (4, 3, 224, 224)
    L = {"all": 4 * 32, "mean": 32, "residual": 2 * 32}[video_input]
    assert tuple(img_list[0].shape) == (1, L, model.config.hidden_size)
    chat.ask("describe the video", state)
    text, tokens = chat.answer(conv=state, img_list=img_list, num_beams=5, do_sample=False, temperature=1,
                               max_new_tokens=12, max_length=2000)            # demo.py:59-65 / msvd_qa.py:108-115
    assert isinstance(text, str) and isinstance(tokens, np.ndarray) and 1 <= tokens.size <= 13
    assert any(c[0] == "prefill" and c[1][0] == 1 and c[1][2] == model.config.hidden_size for c in fake.calls)
    text2, tokens2 = chat.answer(conv=state, img_list=img_list, num_beams=1, do_sample=False, max_new_tokens=6)
    assert tokens2.size <= 7


def test_reference_infer_mvbench_walks_the_product_model(monkeypatch):
    _, mvc, mvb = _load_reference_callers()
    from stllm_b200 import models_api as A
    model = A.load_model("st_llm_hf", "instructblip_vicuna0", is_eval=True, device="cpu", **_overrides("all", text=True))
    fake = _FakeKernels(model, monkeypatch)
    # mv_bench.py:241 / mvbench_conversation.py:103-104 hard-code "cuda:0": map it to the CPU for this attribute walk
    real_to = torch.Tensor.to
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: real_to(
        self, *[("cpu" if isinstance(x, str) and x.startswith("cuda") else x) for x in a], **k))
    sample = {"video": torch.zeros(4 * 3, 224, 224), "question": "Question: what is shown ?\nOptions:\n(A) a\n(B) b\n",
              "answer": "(A) a"}
    out = mvb.infer_mvbench(model, sample, system="", question_prompt="\nOnly give the best option.",
                            answer_prompt="Best option:(", return_prompt="(", all_token=True, ask_simple=True)
    assert isinstance(out, str) and out.startswith("(")
    assert fake.calls[0] == ("encode_img", (4, 3, 224, 224), sample["question"])
    assert mvb.check_ans("(A) a", "(A) a") and not mvb.check_ans("(B) b", "(A) a")       # mv_bench.py:285-300
