"""The `stllm.models` loader API on top of the B200 implementation (reference stllm/models/__init__.py:27-158,
base_model.py:29-79): `load_model`, `load_preprocess`, `load_model_and_preprocess`, `model_zoo`, with the reference's
signatures, defaults and error behaviour.

The reference resolves `model_type` through OmegaConf YAML files (stllm/configs/models/*.yaml) and downloads tokenizers /
checkpoints; neither OmegaConf nor a network exists here, so the four model types' default configurations are stated below
as plain dicts (same keys and values as the YAMLs) and every loader takes `**overrides` that are merged over the `model`
section -- e.g. `llama_model=<HF directory>`, `ckpt=<ST-LLM checkpoint>`, injected `tokenizer=` / `llama_tokenizer=` objects,
or `llama_config=dict(...)` + reduced depths for a random-init model of the same architecture.
Processors (`stllm/processors`, CPU image / caption preprocessing) are outside the hot path: `load_preprocess` resolves them
through a processor registry that holds `BaseProcessor` (identity) and whatever the caller registered -- the reference's own
registry is used when that package is importable.
"""
import copy
import logging
import os

import torch

from .registry import registry

_INSTRUCTBLIP_QFORMER = ("https://storage.googleapis.com/sfr-vision-language-research/LAVIS/models/InstructBLIP/"
                         "instruct_blip_vicuna7b_trimmed.pth")
_BLIP2_QFORMER = "https://storage.googleapis.com/sfr-vision-language-research/LAVIS/models/BLIP2/blip2_pretrained_flant5xxl.pth"
_PREPROCESS = {"vis_processor": {"train": {"name": "blip2_image_train", "image_size": 224},
                                 "eval": {"name": "blip2_image_eval", "image_size": 224}},
               "text_processor": {"train": {"name": "blip_caption"}, "eval": {"name": "blip_caption"}}}


def _model_cfg(q_former_model, llama_model, vit_model=None):
    cfg = {"arch": "st_llm_hf", "image_size": 224, "drop_path_rate": 0, "use_grad_checkpoint": False,
           "vit_precision": "fp16", "freeze_vit": True, "freeze_qformer": True, "q_former_model": q_former_model,
           "num_query_token": 32, "prompt": "", "llama_model": llama_model}
    if vit_model is not None:
        cfg["vit_model"] = vit_model
    return cfg


# stllm/configs/models/{instructblip_vicuna0, instructblip_vicuna0_btadapter, minigpt4_vicuna0, minigpt4_vicuna0_btadapter}.yaml
MODEL_TYPE_CONFIGS = {
    "instructblip_vicuna0": {"model": _model_cfg(_INSTRUCTBLIP_QFORMER, "/path/to/vicuna-7b-v1.1"), "preprocess": _PREPROCESS},
    "instructblip_vicuna0_btadapter": {"model": _model_cfg(_INSTRUCTBLIP_QFORMER, "/path/to/vicuna-7b-v1.1", "eva_btadapter_g"),
                                       "preprocess": _PREPROCESS},
    "minigpt4_vicuna0": {"model": _model_cfg(_BLIP2_QFORMER, "/path/to/vicuna-7b"), "preprocess": _PREPROCESS},
    "minigpt4_vicuna0_btadapter": {"model": _model_cfg(_BLIP2_QFORMER, "/path/to/vicuna-7b", "eva_btadapter_g"),
                                   "preprocess": _PREPROCESS},
}


def default_config(model_cls, model_type):
    """the parsed default configuration of a model type (OmegaConf.load(default_config_path) in the reference)"""
    model_cls.default_config_path(model_type)        # asserts "Unknown model type ..." like base_model.py:74-79
    return copy.deepcopy(MODEL_TYPE_CONFIGS[model_type])


class BaseProcessor:
    """stllm/processors/base_processor.py: the identity processor returned when a config names none"""

    def __init__(self):
        self.transform = lambda x: x

    def __call__(self, item):
        return self.transform(item)

    @classmethod
    def from_config(cls, cfg=None):
        return cls()

    def build(self, **kwargs):
        return self.from_config(kwargs)


_processors = {}


def register_processor(name):
    def wrap(cls):
        _processors[name] = cls
        return cls
    return wrap


def _processor_class(name):
    if name in _processors:
        return _processors[name]
    try:                                            # a full ST-LLM install next to this package: use its processors
        from stllm.common.registry import registry as ref_registry
        cls = ref_registry.get_processor_class(name)
        if cls is not None:
            return cls
    except ImportError:
        pass
    raise KeyError(f"processor {name!r} is not registered: the reference's CPU processors (stllm/processors) are outside "
                   f"the hot path; register one with stllm_b200.models_api.register_processor or install the reference")


def load_model(name, model_type, is_eval=False, device="cpu", checkpoint=None, **overrides):
    """models/__init__.py:27-58"""
    model_cls = registry.get_model_class(name)
    if model_cls is None:
        raise KeyError(f"unknown model {name!r}; registered: {registry.list_models()}")
    model = model_cls.from_pretrained(model_type=model_type, **overrides)
    if checkpoint is not None:
        model.load_checkpoint(checkpoint)
    if is_eval:
        model.eval()
    if device == "cpu" or device == torch.device("cpu"):
        model = model.float()
    return model.to(device)


def load_preprocess(config):
    """models/__init__.py:61-110: {"train","eval"} processors for visual and text inputs; BaseProcessor where none is named"""
    def build(cfg):
        return _processor_class(cfg.get("name")).from_config(cfg) if cfg is not None else BaseProcessor()
    vis_cfg, txt_cfg = config.get("vis_processor"), config.get("text_processor")
    vis = {k: build(vis_cfg.get(k) if vis_cfg is not None else None) for k in ("train", "eval")}
    txt = {k: build(txt_cfg.get(k) if txt_cfg is not None else None) for k in ("train", "eval")}
    return vis, txt


def load_model_and_preprocess(name, model_type, is_eval=False, device="cpu", **overrides):
    """models/__init__.py:113-158"""
    model_cls = registry.get_model_class(name)
    if model_cls is None:
        raise KeyError(f"unknown model {name!r}; registered: {registry.list_models()}")
    model = model_cls.from_pretrained(model_type=model_type, **overrides)
    if is_eval:
        model.eval()
    cfg = default_config(model_cls, model_type)
    if cfg is not None and cfg.get("preprocess") is not None:
        vis_processors, txt_processors = load_preprocess(cfg["preprocess"])
    else:
        vis_processors, txt_processors = None, None
        logging.info(f"No default preprocess for model {name} ({model_type}).")
    if device == "cpu" or device == torch.device("cpu"):
        model = model.float()
    return model.to(device), vis_processors, txt_processors


class ModelZoo:
    """models/__init__.py:161-198: string table of registered architectures and their model types"""

    def __init__(self):
        self.model_zoo = {k: list(v.PRETRAINED_MODEL_CONFIG_DICT.keys())
                          for k, v in registry.mapping["model_name_mapping"].items()}

    def __str__(self):
        return ("=" * 50 + "\n" + f"{'Architectures':<30} {'Types'}\n" + "=" * 50 + "\n"
                + "\n".join(f"{name:<30} {', '.join(types)}" for name, types in self.model_zoo.items()))

    def __iter__(self):
        return iter(self.model_zoo.items())

    def __len__(self):
        return sum(len(v) for v in self.model_zoo.values())


def _checkpoint_state_dict(path):
    """base_model.py:29-52: a local checkpoint file, optionally wrapped in {'model': ...}"""
    if not os.path.isfile(path):
        raise RuntimeError("checkpoint url or path is invalid")
    ckpt = torch.load(path, map_location="cpu")
    return ckpt["model"] if "model" in ckpt.keys() else ckpt
