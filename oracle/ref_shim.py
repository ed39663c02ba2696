"""Per-file loader for the UNMODIFIED reference modules (eva_vit.py, Qformer.py, st_llm.py, utils.py).

TEST INFRASTRUCTURE.  Used only where the reference tree exists (this container: /root/reference) to
(a) generate tests/golden fixtures and (b) time the reference's own CPU path.  Nothing on the GPU box
imports this.  `import stllm` fails offline (omegaconf/timm/peft/... missing), so the three
compute files are loaded one by one with small stubs for the missing third-party packages and five shims
for the transformers 4.28 -> 5.x API drift.  No reference source is copied.
"""
import importlib.util
import os
import sys
import types

import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
# Where the reference's source files are read from: $STLLM_REFERENCE, the read-only tree of the build container, or
# oracle/_ref/ -- a git-ignored copy of the reference's model / common files that `__graft_entry__.build()` materialises while
# the tree exists (oracle/materialize_ref.py), so that the GPU box -- which has no /root/reference -- can still time the
# reference's OWN modules as the CPU baseline (`bench.py --impl reference`, cpu_baseline.kind = "reference").
REF = os.environ.get("STLLM_REFERENCE") or ("/root/reference" if os.path.isfile("/root/reference/stllm/models/st_llm.py")
                                             else os.path.join(_HERE, "_ref"))
_loaded = None


def available():
    return os.path.isfile(os.path.join(REF, "stllm/models/st_llm.py"))


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, f"{REF}/{rel}")
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def load_reference():
    """Returns a namespace with the reference modules: .eva_vit .Qformer .st_llm .utils .blip2"""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    import transformers  # must precede the fake timm (transformers probes timm.__spec__)
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    # transformers 4.28 -> 5.x drift used by Qformer.py:39-44,697,935
    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer
    mu.find_pruneable_heads_and_indices = None
    _orig_init = mu.PreTrainedModel.init_weights
    mu.PreTrainedModel.init_weights = lambda s: s.post_init() if not hasattr(s, "all_tied_weights_keys") else _orig_init(s)
    mu.PreTrainedModel.get_head_mask = lambda s, hm, n, c=False: [None] * n
    # missing third-party packages (init helpers / unused imports only)
    _stub("timm"); _stub("timm.models"); _stub("timm.models.hub")
    _stub("timm.models.layers", drop_path=lambda x, p=0., t=False: x, to_2tuple=lambda x: (x, x),
          trunc_normal_=nn.init.trunc_normal_)
    _stub("timm.models.registry", register_model=lambda f: f)
    _stub("omegaconf", OmegaConf=object)
    _stub("peft", LoraConfig=object, get_peft_model=None,
          peft_model=types.SimpleNamespace(PeftModelForCausalLM=type("P", (), {})))
    _stub("peft.utils", PeftType=object)
    for p in ("stllm", "stllm.common", "stllm.models"):
        _stub(p).__path__ = [f"{REF}/" + p.replace(".", "/")]
    _stub("stllm.common.utils", is_url=lambda s: False, get_abs_path=lambda s: s)
    _load("stllm/common/registry.py", "stllm.common.registry")
    _load("stllm/common/dist_utils.py", "stllm.common.dist_utils")
    _load("stllm/common/logger.py", "stllm.common.logger")
    bm = _load("stllm/models/base_model.py", "stllm.models.base_model")
    sys.modules["stllm.models"].BaseModel = bm.BaseModel
    st = _load("stllm/models/st_llm.py", "stllm.models.st_llm")
    ns = types.SimpleNamespace(
        st_llm=st,
        eva_vit=sys.modules["stllm.models.eva_vit"],
        Qformer=sys.modules["stllm.models.Qformer"],
        blip2=sys.modules["stllm.models.blip2"],
        utils=sys.modules["stllm.models.utils"],
    )
    _loaded = ns
    return ns


def build_vit(ref, img_size, embed_dim, depth, num_heads, mlp_ratio=4.3637, patch=14):
    """the create_eva_vit_g architecture (eva_vit.py:416-428) at arbitrary size, without the download"""
    from functools import partial
    return ref.eva_vit.VisionTransformer(
        img_size=img_size, patch_size=patch, use_mean_pooling=False, embed_dim=embed_dim, depth=depth,
        num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=True, drop_path_rate=0.0,
        norm_layer=partial(nn.LayerNorm, eps=1e-6), use_checkpoint=False).eval()


def build_qformer(ref, hidden, heads, inter, layers, enc_width, n_query, vocab=30522, max_pos=512, text=False):
    """Blip2Base.init_Qformer (blip2.py:46-59) at arbitrary size + the strip of st_llm.py:277-288"""
    import torch
    from transformers import BertConfig
    cfg = BertConfig(hidden_size=hidden, num_attention_heads=heads, intermediate_size=inter,
                     num_hidden_layers=layers, vocab_size=vocab, max_position_embeddings=max_pos)
    cfg.encoder_width = enc_width
    cfg.add_cross_attention = True
    cfg.cross_attention_freq = 2
    cfg.query_length = n_query
    qf = ref.Qformer.BertLMHeadModel(config=cfg)
    query_tokens = nn.Parameter(torch.zeros(1, n_query, hidden))
    query_tokens.data.normal_(mean=0.0, std=cfg.initializer_range)
    if not text:
        qf.bert.embeddings.word_embeddings = None
        qf.bert.embeddings.position_embeddings = None
        for layer in qf.bert.encoder.layer:
            layer.output = None
            layer.intermediate = None
    qf.cls = None
    return qf.eval(), query_tokens
