"""Name -> class registry with the reference's lookup API (stllm/common/registry.py:83-109,236-237):
`registry.register_model(name)` decorator and `registry.get_model_class(name)`.  When the reference package is
importable (a full ST-LLM install), `install_into_reference()` registers this implementation under the same name
in the reference's own registry so demo.py / stllm/test harnesses pick it up unchanged."""


class Registry:
    mapping = {"model_name_mapping": {}}

    @classmethod
    def register_model(cls, name):
        def wrap(model_cls):
            cls.mapping["model_name_mapping"][name] = model_cls
            return model_cls
        return wrap

    @classmethod
    def get_model_class(cls, name):
        return cls.mapping["model_name_mapping"].get(name, None)

    @classmethod
    def list_models(cls):
        return sorted(cls.mapping["model_name_mapping"].keys())


registry = Registry()


def install_into_reference():
    """Overwrite the reference registry's "st_llm_hf" entry with the B200 implementation."""
    from stllm.common.registry import registry as ref_registry  # noqa: needs the reference package installed
    from .st_llm import STLLMForCausalLM
    ref_registry.mapping["model_name_mapping"]["st_llm_hf"] = STLLMForCausalLM
    return ref_registry
