"""Validity of the packed (fused / pre-transposed / norm-folded) fp16 weight copies the kernels read.

Each packing module (LlamaModel, VisionTransformer, EVAVisionTransformer_BTAdapter, the Q-Former BertModel) builds its
packed copies once, at the first forward.  They must be rebuilt whenever a parameter changes; the ways that happens:
  * `.to()` / `.half()` / `.cuda()`                      -> `_apply` override of the module (it is called on every submodule);
  * `load_state_dict` on the module OR ON ANY PARENT     -> `_load_from_state_dict` (nn.Module.load_state_dict recurses through
    it and never calls a child's `load_state_dict`, so overriding that method on the child is not enough);
  * in-place edits (`p.data.copy_(...)`, optimizer steps) -> every parameter's (storage address, version counter) is recorded
    when the pack is built and compared at every eager forward (a CUDA-graph replay runs no Python: its weights are baked).
    Edits through `p.data` bypass autograd's version counter and cannot be seen: call `repack()` after those.
"""


class PackedWeightsMixin:
    _packed = None
    _packed_sig = None

    def _weights_signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _mark_packed(self):
        self._packed_sig = self._weights_signature()

    def _packed_valid(self):
        return self._packed is not None and self._packed_sig == self._weights_signature()

    def repack(self):
        """drop the packed copies (rebuilt at the next forward)"""
        self._packed = None
        self._packed_sig = None

    def _load_from_state_dict(self, *args, **kwargs):
        self.repack()
        return super()._load_from_state_dict(*args, **kwargs)
