"""Video-token prefill pipeline: frames -> encode_img -> aggregate -> [video ‖ prompt] -> LLM prefill -> logits.

This is the inference call stack of the reference (Chat.upload_video -> encode_img -> aggregation,
conversation.py:274-293; get_context_emb_sim + first `generate` step, conversation.py:322-340,231) as ONE
callable with static shapes, captured into a CUDA graph so the ~1100 kernel launches of a step cost one
cudaGraphLaunch (B200-first: streams + graphs instead of a tracing compiler).  Used by bench.py, smoke() and the
clip-parallel multi-GPU driver; the model classes stay usable eagerly (demo.py-style) without it.
"""
import torch

from . import ops


def build_model(llama_layers=32, vit_depth=39, qformer_layers=12, vocab_size=32000, video_input="all", residual_size=4,
                device="cuda", seed=0, llama_hidden=4096, llama_inter=11008, llama_heads=32, use_mask=False,
                mvm_decode=False, qformer_text_input=False, tokenizer=None, llama_tokenizer=None):
    """Random-init ST-LLM of the real architecture (EVA-ViT-g/14 + Q-Former + Vicuna-7B dims), built directly in
    fp16 on the device (no checkpoints/tokenizers exist offline).  Parameters are created on the GPU so that a 7B
    model does not have to be initialised on the host."""
    from .qformer import BertConfig
    from .st_llm import STLLMForCausalLM
    torch.manual_seed(seed)
    cfg = dict(llama_config=dict(vocab_size=vocab_size, hidden_size=llama_hidden, intermediate_size=llama_inter,
                                 num_hidden_layers=llama_layers, num_attention_heads=llama_heads),
               video_input=video_input, residual_size=residual_size, vit_depth=vit_depth,
               qformer_config=BertConfig(num_hidden_layers=qformer_layers), use_mask=use_mask, mvm_decode=mvm_decode,
               qformer_text_input=qformer_text_input, tokenizer=tokenizer, llama_tokenizer=llama_tokenizer,
               vit_precision="fp16")
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float16)
    try:
        with torch.device(device):
            model = STLLMForCausalLM.from_config(cfg)
    finally:
        torch.set_default_dtype(old)
    sm = model.get_model().stllm_model
    if video_input == "residual":
        sm.up_proj.weight.data.normal_(0, 0.02)   # zero-init in the reference makes the global branch vanish
    return model.eval()


class VideoPrefillPipeline:
    """frames [B,T,3,H,W] (fp16, device) + prompt ids [B,P] -> logits [B,S,V]; optionally graph-captured."""

    _next_context = 0

    def __init__(self, model, batch, frames, prompt_len, img_size=224, use_graph=True):
        VideoPrefillPipeline._next_context += 1
        self.context = VideoPrefillPipeline._next_context   # private row-statistics buffers (ops.stats_context)
        self.model = model
        self.sm = model.get_model().stllm_model
        self.B, self.T, self.P = batch, frames, prompt_len
        dev = model.device
        self.frames = torch.zeros(batch, frames, 3, img_size, img_size, dtype=torch.float16, device=dev)
        self.prompt_ids = torch.zeros(batch, prompt_len, dtype=torch.long, device=dev)
        self.logits = None
        self.last_logits = None
        self.graph = None
        self.launches_per_step = 0
        self.use_graph = use_graph

    def _step(self):
        prev, ops.stats_context = ops.stats_context, self.context
        try:
            return self._step_impl()
        finally:
            ops.stats_context = prev

    def _step_impl(self):
        sm = self.sm
        emb, _, _ = sm.encode_img(self.frames)                       # [B,T,Q,D]
        vid = sm.aggregate_video(emb)                                # [B,1,L,D]
        B, _, L, D = vid.shape
        prompt = sm._embed(self.prompt_ids)                          # [B,P,D]
        seq = torch.empty(B, L + self.P, D, dtype=torch.float16, device=vid.device)
        seq[:, :L] = vid[:, 0]
        seq[:, L:] = prompt
        out = self.model(samples=None, inputs_embeds=seq)
        self.logits = out.logits
        self.last_logits = out.logits[:, -1].contiguous()
        return self.logits

    def warmup(self, n=2):
        for _ in range(n):
            c0 = ops.launch_count
            self._step()
            self.launches_per_step = ops.launch_count - c0
        torch.cuda.synchronize()

    def capture(self):
        self.warmup(2)
        if not self.use_graph:
            return
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            self._step()
        self.graph = g
        torch.cuda.synchronize()

    def run(self):
        """one step on the inputs currently in self.frames / self.prompt_ids (device-resident)"""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        return self.logits

    def __call__(self, frames_host, prompt_ids_host, out_host=None):
        """end-to-end call with HOST tensors (pinned for async copies): H2D, step, D2H of the last-position logits"""
        self.frames.copy_(frames_host, non_blocking=True)
        self.prompt_ids.copy_(prompt_ids_host, non_blocking=True)
        self.run()
        if out_host is None:
            out_host = torch.empty(self.last_logits.shape, dtype=self.last_logits.dtype, pin_memory=True)
        out_host.copy_(self.last_logits, non_blocking=True)
        return out_host
