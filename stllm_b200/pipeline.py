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

    def __init__(self, model, batch, frames, prompt_len, img_size=224, use_graph=True, frames_dtype=torch.float16):
        """frames_dtype=torch.float32: the device-resident frames are fp32 (synthetic fp32 frames, cast to fp16 on
        the device) -- the cast happens inside the patch-embed im2col kernel, no extra pass"""
        VideoPrefillPipeline._next_context += 1
        self.context = VideoPrefillPipeline._next_context   # private row-statistics buffers (ops.stats_context)
        self.model = model
        self.sm = model.get_model().stllm_model
        self.B, self.T, self.P = batch, frames, prompt_len
        dev = model.device
        self.frames = torch.zeros(batch, frames, 3, img_size, img_size, dtype=frames_dtype, device=dev)
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


class _GraphedStep:
    """shared capture / replay plumbing of the pipelines below: `_step()` is the device work of one step"""

    def _init_graph(self, use_graph):
        _GraphedStep._ctx = getattr(_GraphedStep, "_ctx", 1000) + 1
        self.context = _GraphedStep._ctx
        self.use_graph, self.graph, self.launches_per_step = use_graph, None, 0

    def _run_step(self):
        prev, ops.stats_context = ops.stats_context, self.context
        try:
            return self._step()
        finally:
            ops.stats_context = prev

    def capture(self):
        for _ in range(2):
            c0 = ops.launch_count
            self._run_step()
            self.launches_per_step = ops.launch_count - c0
        torch.cuda.synchronize()
        if not self.use_graph:
            return
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._run_step()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            self._run_step()
        self.graph = g
        torch.cuda.synchronize()

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._run_step()


class MaskedTrainPipeline(_GraphedStep):
    """BASELINE config 3 -- the training-branch forward `STLLMForCausalLM.forward(samples)` (st_llm.py:116-146, 56-92,
    447-546) with the dynamic token mask and the MVM head, as ONE captured step: encode_img -> aggregation -> masked and
    unmasked sequences (one two-source gather each, rows planned on the host once) -> masked LLM pass -> MVM decoder ->
    unmasked LLM pass -> visible-position gather -> cosine loss; lm_head on the masked pass -> shifted CE; loss = CE + MVM.
    The host half (tokenisation, mask draw from numpy's RNG, row plan) runs once at construction: it depends on the text
    and the mask only, not on the frames."""

    def __init__(self, model, batch, frames, after_ids, answer_ids, mask_rate=0.5, seed=1237, img_size=224, use_graph=True,
                 frames_dtype=torch.float32, before_ids=None):
        import numpy as np
        from . import prompt_assembly as PA
        from .st_llm import RandomMaskingGenerator
        self._init_graph(use_graph)
        self.model, self.llm, self.sm = model, model.get_model(), model.get_model().stllm_model
        sm = self.sm
        dev = model.device
        self.B, self.T = batch, frames
        self.frames = torch.zeros(batch, frames, 3, img_size, img_size, dtype=frames_dtype, device=dev)
        Q = sm.query_tokens.shape[1] if sm.has_qformer else 64
        L = frames * Q if sm.video_input == "all" else (Q if sm.video_input == "mean" else sm.residual_size * Q)
        np.random.seed(seed)
        keep = ~RandomMaskingGenerator(L, mask_rate, batch, "cpu").numpy()               # utils.py:4-16, numpy RNG order
        self.img_len, self.mask_img_len = L, int(keep[0].sum())
        sm.img_len, sm.mask_img_len, sm._keep_host = L, self.mask_img_len, keep
        sm.mask = torch.from_numpy(~keep).unsqueeze(1).to(dev)
        after = [np.asarray(a, dtype=np.int64) for a in after_ids]
        before = [np.zeros(0, dtype=np.int64)] * batch if before_ids is None else \
            [np.asarray(b, dtype=np.int64) for b in before_ids]
        a_ids = np.asarray(answer_ids, dtype=np.int64)
        a_att = np.ones_like(a_ids)
        pad = model.config.pad_token_id if model.config.pad_token_id is not None else 0
        bos = None if sm.qformer_text_input else (model.config.bos_token_id or 1)
        pm = PA.plan_sequences(PA.video_row_codes(batch, L, keep), before, after, a_ids, a_att, pad, bos)
        pu = PA.plan_sequences(PA.video_row_codes(batch, L), before, after, a_ids, a_att, pad, bos)
        up = lambda a: torch.from_numpy(a).to(dev)
        self.rows_m, self.att_m, self.rows_u, self.att_u = up(pm.rows), up(pm.attention_mask), up(pu.rows), up(pu.attention_mask)
        self.targets = up(pm.targets)
        self.sel_rows = self.llm.mvm_select_rows(keep, pu.rows.shape[1]).to(dev)
        self.seq_masked, self.seq_unmasked = pm.rows.shape[1], pu.rows.shape[1]
        self.loss = None
        self.logits = None

    def _step(self):
        sm, llm = self.sm, self.llm
        emb, _, _ = sm.encode_img(self.frames)
        vid = sm.aggregate_video(emb)
        B, _, L, D = vid.shape
        v2d = vid.reshape(B * L, D)
        emb_m = sm._gather_plan(self.rows_m, v2d)
        emb_u = sm._gather_plan(self.rows_u, v2d)
        out, loss_mvm = llm.mvm_passes(emb_m, self.att_m, emb_u, self.att_u, self.sel_rows, self.mask_img_len)
        self.logits = self.model._lm_head(out.last_hidden_state)
        self.loss = self.model._ce(self.logits, self.targets) + loss_mvm
        return self.loss


class ShardedVideoPipeline(_GraphedStep):
    """One clip whose T frames are sharded over the ranks (BASELINE config 4: 64 frames, global-local `residual` merge; also the
    strong-scaling row of config 2 with video_input='all'): every rank encodes its block of T / W frames (captured graph A),
    ONE all-gather of the per-frame tokens [T / W, 32, 4096] fp16 per rank (eager NCCL call between the two graphs), then
    aggregation + [video | prompt] + prefill replicated on every rank (captured graph B).  W = 1: the same two graphs, no
    collective -- the single-GPU result the sharded one must equal bit for bit."""

    def __init__(self, model, T, prompt_len, world=1, rank=0, group=None, img_size=224, use_graph=True,
                 frames_dtype=torch.float32):
        assert T % world == 0, "frames must divide evenly over the ranks"
        self._init_graph(use_graph)
        self.model, self.sm = model, model.get_model().stllm_model
        self.T, self.world, self.rank, self.group, self.P = T, world, rank, group, prompt_len
        dev = model.device
        self.t_loc = T // world
        D = model.config.hidden_size
        Q = self.sm.query_tokens.shape[1]
        self.frames = torch.zeros(self.t_loc, 3, img_size, img_size, dtype=frames_dtype, device=dev)
        self.prompt_ids = torch.zeros(1, prompt_len, dtype=torch.long, device=dev)
        self.tok_local = torch.zeros(self.t_loc, Q, D, dtype=torch.float16, device=dev)
        self.tok_all = torch.zeros(T, Q, D, dtype=torch.float16, device=dev)
        self.graph_b = None
        self.logits = None
        self.last_logits = None
        self.gather_bytes = 0 if world == 1 else self.tok_all.numel() * 2

    def _step(self):           # graph A: the local frames
        tok, _, _ = self.sm.encode_img(self.frames)
        self.tok_local.copy_(tok)

    def _step_b(self):         # graph B: merge + prefill, replicated
        sm = self.sm
        vid = sm.aggregate_video(self.tok_all.unsqueeze(0))
        prompt = sm._embed(self.prompt_ids)
        L = vid.shape[2]
        seq = torch.empty(1, L + self.P, vid.shape[-1], dtype=torch.float16, device=vid.device)
        seq[:, :L] = vid[:, 0]
        seq[:, L:] = prompt
        self.logits = self.model(samples=None, inputs_embeds=seq).logits
        self.last_logits = self.logits[:, -1].contiguous()

    def capture(self):
        import torch.distributed as dist
        super().capture()
        self._exchange()
        prev, ops.stats_context = ops.stats_context, self.context
        try:
            for _ in range(2):
                c0 = ops.launch_count
                self._step_b()
                self.launches_per_step_b = ops.launch_count - c0
            torch.cuda.synchronize()
            if self.use_graph:
                g = torch.cuda.CUDAGraph()
                st = torch.cuda.Stream()
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    self._step_b()
                torch.cuda.current_stream().wait_stream(st)
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._step_b()
                self.graph_b = g
                torch.cuda.synchronize()
        finally:
            ops.stats_context = prev

    def _exchange(self):
        import torch.distributed as dist
        if self.world == 1:
            self.tok_all.copy_(self.tok_local)
        else:
            dist.all_gather_into_tensor(self.tok_all, self.tok_local, group=self.group)   # rank r's block = frames r T/W ..

    def run(self):
        super().run()
        self._exchange()
        if self.graph_b is not None:
            self.graph_b.replay()
        else:
            prev, ops.stats_context = ops.stats_context, self.context
            try:
                self._step_b()
            finally:
                ops.stats_context = prev
        return self.logits
