"""The operator surface: ``model_fn_wan_video`` (one velocity prediction), the denoising loop and ``install()``.

Mirrors ``diffsynth/pipelines/wan_video_new.py:1260-1468`` (model_fn_wan_video) and ``:515-542`` (denoise loop),
``diffsynth/schedulers/flow_match.py`` (FlowMatchScheduler) -- same keyword names, argument meaning and error
behaviour (Python exceptions) -- with the block math on libwvd.so.  ``install(pipe)`` plugs this into a reference
``WanVideoPipeline`` by replacing the instance attribute ``pipe.model_fn`` (set at wan_video_new.py:78), the
same monkey-patch idiom as the reference's ``enable_usp`` (:326-338).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import engine, ops as _cuda_ops
from .ulysses import make_exchange, shard_bounds

Tensor = torch.Tensor

_rope_tables = {}


def _rope_info(dit, f, h, w, device, rope_indices, token_offset, ops):
    if hasattr(dit, "rope_info"):
        info = dit.rope_info(f, h, w, device, rope_indices, token_offset) if ops is _cuda_ops else None
        if info is not None:
            return info
    key = (id(dit), str(device))
    tab = _rope_tables.get(key)
    if tab is None:
        tab = ops.make_rope_table(dit.freqs, device)
        _rope_tables[key] = tab
    fid = None
    if rope_indices is not None:
        fid = torch.as_tensor(rope_indices, dtype=torch.int32).to(device).contiguous()
        f = int(fid.numel())
    return engine.RopeInfo(tab, (f, h, w), token_offset, fid)


def model_fn_wan_video(
    dit,
    motion_controller=None,
    vace=None,
    animate_adapter=None,
    latents: Tensor = None,
    timestep: Tensor = None,
    context: Tensor = None,
    clip_feature: Optional[Tensor] = None,
    y: Optional[Tensor] = None,
    reference_latents=None,
    vace_context=None,
    vace_scale=1.0,
    audio_embeds: Optional[Tensor] = None,
    motion_latents: Optional[Tensor] = None,
    s2v_pose_latents: Optional[Tensor] = None,
    drop_motion_frames: bool = True,
    tea_cache=None,
    use_unified_sequence_parallel: bool = False,
    motion_bucket_id: Optional[Tensor] = None,
    pose_latents=None,
    face_pixel_values=None,
    sliding_window_size: Optional[int] = None,
    sliding_window_stride: Optional[int] = None,
    cfg_merge: bool = False,
    use_gradient_checkpointing: bool = False,
    use_gradient_checkpointing_offload: bool = False,
    control_camera_latents_input=None,
    fuse_vae_embedding_in_latents: bool = False,
    rope_indices: Optional[Tensor] = None,
    sp_group=None,
    text_cache: Optional[engine.TextCache] = None,
    ops=_cuda_ops,
    **kwargs,
) -> Tensor:
    """One velocity prediction (B,16,F,H,W) -> (B,16,F,H,W); keyword-compatible with the reference model_fn.

    Extra keywords: ``rope_indices`` (the fixed WanModel.forward variant, wan_video_dit.py:378-384), ``sp_group``
    (torch.distributed group for Ulysses; default group when ``use_unified_sequence_parallel``), ``text_cache`` (an
    engine.TextCache: text embedding and cross-attention K / V computed once per prompt instead of once per call --
    identical outputs; off unless given), ``ops`` (kernel backend; tests inject a CPU restatement to exercise the host
    logic -- the product default is libwvd.so)."""
    for name, val in (("audio_embeds", audio_embeds), ("clip_feature", clip_feature), ("y", y),
                      ("reference_latents", reference_latents), ("pose_latents", pose_latents),
                      ("face_pixel_values", face_pixel_values), ("control_camera_latents_input", control_camera_latents_input),
                      ("sliding_window_size", sliding_window_size)):
        if val is not None:
            raise NotImplementedError(f"model_fn_wan_video({name}=...) is outside the T2V / VACE hot path of this build")
    if getattr(dit, "seperated_timestep", False) and fuse_vae_embedding_in_latents:
        raise NotImplementedError("seperated_timestep (Wan2.2 TI2V) is outside the T2V / VACE hot path of this build")
    if torch.is_grad_enabled() and any(p.requires_grad for p in dit.parameters()):
        raise NotImplementedError("the wvd path is forward-only: call under torch.no_grad()")

    # ---- timestep / text embeddings: tiny library GEMMs, kept in PyTorch (wan_video_new.py:1351-1357) ----
    from .wan_video_dit import sinusoidal_embedding_1d
    t = dit.time_embedding(sinusoidal_embedding_1d(dit.freq_dim, timestep))
    t_mod = dit.time_projection(t).unflatten(1, (6, dit.dim))
    if motion_bucket_id is not None and motion_controller is not None:
        t_mod = t_mod + motion_controller(motion_bucket_id).unflatten(1, (6, dit.dim))
    text_entry = None
    if text_cache is not None:
        text_entry = text_cache.entry(context)
        te = dit.text_embedding
        stamp = engine.TextCache.stamp(*[p_ for p_ in te.parameters()])
        if text_entry["emb"] is None or text_entry["emb_stamp"] != stamp:
            text_entry["emb"], text_entry["emb_stamp"] = te(context), stamp
            text_entry["kv"].clear()
            text_cache.misses += 1
        else:
            text_cache.hits += 1
        context = text_entry["emb"]
    else:
        context = dit.text_embedding(context)

    x = latents
    if x.shape[0] != context.shape[0]:                       # merged CFG (:1361-1364)
        x = torch.concat([x] * context.shape[0], dim=0)
    if t_mod.shape[0] != context.shape[0]:
        t_mod = torch.concat([t_mod] * context.shape[0], dim=0)
        t = torch.concat([t] * context.shape[0], dim=0)

    # patchify: Conv3d k=s=(1,2,2) (:1374) + 'b c f h w -> b (f h w) c' (:1381-1382), as im2col + GEMM
    pt, ph, pw = engine._unwrap(dit.patch_embedding).kernel_size
    bsz, f, h, w = x.shape[0], x.shape[2] // pt, x.shape[3] // ph, x.shape[4] // pw
    x = torch.stack([engine.patch_embed(dit.patch_embedding, x[i:i + 1], ops) for i in range(bsz)], dim=0)
    n_tokens = x.shape[1]

    # ---- sequence parallel plan (Ulysses; replaces xfuser USP, wan_video_new.py:1412-1417) ----
    exchange = engine._LOCAL
    lo, hi, n_loc = 0, n_tokens, n_tokens
    if use_unified_sequence_parallel:
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size(sp_group) > 1:
            heads = dit.blocks[0].self_attn.num_heads
            if heads % dist.get_world_size(sp_group) != 0:       # before any launch or collective: every rank raises alike
                raise ValueError(f"Ulysses needs num_heads ({heads}) divisible by the sequence-parallel world size "
                                 f"({dist.get_world_size(sp_group)})")
            exchange = make_exchange(sp_group, n_tokens, x.device)
            lo, hi, n_loc = shard_bounds(n_tokens, exchange.world, exchange.rank)

    tea_cache_update = tea_cache.check(dit, x, t_mod) if tea_cache is not None else False

    outs = []
    for b in range(bsz):
        xb_full = x[b]
        ctx = context[b]
        tm, tb = t_mod[b:b + 1], t[b:b + 1]
        rope = _rope_info(dit, f, h, w, x.device, rope_indices, lo, ops)
        ffn = _ffn_dim(dit)
        ws = engine.workspace(n_loc, dit.dim, ffn, ctx.shape[0], x.dtype, x.device)
        xb = ws.get("x", (n_loc, dit.dim))
        xb[:hi - lo].copy_(xb_full[lo:hi])
        if hi - lo < n_loc:
            xb[hi - lo:].zero_()                              # zero-pad the last shard (:1414-1416)
        hints = None
        if vace_context is not None:
            vc = vace_context[b:b + 1] if vace_context.shape[0] == bsz else vace_context
            hints = engine.vace_forward(vace, xb, vc, ctx, tm, rope, ws, ops, exchange,
                                        token_slice=slice(lo, hi) if exchange.world > 1 else None,
                                        text_entry=_sample_entry(text_entry, b, bsz))
        if tea_cache_update:
            xb = ops.as_2d(tea_cache.update(xb.unsqueeze(0)))
        else:
            for block_id, block in enumerate(dit.blocks):
                engine.dit_block_forward(block, xb, ctx, tm, rope, ws, ops, exchange, _sample_entry(text_entry, b, bsz))
                if hints is not None and block_id in vace.vace_layers_mapping:
                    ops.scale_add(xb, hints[vace.vace_layers_mapping[block_id]], float(vace_scale), out=xb)
            if tea_cache is not None:
                tea_cache.store(xb.unsqueeze(0))
        y_loc = engine.head_forward(dit.head, xb, tb, ws, ops)            # (n_loc, out_dim*prod(patch))
        if exchange.world > 1:
            y_all = exchange.all_gather_tokens(y_loc)[:n_tokens]          # :1459-1462
        else:
            y_all = y_loc
        outs.append(y_all.clone())
    out = torch.stack(outs, dim=0)
    return dit.unpatchify(out, (f, h, w))


def _sample_entry(text_entry, b: int, bsz: int):
    """Per-sample view of a TextCache entry: a merged-CFG call (batch 2: posi | nega contexts in one tensor) keeps one
    K / V set per sample."""
    if text_entry is None:
        return None
    if bsz == 1:
        return text_entry
    sub = text_entry.setdefault("samples", {})
    return sub.setdefault(b, dict(kv={}))


def _ffn_dim(dit) -> int:
    blk = dit.blocks[0]
    return int(getattr(blk, "ffn_dim", None) or engine._unwrap(blk.ffn[0]).weight.shape[0])


class FlowMatchScheduler:
    """Flow-matching Euler scheduler as WanVideoPipeline builds it (wan_video_new.py:39:
    FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True); flow_match.py:34-100), with ``step`` kept on
    the device: the sigma pair is looked up on the host from the step index, so there is no per-step ``.cpu()`` sync."""

    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0,
                 sigma_min=0.003 / 1.002, extra_one_step=False):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.sigma_max, self.sigma_min, self.extra_one_step = sigma_max, sigma_min, extra_one_step
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, shift=None, **_):
        if shift is not None:
            self.shift = shift
        start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            s = torch.linspace(start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            s = torch.linspace(start, self.sigma_min, num_inference_steps)
        self.sigmas = self.shift * s / (1 + (self.shift - 1) * s)
        self.timesteps = self.sigmas * self.num_train_timesteps

    def _index(self, timestep) -> int:
        ts = torch.as_tensor(timestep).detach().float().cpu()
        return int(torch.argmin((self.timesteps - ts).abs()))

    def dsigma(self, timestep, to_final: bool = False) -> float:
        """sigma_next - sigma of the step at ``timestep``, evaluated in fp32 like flow_match.py:76-81 does."""
        i = self._index(timestep)
        nxt = torch.zeros((), dtype=self.sigmas.dtype) if (to_final or i + 1 >= len(self.timesteps)) else self.sigmas[i + 1]
        return float(nxt - self.sigmas[i])

    def step(self, model_output: Tensor, timestep, sample: Tensor, to_final: bool = False, **_) -> Tensor:
        return sample + model_output * self.dsigma(timestep, to_final)

    def add_noise(self, original_samples: Tensor, noise: Tensor, timestep) -> Tensor:
        sigma = float(self.sigmas[self._index(timestep)])
        return (1 - sigma) * original_samples + sigma * noise


class GraphedModelFn:
    """One velocity prediction as a replayed CUDA graph: the ~900 kernel launches of a call (C-ABI kernels and the few
    PyTorch ones alike) are captured once per (shapes, prompt, flags) and replayed with no host work in between.
    Inputs are copied into static buffers; the output buffer is reused by the next replay (clone it to keep it).
    The TextCache, if any, is filled by the warm-up call, so the captured graph reads the cached K / V.
    What the graph buys is launch latency: nothing at c3 on one GPU (1.5 s of long kernels), the gaps between the short
    kernels of a 3,705-token rank at 8 GPUs or of the 1,280-token c1 model."""

    def __init__(self, **fixed):
        self.fixed = fixed            # dit, vace, vace_scale, use_unified_sequence_parallel, text_cache ...
        self._graphs = {}

    def __call__(self, latents: Tensor, timestep: Tensor, context: Tensor, vace_context: Optional[Tensor] = None) -> Tensor:
        key = (tuple(latents.shape), latents.dtype, id(context), context._version,
               None if vace_context is None else tuple(vace_context.shape))
        g = self._graphs.get(key)
        if g is None:
            st = dict(latents=latents.clone(), timestep=timestep.clone(), context=context,
                      vace_context=None if vace_context is None else vace_context.clone())
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                       # warm-up off the capture: workspaces, tensor maps, caches
                for _ in range(2):
                    model_fn_wan_video(**self.fixed, **st)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = model_fn_wan_video(**self.fixed, **st)
            g = self._graphs[key] = (graph, st, out)
        graph, st, out = g
        st["latents"].copy_(latents)
        st["timestep"].copy_(timestep)
        if vace_context is not None and vace_context.data_ptr() != st["vace_context"].data_ptr():
            st["vace_context"].copy_(vace_context)
        graph.replay()
        return out


@torch.no_grad()
def denoise(dit, vace, latents: Tensor, context_posi: Tensor, context_nega: Optional[Tensor] = None,
            vace_context: Optional[Tensor] = None, vace_scale: float = 1.0, num_inference_steps: int = 50,
            cfg_scale: float = 5.0, sigma_shift: float = 5.0, torch_dtype=torch.bfloat16, progress=None,
            use_unified_sequence_parallel: bool = False, scheduler: Optional[FlowMatchScheduler] = None,
            cache_text: bool = True, use_cuda_graph: bool = False, fused_step: bool = True) -> Tensor:
    """The denoising loop of WanVideoPipeline.__call__ (wan_video_new.py:515-542): per step the timestep is rounded
    to the model dtype (:526), posi (+ nega) velocity, CFG combine (:535), Euler step (:540).

    Loop-level fusion on top of the per-call kernels, all switchable and all bit-identical to the plain loop:
    cache_text      text embedding + cross-attention K / V once per prompt (engine.TextCache) instead of 100x per video
    fused_step      CFG combine + Euler update as ONE kernel on the device (wvd_cfg_euler_step); sigma pairs are looked
                    up on the host from the step index, so there is no per-step .cpu() sync (flow_match.py:73-75)
    use_cuda_graph  each velocity prediction replayed as a CUDA graph (GraphedModelFn)"""
    sch = scheduler or FlowMatchScheduler(shift=5, sigma_min=0.0, extra_one_step=True)
    sch.set_timesteps(num_inference_steps, shift=sigma_shift)
    it = sch.timesteps if progress is None else progress(sch.timesteps)
    fixed = dict(dit=dit, vace=vace, vace_scale=vace_scale, use_unified_sequence_parallel=use_unified_sequence_parallel)
    if cache_text:
        fixed["text_cache"] = engine.TextCache()
    on_gpu = latents.is_cuda
    call = GraphedModelFn(**fixed) if (use_cuda_graph and on_gpu) else (
        lambda latents, timestep, context, vace_context=None: model_fn_wan_video(latents=latents, timestep=timestep, context=context,
                                                                               vace_context=vace_context, **fixed))
    for i, ts in enumerate(it):
        timestep = ts.unsqueeze(0).to(dtype=torch_dtype, device=latents.device)
        v = call(latents, timestep, context_posi, vace_context)
        vn = None
        if cfg_scale != 1.0:
            if use_cuda_graph and on_gpu:
                v = v.clone()                                  # the graph's output buffer is shared between replays of one graph only,
            vn = call(latents, timestep, context_nega, vace_context)      # but posi / nega are different graphs: the clone is for safety
        if fused_step and on_gpu and latents.is_contiguous():
            latents = _cuda_ops.cfg_euler_step(latents, v.contiguous(), None if vn is None else vn.contiguous(), cfg_scale,
                                               sch.dsigma(sch.timesteps[i]))
        else:
            if vn is not None:
                v = vn + cfg_scale * (v - vn)
            latents = sch.step(v, sch.timesteps[i], latents)
    return latents


def install(pipe, use_usp: bool = False, cache_text: bool = True):
    """Plug the B200 path into a reference ``WanVideoPipeline`` (diffsynth/pipelines/wan_video_new.py):

        pipe = WanVideoPipeline.from_pretrained(...); pipe.load_lora(pipe.vace, ...); pipe.enable_vram_management()
        video_styler_b200.install(pipe)         # <- one line; infer_ditto.py is otherwise unchanged

    ``pipe.model_fn`` is the instance attribute every caller goes through (:529,534 denoise loop; :117 training_loss).
    The module tree, state-dict keys, merged LoRA weights and vram wrappers are left untouched."""
    import functools
    extra = {}
    if use_usp:
        extra["use_unified_sequence_parallel"] = True
    if cache_text:
        # the pipeline passes the same posi / nega context tensors at every step (wan_video_new.py:529,534): their text
        # embedding and cross-attention K / V are computed once per prompt instead of 100x per video
        pipe.wvd_text_cache = extra["text_cache"] = engine.TextCache()
    fn = functools.partial(model_fn_wan_video, **extra) if extra else model_fn_wan_video
    pipe.model_fn = fn
    pipe.use_unified_sequence_parallel = bool(use_usp) or getattr(pipe, "use_unified_sequence_parallel", False)
    return pipe
