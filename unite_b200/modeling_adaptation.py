"""Adaptation student (stages 1 and 3): ViT encoder on visible tokens + K linear CLIP-alignment decoders.

Drop-in surface of reference src/models/modeling_adaptation.py:
  AdaptationVisionTransformerEncoder :54, Linear_Decoder :182, AdaptationVisionTransformer :216,
  factories adaptation_umt_{base,large}_patch16_224 :337-378 — same kwargs, attributes
  (`model.encoder.patch_embed.{patch_size,num_patches,tubelet_size}`), `forward(x, mask, clip_only)` signature and
  returns, state_dict keys/shapes (pos_embed / clip_pos_embed are plain tensor attributes, absent from it).
The arithmetic runs in vit_core.ViTTrunk + the decoder kernels; autograd sees ONE Function per forward.
"""
from functools import partial
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .arena import ParamArena
from .modeling_finetune import Block, PatchEmbed, _ParamsOnly, drop_path_factors, get_sinusoid_encoding_table
from .registry import register_model
from .vit_core import DropPathSource, ViTTrunk

BF16, F32, I32 = torch.bfloat16, torch.float32, torch.int32


def _xavier_init(m):
    """modeling_adaptation.py:108-115."""
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class AdaptationVisionTransformerEncoder(_ParamsOnly):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0,
                 norm_layer=nn.LayerNorm, init_values=None, num_frames=16, tubelet_size=2, use_checkpoint=False,
                 checkpoint_num=0, use_learnable_pos_emb=False, clip_return_layers=[6, 7, 8, 9, 10, 11],
                 clip_student_return_interval=1, use_cls_token=False):
        super().__init__()
        if use_cls_token or use_learnable_pos_emb or num_classes:
            raise NotImplementedError("use_cls_token / use_learnable_pos_emb / encoder head are off in every shipped UNITE config")
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      num_frames=num_frames, tubelet_size=tubelet_size)
        # activation checkpointing is accepted for signature compatibility; 180 GB of HBM makes it unnecessary
        self.use_checkpoint, self.checkpoint_num = use_checkpoint, checkpoint_num
        self.return_index = list(clip_return_layers)
        self.use_learnable_pos_emb = use_learnable_pos_emb
        self.pos_embed = get_sinusoid_encoding_table(self.patch_embed.num_patches, embed_dim)
        self.drop_path_rates = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                  attn_drop=attn_drop_rate, drop_path=self.drop_path_rates[i], norm_layer=norm_layer, init_values=init_values)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        self.apply(_xavier_init)

    def get_num_layers(self):
        return len(self.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}


class Linear_Decoder(_ParamsOnly):
    def __init__(self, num_classes=768, embed_dim=768, norm_layer=nn.LayerNorm, clip_norm_type="l2"):
        super().__init__()
        if clip_norm_type != "l2":
            raise NotImplementedError("clip_norm_type must be 'l2' (the only value the shipped configs use)")
        self.clip_norm_type = clip_norm_type
        self.head = nn.Linear(embed_dim, num_classes)
        self.norm = norm_layer(num_classes)
        self.apply(_xavier_init)


def student_order_key(depth):
    """Arena order = backward completion order (see arena.py); q_bias directly before v_bias."""
    sub = ["norm1.weight", "norm1.bias", "attn.q_bias", "attn.v_bias", "attn.qkv.weight", "attn.proj.weight", "attn.proj.bias",
           "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"]

    def key(name):
        n = name[len("encoder."):] if name.startswith("encoder.") else name
        if n.startswith("clip_decoder.") or n.startswith("head.") or n.startswith("fc_norm."):
            return (0, 0, 0, n)
        if n.startswith("blocks."):
            _, l, rest = n.split(".", 2)
            return (1, depth - 1 - int(l), sub.index(rest) if rest in sub else 99, n)
        return (2, 0, 0, n)
    return key


def no_decay_rule(skip):
    """src/optim_factory.py:83-88."""
    def rule(name, p):
        return p.ndim == 1 or name.endswith(".bias") or name in skip
    return rule


class AdaptationVisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, encoder_in_chans=3, encoder_num_classes=0, encoder_embed_dim=768,
                 encoder_depth=12, encoder_num_heads=12, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.0, norm_layer=nn.LayerNorm, init_values=0.0, use_learnable_pos_emb=False,
                 use_cls_token=False, use_checkpoint=False, checkpoint_num=0, num_frames=16, tubelet_size=2,
                 clip_decoder_embed_dim=768, clip_output_dim=512, clip_norm_type="l2", clip_return_layers=[6, 7, 8, 9, 10, 11],
                 clip_student_return_interval=1):
        super().__init__()
        if clip_decoder_embed_dim != encoder_embed_dim:
            raise NotImplementedError("clip_decoder_embed_dim must equal encoder_embed_dim (as in the shipped configs)")
        self.encoder = AdaptationVisionTransformerEncoder(
            img_size=img_size, patch_size=patch_size, in_chans=encoder_in_chans, num_classes=encoder_num_classes,
            embed_dim=encoder_embed_dim, depth=encoder_depth, num_heads=encoder_num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
            qk_scale=qk_scale, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate,
            norm_layer=norm_layer, init_values=init_values, num_frames=num_frames, tubelet_size=tubelet_size,
            use_checkpoint=use_checkpoint, checkpoint_num=checkpoint_num, use_learnable_pos_emb=use_learnable_pos_emb,
            clip_return_layers=clip_return_layers, clip_student_return_interval=clip_student_return_interval,
            use_cls_token=use_cls_token)
        self.clip_decoder = nn.ModuleList([
            Linear_Decoder(num_classes=clip_output_dim, embed_dim=clip_decoder_embed_dim, norm_layer=norm_layer,
                           clip_norm_type=clip_norm_type) for _ in range(len(clip_return_layers))])
        self.clip_pos_embed = get_sinusoid_encoding_table(self.encoder.patch_embed.num_patches, clip_decoder_embed_dim)
        self.ln_eps = self.encoder.norm.eps
        self.clip_output_dim = clip_output_dim
        self.mlp_hidden = int(encoder_embed_dim * mlp_ratio)
        self.num_heads = encoder_num_heads
        self._core: Optional["AdaptationCore"] = None

    def get_num_layers(self):
        return len(self.encoder.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "mask_token", "clip_mask_token", "clip_pos_embed"}

    # ---- compute ---------------------------------------------------------------------------
    def core(self) -> "AdaptationCore":
        if self._core is None:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("unite_b200 models compute on CUDA only: call model.cuda() first (there is no CPU path)")
            self._core = AdaptationCore(self, dev)
        return self._core

    def _apply(self, fn, *a, **k):
        if self._core is not None:
            raise RuntimeError("the model's parameters already live in its device arena; move it before the first forward")
        return super()._apply(fn, *a, **k)

    def forward(self, x, mask, clip_only=False, vis_idx=None, patches=None, drop_path_factors_=None):
        """x [B,3,T,H,W] fp32, mask bool [B,N] (True = masked).  Returns x_clip [K,B,N_vis,C] if clip_only else
        (x_vis [B,N_vis,D], x_clip) — reference modeling_adaptation.py:304-334.

        Extra optional inputs (all derivable from the reference arguments): `vis_idx` int32 [B,N_vis] ascending
        visible-token indices (avoids the nonzero() host sync the boolean mask needs), `patches` bf16 im2col rows
        shared with the teacher, `drop_path_factors_` explicit DropPath draws [depth,2,B]."""
        core = self.core()
        if vis_idx is None:
            B = mask.shape[0]
            vis_idx = (~mask).nonzero()[:, 1].reshape(B, -1).to(I32)   # host sync, as in the reference's x[~mask]
        dp = drop_path_factors_
        if dp is None and self.training:
            dp = core.drop_path.draw(x.shape[0])
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if need_grad:
            anchor = torch.empty(0, device=x.device, requires_grad=True)
            outs = _StudentFn.apply(anchor, core, x, vis_idx, patches, dp, clip_only)
        else:
            outs = core.run_forward(x, vis_idx, patches, dp, clip_only, save=False)[:2]
        x_vis, x_clip = outs
        return x_clip if clip_only else (x_vis, x_clip)


class _StudentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, core, x, vis_idx, patches, dp, clip_only):
        x_vis, x_clip, state = core.run_forward(x, vis_idx, patches, dp, clip_only, save=True)
        ctx.core, ctx.state, ctx.clip_only = core, state, clip_only
        if clip_only:
            x_vis = x_clip.new_zeros(1)
            ctx.mark_non_differentiable(x_vis)
        return x_vis, x_clip

    @staticmethod
    def backward(ctx, g_vis, g_clip):
        ctx.core.run_backward(ctx.state, g_clip=g_clip, g_vis=None if ctx.clip_only else g_vis)
        return (None,) * 7


class AdaptationCore:
    """Owns the arena, the trunk and the decoder workspaces of one AdaptationVisionTransformer."""

    def __init__(self, model: AdaptationVisionTransformer, dev):
        enc = model.encoder
        self.model = model
        self.D, self.depth = enc.embed_dim, len(enc.blocks)
        self.C = model.clip_output_dim
        self.taps: List[int] = list(enc.return_index)
        self.N = enc.patch_embed.num_patches
        self.tubelet = enc.patch_embed.tubelet_size
        self.eps = model.ln_eps
        D = self.D
        self.arena = ParamArena(model, dev, student_order_key(self.depth), no_decay_rule(model.no_weight_decay()),
                                gap_after=lambda n, p: p.numel() if n.endswith("attn.q_bias") else 0)
        self.trunk = ViTTrunk(self.arena, "encoder.", D, self.depth, model.num_heads, model.mlp_hidden, self.eps)
        self.pos = enc.pos_embed[0].to(dev).contiguous()               # [N, D] fp32 (constant, uploaded ONCE)
        self.clip_pos = model.clip_pos_embed[0].to(dev).contiguous()
        self._shadow_version = None
        self._dec_ws: Dict = {}
        # The K decoder heads as ONE grouped GEMM (forward, dgrad, wgrad): possible when their weights sit back to back in the
        # arena (they do: the decay segment starts with clip_decoder.{k}.head.weight in order) and their biases at a fixed stride
        K = len(self.taps)
        offs = [self.arena.offsets[f"clip_decoder.{k}.head.weight"][0] for k in range(K)]
        boffs = [self.arena.offsets[f"clip_decoder.{k}.head.bias"][0] for k in range(K)]
        wsz = self.C * self.D
        import os
        self._group_dec = (os.environ.get("UB_GROUP_DECODERS", "1") == "1" and K > 1 and self.C % 256 == 0
                           and all(offs[k + 1] - offs[k] == wsz for k in range(K - 1))
                           and len({boffs[k + 1] - boffs[k] for k in range(K - 1)}) == 1 and boffs[1] > boffs[0])
        self._dec_w_off, self._dec_b_off = offs[0], boffs[0]
        self._dec_b_stride = (boffs[1] - boffs[0]) if K > 1 else 0
        import os
        self.drop_path = DropPathSource(enc.drop_path_rates, dev, seed=int(os.environ.get("UB_DROP_PATH_SEED", "0")))

    # bf16 shadow of the weights: refreshed whenever a parameter changed outside the fused optimizer
    def sync_shadow(self, force=False):
        v = self.arena.params_version()
        if force or v != self._shadow_version:
            ops.cast_bf16(self.arena.params, self.arena.w16)
            self._shadow_version = v

    def mark_shadow_fresh(self):
        self._shadow_version = self.arena.params_version()

    def _decoder_bufs(self, M, save):
        key = (M, save)
        if key not in self._dec_ws:
            dev, K = self.arena.device, len(self.taps)
            d = dict(z=torch.empty(K if save else 1, M, self.D, device=dev, dtype=BF16),
                     y=torch.empty(K if save else 1, M, self.C, device=dev, dtype=F32))
            if save:
                d["dy"] = torch.empty(K if self._grouped(M) else 1, M, self.C, device=dev, dtype=BF16)
                d["dz"] = torch.empty(K, M, self.D, device=dev, dtype=BF16)
            self._dec_ws[key] = d
        return self._dec_ws[key]

    def _grouped(self, M):
        """The decoder heads run as one grouped GEMM when every group's rows are whole 256-row tiles."""
        return self._group_dec and M % 256 == 0

    def run_forward(self, x, vis_idx, patches, dp, clip_only, save, targets=None, loss_acc=None, want_clip=True, abs_rows=None,
                    loss_clips=None):
        """Returns (x_vis or None, x_clip [K,B,Nv,C] fp32 or None, state).  With `targets` ([K,B,Nv,C] fp32) the decoder
        tail also accumulates the alignment loss mean(2 - 2<out,tgt>) into loss_acc (fp32 [1]).
        want_clip=False skips the alignment decoders (stage 3 only uses the encoder output, run_stage3.py:475-483);
        abs_rows int32 [B*Nv]: absolute rows of `patches` to gather (committee members share one clip's patches).
        loss_clips = (b_lo, b_hi): only these clips enter the loss (clip_loss_data 'source' / 'target', run_stage1.py:418-423)."""
        self.sync_shadow()
        B = x.shape[0]
        Nv = vis_idx.shape[1]
        M = B * Nv
        dev = x.device
        vis_flat = vis_idx.reshape(-1).contiguous()
        if patches is None:
            patches = torch.empty(B * self.N, 3 * self.tubelet * 256, device=dev, dtype=BF16)
            ops.patchify(x.contiguous(), patches, self.tubelet)
        p_vis = torch.empty(M, patches.shape[1], device=dev, dtype=BF16)
        if abs_rows is not None:
            ops.gather_rows(patches, abs_rows, p_vis)
        else:
            ops.gather_rows(patches, vis_flat, p_vis, rows_per_group=Nv, group_stride_rows=self.N)
        pos_vis = torch.empty(M, self.D, device=dev, dtype=F32)
        ops.gather_rows(self.pos, vis_flat, pos_vis)
        n_layers = (max(self.taps) + 1) if clip_only else self.depth
        K = len(self.taps)
        dws = self._decoder_bufs(M, save) if want_clip else None
        out = torch.empty(K, M, self.C, device=dev, dtype=F32) if want_clip else None
        a = self.arena
        enc_w, enc_b = a.p32("encoder.norm.weight"), a.p32("encoder.norm.bias")
        tap_of = {l: k for k, l in enumerate(self.taps)}
        loss_rows = None if loss_clips is None else (loss_clips[0] * Nv, loss_clips[1] * Nv)
        loss_scale = 1.0 / (K * (M if loss_rows is None else max(1, loss_rows[1] - loss_rows[0])))

        grouped = save and want_clip and self._grouped(M)

        def dec_tail(k, y):
            ops.dec_tail_fwd(y, a.p32(f"clip_decoder.{k}.norm.weight"), a.p32(f"clip_decoder.{k}.norm.bias"), self.eps, out[k],
                             None if targets is None else targets[k].reshape(M, self.C), loss_acc, loss_scale, loss_rows)

        def after_layer(l, x_l):
            if l not in tap_of or not want_clip:
                return
            k = tap_of[l]
            z = dws["z"][k if save else 0]
            y = dws["y"][k if save else 0]
            ops.layernorm_fwd(x_l, enc_w, enc_b, self.eps, z, post_add=self.clip_pos, post_idx=vis_flat)
            if grouped:
                return                                   # the K head GEMMs run as one grouped launch after the trunk
            ops.gemm(z, a.b16(f"clip_decoder.{k}.head.weight"), y, bias=a.p32(f"clip_decoder.{k}.head.bias"))
            dec_tail(k, y)

        ws = self.trunk.forward(p_vis, pos_vis, B, Nv, n_layers, save, dp, after_layer=after_layer)
        if grouped:
            # y_k = z_k W_k^T + b_k for all K decoders (modeling_adaptation.py:203-213, 322-325) in one launch: 480 pair tiles
            # instead of 6 launches of 80 (1.08 waves each on 74 CTA pairs)
            W = a.w16[self._dec_w_off:self._dec_w_off + K * self.C * self.D].view(K * self.C, self.D)
            bias = a.params[self._dec_b_off:self._dec_b_off + (K - 1) * self._dec_b_stride + self.C]
            ops.gemm(dws["z"].view(K * M, self.D), W, dws["y"].view(K * M, self.C), bias=bias,
                     group=dict(rows=M, K=self.D, b_n=self.C, bias=self._dec_b_stride))
            for k in range(K):
                dec_tail(k, dws["y"][k])
        x_vis = None
        if not clip_only:
            x_vis = torch.empty(M, self.D, device=dev, dtype=F32)
            ops.layernorm_fwd(ws.x_at(self.depth), enc_w, enc_b, self.eps, x_vis)
            x_vis = x_vis.view(B, Nv, self.D)
        state = dict(ws=ws, dws=dws, B=B, Nv=Nv, M=M, vis_flat=vis_flat, clip_only=clip_only, loss_rows=loss_rows) if save else None
        return x_vis, (out.view(K, B, Nv, self.C) if want_clip else None), state

    def block_grad_hi(self, l):
        """End of the decay-segment prefix that is final once block l's backward has run (decoders + blocks >= l)."""
        return self.arena.range_of([f"encoder.blocks.{l}.mlp.fc2.weight", f"encoder.blocks.{l}.attn.qkv.weight"])[1]

    def run_backward(self, state, g_clip=None, g_vis=None, targets=None, grad_sync=None):
        """Gradient of (sum g_clip*x_clip + sum g_vis*x_vis), or — engine fast path — of the alignment loss
        mean(2 - 2<x_clip, targets>) when `targets` is given.  Parameter gradients are ACCUMULATED in the arena."""
        a = self.arena
        a.attach_grads()
        ws, dws, M = state["ws"], state["dws"], state["M"]
        K = len(self.taps)
        enc_w = a.p32("encoder.norm.weight")
        g_enc_w, g_enc_b = a.g32("encoder.norm.weight"), a.g32("encoder.norm.bias")
        have_clip = targets is not None or g_clip is not None
        if have_clip:
            go = (targets if targets is not None else g_clip).reshape(K, M, self.C)
            if not go.is_contiguous() or go.dtype != F32:
                go = go.contiguous().float()
            go_rows = state.get("loss_rows") if targets is not None else None
            go_scale = -2.0 / (K * (M if go_rows is None else max(1, go_rows[1] - go_rows[0]))) if targets is not None else 1.0
            grouped = dws["dy"].shape[0] == K and K > 1
            for k in range(K):
                dy = dws["dy"][k if grouped else 0]
                ops.dec_tail_bwd(dws["y"][k], a.p32(f"clip_decoder.{k}.norm.weight"), a.p32(f"clip_decoder.{k}.norm.bias"), self.eps,
                                 go[k], go_scale, dy, a.g32(f"clip_decoder.{k}.norm.weight"), a.g32(f"clip_decoder.{k}.norm.bias"),
                                 go_rows)
                ops.colsum_bf16(dy, a.g32(f"clip_decoder.{k}.head.bias"))
                if not grouped:
                    self.trunk._wgrad(dy, dws["z"][k], a.g32(f"clip_decoder.{k}.head.weight"))
                    ops.gemm(dy, a.b16(f"clip_decoder.{k}.head.weight"), dws["dz"][k], b_t=True)
            if grouped:
                # the K weight gradients dW_k += dy_k^T z_k and the K input gradients dz_k = dy_k W_k as two grouped launches
                C_, D_ = self.C, self.D
                n_w = K * C_ * D_
                gW = a.grads[self._dec_w_off:self._dec_w_off + n_w].view(K * C_, D_)
                W = a.w16[self._dec_w_off:self._dec_w_off + n_w].view(K * C_, D_)
                dy_all = dws["dy"].view(K * M, C_)
                from .vit_core import _splits_for
                ops.gemm(dy_all, dws["z"].view(K * M, D_), gW, a_t=True, b_t=True, accumulate=True,
                         split_k=_splits_for(K * C_, D_, self.trunk.sms), group=dict(rows=C_, K=M, a_k=M, a_m=C_, b_k=M))
                ops.gemm(dy_all, W, dws["dz"].view(K * M, D_), b_t=True, group=dict(rows=M, K=C_, b_k=C_))
        taps = {}
        N = state["Nv"]

        def make_tap(l, grads):
            def tap(dx_in, dxs_out, row_scale, dsum):
                for i, dz in enumerate(grads):
                    last = i == len(grads) - 1
                    ops.layernorm_bwd(dz, ws.x_at(l + 1), enc_w, self.eps, dx_in, ws.dx, dxs_out if last else None,
                                      row_scale, N, g_enc_w, g_enc_b, dsum=dsum if last else None)
                    dx_in = ws.dx
            return tap

        per_layer: Dict[int, list] = {}
        if have_clip:
            for k, l in enumerate(self.taps):
                if l < ws.n_layers:
                    per_layer.setdefault(l, []).append(dws["dz"][k])
        if g_vis is not None:
            gv = torch.empty(M, self.D, device=a.device, dtype=BF16)
            ops.cast_scale_bf16(g_vis.reshape(M, self.D).contiguous().float(), gv)
            per_layer.setdefault(self.depth - 1, []).append(gv)
        for l, grads in per_layer.items():
            taps[l] = make_tap(l, grads)
        if (ws.n_layers - 1) not in taps:
            raise RuntimeError("no gradient reaches the last computed block (neither x_clip nor x_vis was used)")
        on_done = None
        if grad_sync is not None and grad_sync.world > 1:
            on_done = lambda l: grad_sync.range_ready(a.grads, self.block_grad_hi(l))
        self.trunk.backward(ws, taps, on_block_done=on_done)
        ws.busy = False


@register_model
def adaptation_umt_base_patch16_224(pretrained=False, **kwargs):
    if pretrained:
        raise NotImplementedError("pretrained weights need network access; load a state_dict instead")
    return AdaptationVisionTransformer(img_size=224, patch_size=16, encoder_embed_dim=768, encoder_depth=12, encoder_num_heads=12,
                                       encoder_num_classes=0, mlp_ratio=4, qkv_bias=True,
                                       norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


@register_model
def adaptation_umt_large_patch16_224(pretrained=False, **kwargs):
    if pretrained:
        raise NotImplementedError("pretrained weights need network access; load a state_dict instead")
    return AdaptationVisionTransformer(img_size=224, patch_size=16, encoder_embed_dim=1024, encoder_depth=24, encoder_num_heads=16,
                                       encoder_num_classes=0, mlp_ratio=4, qkv_bias=True,
                                       norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
