"""CPU oracle for the UNITE training-step hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch restatement (functional, state_dict-driven, runs on CPU) of the algorithm the
reference (reddyav1/unite, mounted at /root/reference in the build container) executes on this path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import it;
the product package `unite_b200` never does.

Every function cites the reference file:line it restates.  The restatement is PINNED against the
reference's own nn.Modules: oracle/make_golden.py imports /root/reference/src/models (4-symbol timm shim),
checks this file against them on identical state_dicts and inputs, and writes tests/golden/*.pt; the
`-m "not gpu"` tests re-check this file against those fixtures.  The reference itself ships no tests or
golden vectors, so that self-generated pin is the only one available.

Third-party arithmetic outside /root/reference that the path relies on: torch (reference pins 1.13.0,
environment.yaml:122; here 2.11) for multinomial / MultiheadAttention / LayerNorm / GELU, timm 0.4.12
(environment.yaml:325) for drop_path, OpenAI CLIP (unpinned git dependency, environment.yaml:353) for the
stage-3 zero-shot head, which is absent: `clip_zero_shot` takes image features and a text matrix instead.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import math
import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------
@dataclass
class StudentCfg:
    """adaptation_umt_{base,large}_patch16_224 (modeling_adaptation.py:337-378) / vit_* (modeling_finetune.py:386-415)."""
    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    mlp_ratio: float = 4.0
    img_size: int = 224
    patch_size: int = 16
    num_frames: int = 8
    tubelet_size: int = 1
    ln_eps: float = 1e-6           # norm_layer=partial(nn.LayerNorm, eps=1e-6)
    return_layers: Sequence[int] = (6, 7, 8, 9, 10, 11)
    clip_output_dim: int = 512
    num_classes: int = 12          # stage-2 head

    @property
    def grid(self):
        return self.img_size // self.patch_size

    @property
    def num_patches(self):
        return self.grid * self.grid * (self.num_frames // self.tubelet_size)


@dataclass
class TeacherCfg:
    """clip_b16 (clip.py:234-253)."""
    width: int = 768
    layers: int = 12
    heads: int = 12
    output_dim: int = 512
    input_resolution: int = 224
    patch_size: int = 16
    kernel_size: int = 1
    ln_eps: float = 1e-5           # nn.LayerNorm default
    return_layers: Sequence[int] = (6, 7, 8, 9, 10, 11)


# --------------------------------------------------------------------------------------------------
# shared pieces
# --------------------------------------------------------------------------------------------------
def sinusoid_table(n_position: int, d_hid: int) -> Tensor:
    """modeling_adaptation.py:41-51 / modeling_finetune.py:225-235: fp64 numpy table cast to fp32, [1, n, d].

    angle(pos, j) = pos / 10000^(2*(j//2)/d); even columns sin, odd columns cos.
    """
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    angle = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab = np.empty_like(angle)
    tab[:, 0::2] = np.sin(angle[:, 0::2])
    tab[:, 1::2] = np.cos(angle[:, 1::2])
    return torch.tensor(tab, dtype=torch.float).unsqueeze(0)


def normalize_frames_u8(frames: Tensor, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> Tensor:
    """Decoded frames uint8 [B,T,H,W,3] -> the normalised fp32 clip [B,3,T,H,W] the models consume.
    Follows src/datasets/kinetics_sparse.py:236-247 per clip: ToTensor (uint8 -> float / 255, :236) ... tensor_normalize
    (:241-243, body :434-451: `tensor - mean`, then `tensor / std`, fp32) ... permute T H W C -> C T H W (:245)."""
    assert frames.dtype == torch.uint8 and frames.shape[-1] == 3
    t = frames.float() / 255.0
    t = t - torch.tensor(mean, dtype=torch.float32)
    t = t / torch.tensor(std, dtype=torch.float32)
    return t.permute(0, 4, 1, 2, 3).contiguous()


def patchify(x: Tensor, tubelet: int, patch: int) -> Tensor:
    """Conv3d with stride == kernel (clip.py:123-128,146; modeling_finetune.py:165-174) as an im2col.

    x [B,3,T,H,W] -> [B, (T/tub)*(H/p)*(W/p), 3*tub*p*p]; token order (t,h,w), feature order (c,kt,kh,kw) —
    the flattening of a Conv3d weight [D,3,tub,p,p].
    """
    B, C, T, H, W = x.shape
    t, h, w = T // tubelet, H // patch, W // patch
    x = x.reshape(B, C, t, tubelet, h, patch, w, patch)
    x = x.permute(0, 2, 4, 6, 1, 3, 5, 7)  # B t h w C kt kh kw
    return x.reshape(B, t * h * w, C * tubelet * patch * patch)


def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


# --------------------------------------------------------------------------------------------------
# teacher: CLIP ViT  (clip.py)
# --------------------------------------------------------------------------------------------------
def teacher_forward(sd: SD, x: Tensor, cfg: TeacherCfg, return_cls: bool = False):
    """clip.VisionTransformer.forward with return_attn=True, mask=None (clip.py:145-188).

    Returns (feat [K,B,T'*HW,C_out] L2-normalised, attn [B*T', HW]) with T' = T / kernel_size.
    """
    B = x.shape[0]
    D, H = cfg.width, cfg.heads
    d = D // H
    # conv1, no bias (clip.py:123-128,146); tokens per frame in (h,w) order (clip.py:148)
    w = sd["conv1.weight"].reshape(D, -1)
    e = patchify(x, cfg.kernel_size, cfg.patch_size) @ w.t()          # [B, T'*HW, D]
    HW = (cfg.input_resolution // cfg.patch_size) ** 2
    Tp = e.shape[1] // HW
    e = e.reshape(B * Tp, HW, D)
    # CLS + positional embedding + ln_pre (clip.py:150-152)
    cls = sd["class_embedding"].reshape(1, 1, D).expand(B * Tp, 1, D)
    h = torch.cat([cls, e], dim=1) + sd["positional_embedding"]
    h = _ln(h, sd["ln_pre.weight"], sd["ln_pre.bias"], cfg.ln_eps)
    L = HW + 1
    keep = []
    attn_cls = None
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        # ResidualAttentionBlock (clip.py:55-64): nn.MultiheadAttention = packed in_proj, heads split of E,
        # softmax(q k^T / sqrt(d)) v, out_proj
        y = _ln(h, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], cfg.ln_eps)
        qkv = y @ sd[p + "attn.in_proj_weight"].t() + sd[p + "attn.in_proj_bias"]
        q, k, v = qkv.reshape(B * Tp, L, 3, H, d).permute(2, 0, 3, 1, 4)   # [BT,H,L,d]
        pr = torch.softmax((q * (d ** -0.5)) @ k.transpose(-2, -1), dim=-1)
        if i == cfg.layers - 1:
            # need_weights=True averages heads (clip.py:50, 95-96); caller keeps CLS row, drops CLS column (:183)
            attn_cls = pr.mean(dim=1)[:, 0, 1:]
        o = (pr @ v).transpose(1, 2).reshape(B * Tp, L, D)
        h = h + (o @ sd[p + "attn.out_proj.weight"].t() + sd[p + "attn.out_proj.bias"])
        y = _ln(h, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], cfg.ln_eps)
        y = y @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"]
        y = y * torch.sigmoid(1.702 * y)                                   # QuickGELU (clip.py:29-31)
        h = h + (y @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"])
        if i in cfg.return_layers:                                          # clip.py:99-100
            keep.append(h)
    z = torch.stack(keep)                                                   # [K, BT, L, D]
    K = z.shape[0]
    # ln_post on patch tokens, regroup frames of a clip, project, L2-normalise (clip.py:167-173)
    z = _ln(z[:, :, 1:, :], sd["ln_post.weight"], sd["ln_post.bias"], cfg.ln_eps)
    z = z.reshape(K, B, Tp * HW, D) @ sd["proj"]
    z = z / z.norm(dim=-1, keepdim=True)
    if return_cls:
        # image embedding of OpenAI CLIP's encode_image: ln_post(CLS of the last block) @ proj (then L2-normalised by clip_infer)
        c = _ln(h[:, 0, :], sd["ln_post.weight"], sd["ln_post.bias"], cfg.ln_eps) @ sd["proj"]
        return z, attn_cls, c / c.norm(dim=-1, keepdim=True)
    return z, attn_cls


# --------------------------------------------------------------------------------------------------
# mask construction
# --------------------------------------------------------------------------------------------------
def n_visible(n_tokens: int, mask_ratio: float) -> int:
    """run_stage1.py:380."""
    return n_tokens - int(n_tokens * mask_ratio)


def multinomial_mask(attn: Tensor, q: Tensor, mask_ratio: float, clips: int) -> Tensor:
    """run_stage1.py:379-387 with the sampler's noise made explicit.

    torch.multinomial(attn, N) without replacement draws q ~ Exp(1) and returns topk(attn / q, N)
    (ATen MultinomialKernel; verified against torch.multinomial under a shared generator in
    oracle/make_golden.py).  The first N_vis draws are the visible patches.  Returns bool [clips, T*N],
    True = masked.
    """
    BT, N = attn.shape
    n_vis = n_visible(N, mask_ratio)
    order = torch.topk(attn / q, N, dim=-1).indices
    m = torch.ones(BT, N)
    m[torch.arange(BT).view(-1, 1).repeat(1, n_vis), order[:, :n_vis]] = 0
    return m.view(clips, -1).to(torch.bool)


def greedy_masks(attn: Tensor, mask_ratio: float, k: int) -> Tensor:
    """utils.get_greedy_masks (utils.py:89-120): committee member i unmasks attention ranks i, i+k, ...

    Returns bool [k, BT, N], True = masked.
    """
    BT, N = attn.shape
    n_unmask = N - int(N * mask_ratio)
    order = attn.sort(dim=1, descending=True).indices
    masks = torch.ones(k, BT, N, dtype=torch.bool)
    for i in range(k):
        masks[i].scatter_(1, order[:, i::k][:, :n_unmask], False)
    return masks


def visible_indices(mask: Tensor) -> Tensor:
    """Ascending positions of visible tokens, the order boolean indexing `x[~mask]` yields
    (modeling_adaptation.py:153, run_stage1.py:393). [B, N_vis] int64."""
    B = mask.shape[0]
    return (~mask).nonzero()[:, 1].reshape(B, -1)


# --------------------------------------------------------------------------------------------------
# student blocks (modeling_finetune.py)
# --------------------------------------------------------------------------------------------------
def student_block(sd: SD, p: str, x: Tensor, heads: int, eps: float, keep_scale: Optional[Tensor] = None) -> Tensor:
    """Block.forward without layer-scale (modeling_finetune.py:143-146) = Attention.forward (:100-119) +
    Mlp.forward (:66-73).  keep_scale [2,B] = DropPath factors floor(keep+u)/keep (timm 0.4.12 drop_path) of the
    attention branch and of the MLP branch (two independent draws per block, :145-146), or None."""
    B, N, D = x.shape
    d = D // heads
    y = _ln(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    bias = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.v_bias"]), sd[p + "attn.v_bias"]])
    qkv = F.linear(y, sd[p + "attn.qkv.weight"], bias).reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (d ** -0.5), qkv[1], qkv[2]
    a = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    y = (a @ v).transpose(1, 2).reshape(B, N, D)
    y = F.linear(y, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    if keep_scale is not None:
        y = y * keep_scale[0].view(B, 1, 1)
    x = x + y
    y = _ln(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    y = F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    if keep_scale is not None:
        y = y * keep_scale[1].view(B, 1, 1)
    return x + y


def student_forward(sd: SD, x: Tensor, mask: Tensor, cfg: StudentCfg, clip_only: bool = False,
                    keep_scales: Optional[Tensor] = None):
    """AdaptationVisionTransformer.forward (modeling_adaptation.py:304-334) over
    AdaptationVisionTransformerEncoder.forward_features (:131-169), no CLS token, sinusoid pos-embed.

    mask bool [B, N] (True = masked).  keep_scales [depth, 2, B] optional DropPath factors.
    Returns x_clip [K,B,N_vis,C_out] if clip_only else (x_vis_normed [B,N_vis,D], x_clip).
    """
    B = x.shape[0]
    D = cfg.embed_dim
    w = sd["encoder.patch_embed.proj.weight"].reshape(D, -1)
    t = patchify(x, cfg.tubelet_size, cfg.patch_size) @ w.t() + sd["encoder.patch_embed.proj.bias"]   # :132
    pos = sinusoid_table(cfg.num_patches, D)
    t = t + pos                                                                                       # :141-144
    xv = t[~mask].reshape(B, -1, D)                                                                   # :153
    taps = []
    last = max(cfg.return_layers)
    for i in range(cfg.depth):
        ks = None if keep_scales is None else keep_scales[i]
        xv = student_block(sd, f"encoder.blocks.{i}.", xv, cfg.num_heads, cfg.ln_eps, ks)
        if i in cfg.return_layers:
            taps.append(xv)                                                                           # :163-164
        if i == last and clip_only:
            break                                                                                     # :165-166
    z = _ln(torch.stack(taps), sd["encoder.norm.weight"], sd["encoder.norm.bias"], cfg.ln_eps)        # :168
    K = z.shape[0]
    cpos = sinusoid_table(cfg.num_patches, D).repeat(B, 1, 1)[~mask].view(B, -1, D)                   # :318-319
    z = z + cpos.unsqueeze(0)
    outs = []
    for kk in range(K):                                                                               # :322-325
        p = f"clip_decoder.{kk}."
        y = F.linear(z[kk], sd[p + "head.weight"], sd[p + "head.bias"])                               # :204
        y = _ln(y, sd[p + "norm.weight"], sd[p + "norm.bias"], cfg.ln_eps)
        outs.append(y / y.norm(dim=-1, keepdim=True))                                                 # :207
    x_clip = torch.stack(outs)
    if clip_only:
        return x_clip
    x_vis = _ln(xv, sd["encoder.norm.weight"], sd["encoder.norm.bias"], cfg.ln_eps)                   # :177-178
    return x_vis, x_clip


def alignment_loss(outputs: Tensor, targets: Tensor, kind: str = "l2") -> Tensor:
    """run_stage1.py:430-433: 'l2' is the cosine form of the shipped config; 'mse' / 'smooth_l1' / 'l1' are the
    nn.MSELoss / nn.SmoothL1Loss / nn.L1Loss instances built at :403-408 (default reduction 'mean', beta 1)."""
    if kind == "l2":
        return (2 - 2 * (outputs * targets).sum(dim=-1)).mean()
    if kind == "mse":
        return F.mse_loss(outputs, targets)
    if kind == "smooth_l1":
        return F.smooth_l1_loss(outputs, targets)
    if kind == "l1":
        return F.l1_loss(outputs, targets)
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------------------------------
# stage-1 step  (run_stage1.py:360-456)
# --------------------------------------------------------------------------------------------------
def stage1_step(student_sd: SD, teacher_sd: SD, videos: Tensor, q: Tensor, scfg: StudentCfg, tcfg: TeacherCfg,
                mask_ratio: float = 0.8, keep_scales: Optional[Tensor] = None, with_grads: bool = True,
                attn_override: Optional[Tensor] = None, clip_loss_type: str = "l2", clip_loss_data: str = "mixed",
                n_source: Optional[int] = None):
    """One UMT masked-distillation step with mask_type='attention', src_classifier=None.  q = Exp(1) noise [B*T', HW] consumed
    by the mask sampler.  clip_loss_data 'source' / 'target' (run_stage1.py:418-423): only the first n_source (= B_s) clips /
    the remaining ones enter the loss; 'mixed' (:424-425) uses all.  Returns a dict with every intermediate the parity tests
    compare (outputs / targets are the UNSLICED tensors)."""
    B = videos.shape[0]
    with torch.no_grad():
        feat, attn = teacher_forward(teacher_sd, videos, tcfg)                        # :375
        a = attn if attn_override is None else attn_override
        mask = multinomial_mask(a, q, mask_ratio, B)                                   # :379-387
        K, C = feat.shape[0], feat.shape[-1]
        targets = feat[~mask.unsqueeze(0).repeat(K, 1, 1)].reshape(K, B, -1, C)        # :389-393
    params = {k: v.detach().clone().requires_grad_(with_grads) for k, v in student_sd.items()}
    out = student_forward(params, videos, mask, scfg, clip_only=True, keep_scales=keep_scales)   # :415
    if clip_loss_data == "source":
        loss = alignment_loss(out[:, :n_source], targets[:, :n_source], clip_loss_type)                 # :418-420
    elif clip_loss_data == "target":
        loss = alignment_loss(out[:, n_source:], targets[:, n_source:], clip_loss_type)                 # :421-423
    elif clip_loss_data == "mixed":
        loss = alignment_loss(out, targets, clip_loss_type)                           # :430-433
    else:
        raise NotImplementedError(clip_loss_data)                                      # :426-427
    res = dict(attn=attn, mask=mask, vis_idx=visible_indices(mask), targets=targets, outputs=out.detach(),
               loss=loss.detach())
    if with_grads:
        loss.backward()                                                                # utils.py:609
        grads = {k: v.grad for k, v in params.items() if v.grad is not None}
        res["grads"] = grads
        res["grad_norm"] = torch.norm(torch.stack([g.norm(2) for g in grads.values()]), 2)   # utils.py:631-643
    return res


# --------------------------------------------------------------------------------------------------
# stage-2 step  (engine_for_finetuning.py:37-40 + modeling_finetune.py:356-383)
# --------------------------------------------------------------------------------------------------
def vit_forward(sd: SD, x: Tensor, cfg: StudentCfg, keep_scales: Optional[Tensor] = None) -> Tensor:
    """modeling_finetune.VisionTransformer.forward with use_mean_pooling=True, linear head."""
    D = cfg.embed_dim
    w = sd["patch_embed.proj.weight"].reshape(D, -1)
    t = patchify(x, cfg.tubelet_size, cfg.patch_size) @ w.t() + sd["patch_embed.proj.bias"]
    t = t + sinusoid_table(cfg.num_patches, D)                                         # :364-365
    for i in range(cfg.depth):
        ks = None if keep_scales is None else keep_scales[i]
        t = student_block(sd, f"blocks.{i}.", t, cfg.num_heads, cfg.ln_eps, ks)
    t = _ln(t.mean(1), sd["fc_norm.weight"], sd["fc_norm.bias"], cfg.ln_eps)           # :374-376
    return F.linear(t, sd["head.weight"], sd["head.bias"])                             # :382


def stage2_step(sd: SD, videos: Tensor, labels: Tensor, cfg: StudentCfg, with_grads: bool = True):
    params = {k: v.detach().clone().requires_grad_(with_grads) for k, v in sd.items()}
    logits = vit_forward(params, videos, cfg)
    loss = F.cross_entropy(logits, labels)                                             # engine_for_finetuning.py:39
    res = dict(logits=logits.detach(), loss=loss.detach())
    if with_grads:
        loss.backward()
        res["grads"] = {k: v.grad for k, v in params.items() if v.grad is not None}
    return res


# --------------------------------------------------------------------------------------------------
# stage-3 pieces  (run_stage3.py:427-625, utils.py:55-68)
# --------------------------------------------------------------------------------------------------
def pool_outputs(x: Tensor) -> Tensor:
    """run_stage3.py:333-338, use_cls_token=False: mean over tokens of the normed encoder output."""
    return x.mean(dim=1)


def clip_zero_shot(image_features: Tensor, text_features: Tensor, clips: int) -> Tensor:
    """utils.clip_infer after encode_image (utils.py:62-68): per-frame softmax(100 * cos) averaged over frames.
    image_features [clips*T, C], text_features [n_cls, C] -> [clips, n_cls]."""
    i = image_features / image_features.norm(dim=-1, keepdim=True)
    t = text_features / text_features.norm(dim=-1, keepdim=True)
    sim = (100 * i @ t.t()).softmax(dim=-1)
    return sim.reshape(clips, -1, sim.shape[-1]).mean(dim=1)


def pseudo_label_fusion(logits_full_t: Tensor, logits_masked_t: Tensor, clip_probs: Tensor,
                        clip_threshold: float = 0.5, tgt_ratio: float = 1.0, conf_weighted: bool = True):
    """selection_strategy == 'clip_matchORconf', train_masked=True (run_stage3.py:489-490, 556-616).

    logits_full_t [B,C] student logits on the full target clip (no grad), logits_masked_t [k,B,C] committee
    logits, clip_probs [B,C] zero-shot CLIP probabilities.  Returns dict(sel_mask, pseudo, msp, loss_t).
    """
    probs = torch.softmax(logits_full_t.detach(), dim=-1)
    msp, preds = probs.max(dim=-1)                                                     # :489-490
    clip_msp, clip_preds = clip_probs.max(dim=-1)                                      # :558
    match = clip_preds == preds                                                        # :562
    conf = torch.logical_xor(msp >= clip_threshold, clip_msp >= clip_threshold) & ~match   # :565-569
    sel = conf | match                                                                 # :572
    pseudo = preds                                                                     # :576 overrides :575
    if sel.sum() > 0:
        ratio = sel.sum().item() / len(sel)                                            # :601
        w = msp[sel] if conf_weighted else torch.ones_like(msp[sel])
        ce = F.cross_entropy(logits_masked_t[-1][sel], pseudo[sel], reduction="none")  # :606-612
        loss_t = tgt_ratio * ratio * torch.mean(w * ce)                                # :614-615
    else:
        loss_t = torch.zeros(())
    return dict(sel_mask=sel, pseudo=pseudo, msp=msp, match=match, conf=conf, loss_t=loss_t)


def stage3_step(student_sd: SD, teacher_sd: SD, cls_w: Tensor, cls_b: Tensor, text_features: Tensor, videos_s: Tensor,
                labels_s: Tensor, videos_t: Tensor, scfg: StudentCfg, tcfg: TeacherCfg, mask_ratio: float = 0.8, k: int = 2,
                clip_threshold: float = 0.5, src_ratio: float = 1.0, tgt_ratio: float = 1.0, with_grads: bool = True,
                videos_t_aug: Optional[Tensor] = None, keep_scales: Optional[dict] = None):
    """Collaborative self-training step, masking_type='clip_attention', selection_strategy='clip_matchORconf',
    train_masked=True, conf_weighted_loss=True (run_stage3.py:427-642).  The OpenAI-CLIP zero-shot tower (absent here) is
    replaced by the teacher trunk's CLS embedding and the given text matrix; src_classifier is frozen (run_stage3.py:1193,1264).
    videos_t_aug: the augmented view of the target clips (return_aug_for_val, run_stage3.py:405-413): it replaces videos_t for the
    teacher attention (:434-451, `videos[B_s:]`) and the masked committee (:499); the full-token target pass (:480) and the
    zero-shot head (:557) keep the plain view.  keep_scales: optional DropPath factors per forward, keys 's', 't', 'm'
    ([depth,2,B] each; 'm' has k*B_t samples) — the model is in train mode for every pass (:352)."""
    Bs, Bt = videos_s.shape[0], videos_t.shape[0]
    v_mask = videos_t if videos_t_aug is None else videos_t_aug
    ks = keep_scales or {}
    with torch.no_grad():
        _, attn, img = teacher_forward(teacher_sd, v_mask, tcfg, return_cls=True)              # :434-451 (attn only is used)
        if videos_t_aug is not None:
            _, _, img = teacher_forward(teacher_sd, videos_t, tcfg, return_cls=True)            # clip_infer sees videos_t (:557)
    params = {k_: v.detach().clone().requires_grad_(with_grads) for k_, v in student_sd.items()}
    full_s = torch.zeros(Bs, scfg.num_patches, dtype=torch.bool)
    full_t = torch.zeros(Bt, scfg.num_patches, dtype=torch.bool)
    enc_s, _ = student_forward(params, videos_s, full_s, scfg, clip_only=False, keep_scales=ks.get("s"))   # :475
    logits_s = F.linear(pool_outputs(enc_s), cls_w, cls_b)                                      # :476-477
    with torch.no_grad():
        enc_t, _ = student_forward(params, videos_t, full_t, scfg, clip_only=False, keep_scales=ks.get("t"))   # :480-481
        logits_full_t = F.linear(pool_outputs(enc_t), cls_w, cls_b)                             # :482-483
    loss_s = F.cross_entropy(logits_s, labels_s)                                                # :486
    gm = greedy_masks(attn, mask_ratio, k)                                                      # :496
    T = attn.shape[0] // Bt
    masks = gm.reshape(k, Bt, T * attn.shape[1]).reshape(k * Bt, -1)                            # 'k (B T) N -> (k B) (T N)'  :497
    videos_tk = v_mask.repeat(k, 1, 1, 1, 1)                                                    # :499
    enc_m, _ = student_forward(params, videos_tk, masks, scfg, clip_only=False, keep_scales=ks.get("m"))   # :502
    logits_masked = F.linear(pool_outputs(enc_m), cls_w, cls_b).reshape(k, Bt, -1)              # :503-505
    clip_probs = clip_zero_shot(img, text_features, Bt)                                         # :557
    fus = pseudo_label_fusion(logits_full_t, logits_masked, clip_probs, clip_threshold, tgt_ratio)
    loss = src_ratio * loss_s + fus["loss_t"]                                                   # :625
    res = dict(attn=attn, masks=gm, logits_s=logits_s.detach(), logits_full_t=logits_full_t, logits_masked=logits_masked.detach(),
               clip_probs=clip_probs, sel_mask=fus["sel_mask"], pseudo=fus["pseudo"], msp=fus["msp"], loss_s=loss_s.detach(),
               loss_t=fus["loss_t"].detach(), loss=loss.detach())
    if with_grads:
        loss.backward()
        res["grads"] = {k_: v.grad for k_, v in params.items() if v.grad is not None}
    return res
