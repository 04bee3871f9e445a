"""Stage-3 collaborative self-training step (run_stage3.py:427-642) on the fused kernels.

masking_type='clip_attention', selection_strategy='clip_matchORconf', train_masked=True, conf_weighted_loss=True — the
shipped configuration (configs/stage3_config.yaml).  Per step:
    teacher attention on the target clips (no projection needed)                   run_stage3.py:434-451
    student, all 1568 tokens, source clips (grad)  -> mean-pool -> src_classifier   :475-477
    student, all tokens, target clips (no grad)    -> logits_full_t                 :480-483
    greedy round-robin masks for k=2 committee members (ub_mask_select, q=NULL)     :493-497, utils.py:89-120
    student on the masked target views (member k-1 with grad) -> logits_masked      :499-505
    zero-shot CLIP probabilities, MatchOrConf selection, pseudo labels              :556-587   (ub_clip_zero_shot, ub_pseudo_label_fusion)
    loss = src_ratio * CE_s + tgt_ratio * |sel|/B_t * mean_sel(msp * CE(masked[-1], pseudo))    :599-625
    backward through both grad-carrying forwards, AdamW on the student only (src_classifier is frozen, :1193/:1264)
The OpenAI-CLIP image/text towers of `clip_infer` are not available offline: the teacher trunk's CLS embedding
(clip.VisionTransformer.cls_features) and a caller-supplied text matrix [n_classes, 512] stand in.
No host synchronisation: |sel| never leaves the device — tgt_ratio*|sel|/B_t*mean_sel(w*CE) == tgt_ratio/B_t * sum_b sel_b*w_b*CE_b.
"""
from typing import Optional

import torch

from . import ops
from .engine import FusedAdamW

BF16, F32, I32, U8 = torch.bfloat16, torch.float32, torch.int32, torch.uint8


class Stage3Engine:
    def __init__(self, student, teacher, cls_weight: torch.Tensor, cls_bias: torch.Tensor, text_features: torch.Tensor,
                 mask_ratio: float = 0.8, k: int = 2, clip_threshold: float = 0.5, src_ratio: float = 1.0, tgt_ratio: float = 1.0,
                 conf_weighted: bool = True, lr: float = 1e-4, weight_decay: float = 0.05, betas=(0.9, 0.999), grad_sync=None):
        self.student, self.teacher = student, teacher
        self.core = student.core()
        self.core.sync_shadow(force=True)
        dev = self.core.arena.device
        self.W = cls_weight.detach().to(dev, F32).contiguous()
        self.b = cls_bias.detach().to(dev, F32).contiguous()
        self.text = text_features.detach().to(dev, F32).contiguous()
        self.mask_ratio, self.k, self.thr = mask_ratio, k, clip_threshold
        self.src_ratio, self.tgt_ratio, self.conf_weighted = src_ratio, tgt_ratio, conf_weighted
        self.optimizer = FusedAdamW(self.core.arena, lr, weight_decay, betas)
        self.grad_sync = grad_sync
        self.loss = torch.zeros(1, device=dev, dtype=F32)
        self.loss_s = torch.zeros(1, device=dev, dtype=F32)
        self.loss_t = torch.zeros(1, device=dev, dtype=F32)
        self._full_idx = {}
        self.last = {}

    def _all_visible(self, B, N, dev):
        if (B, N) not in self._full_idx:
            self._full_idx[(B, N)] = torch.arange(N, device=dev, dtype=I32).repeat(B, 1).contiguous()
        return self._full_idx[(B, N)]

    def _classify(self, x_vis):
        B, N, D = x_vis.shape
        pooled = torch.empty(B, D, device=x_vis.device, dtype=F32)
        ops.meanpool_fwd(x_vis, pooled)                                            # pool_outputs, run_stage3.py:333-338
        logits = torch.empty(B, self.W.shape[0], device=x_vis.device, dtype=F32)
        ops.linear_small_fwd(pooled, self.W, self.b, logits)
        return pooled, logits

    def _backward_from_logits(self, state, pooled, dlogits, n_tokens):
        B, D = pooled.shape
        d_pooled = torch.empty(B, D, device=pooled.device, dtype=F32)
        ops.linear_small_bwd(pooled, self.W, dlogits, d_pooled, None, None)       # classifier frozen: dx only
        g_vis = torch.empty(B, n_tokens, D, device=pooled.device, dtype=F32)
        ops.meanpool_bwd(d_pooled, g_vis)
        self.core.run_backward(state, g_vis=g_vis, grad_sync=None)

    def forward_backward(self, videos_s, labels_s, videos_t, attn_override: Optional[torch.Tensor] = None):
        core, teacher, k = self.core, self.teacher, self.k
        dev = videos_s.device
        Bs, Bt, N = videos_s.shape[0], videos_t.shape[0], core.N
        C = self.W.shape[0]
        # ---- teacher on the target clips: attention map + CLS embedding for the zero-shot head
        _, attn, patches_t = teacher.forward_features(videos_t)
        frames, P = attn.shape
        T = frames // Bt
        img = teacher.cls_features(frames, P)                                      # [Bt*T, 512]
        share = core.tubelet == teacher.kernel_size
        # ---- source clips, all tokens, with grad
        self.loss.zero_(); self.loss_s.zero_(); self.loss_t.zero_()
        xs, _, st_s = core.run_forward(videos_s, self._all_visible(Bs, N, dev), None, None, False, True, want_clip=False)
        pooled_s, logits_s = self._classify(xs)
        dl_s = torch.empty_like(logits_s)
        ops.softmax_ce(logits_s, labels_s.to(I32), None, self.src_ratio / Bs, self.loss_s, dl_s)
        self._backward_from_logits(st_s, pooled_s, dl_s, N)
        # ---- target clips, all tokens, no grad
        xt, _, _ = core.run_forward(videos_t, self._all_visible(Bt, N, dev), patches_t if share else None, None, False, False,
                                    want_clip=False)
        _, logits_full_t = self._classify(xt)
        # ---- committee masks and masked views
        n_vis = P - int(P * self.mask_ratio)
        mask = torch.empty(k, frames * P, device=dev, dtype=U8)
        vis = torch.empty(k, Bt, T * n_vis, device=dev, dtype=I32)
        ops.mask_select(attn if attn_override is None else attn_override, None, mask, vis, None, T, k, n_vis)
        logits_masked = torch.empty(k, Bt, C, device=dev, dtype=F32)
        base = (torch.arange(Bt, device=dev, dtype=I32) * N).view(Bt, 1)
        st_m = pooled_m = None
        for m in range(k):
            grad = m == k - 1                                                      # only the last member trains (run_stage3.py:606)
            abs_rows = (vis[m] + base).reshape(-1).contiguous() if share else None
            xm, _, st = core.run_forward(videos_t, vis[m].contiguous(), patches_t if share else None, None, False, grad, want_clip=False,
                                         abs_rows=abs_rows)
            pooled, lg = self._classify(xm)
            logits_masked[m].copy_(lg)
            if grad:
                st_m, pooled_m = st, pooled
        # ---- zero-shot CLIP + MatchOrConf fusion
        clip_probs = torch.empty(Bt, C, device=dev, dtype=F32)
        ops.clip_zero_shot(img, self.text, clip_probs, T)
        msp = torch.empty(Bt, device=dev, dtype=F32)
        pseudo = torch.empty(Bt, device=dev, dtype=I32)
        sel = torch.empty(Bt, device=dev, dtype=U8)
        weight = torch.empty(Bt, device=dev, dtype=F32)
        ops.pseudo_label_fusion(logits_full_t, clip_probs, self.thr, self.conf_weighted, msp, pseudo, sel, weight)
        dl_t = torch.empty(Bt, C, device=dev, dtype=F32)
        ops.softmax_ce(logits_masked[k - 1], pseudo, weight, self.tgt_ratio / Bt, self.loss_t, dl_t)
        self._backward_from_logits(st_m, pooled_m, dl_t, T * n_vis)
        self.loss.copy_(self.loss_s + self.loss_t)
        self.last = dict(attn=attn, masks=mask.view(k, frames, P).bool(), logits_s=logits_s, logits_full_t=logits_full_t,
                         logits_masked=logits_masked, clip_probs=clip_probs, sel_mask=sel.bool(), pseudo=pseudo, msp=msp)
        return self.loss

    def step(self, videos_s, labels_s, videos_t):
        self.optimizer.zero_grad()
        loss = self.forward_backward(videos_s, labels_s, videos_t)
        scale = self.grad_sync.all_reduce(self.core.arena.grads) if self.grad_sync is not None else 1.0
        self.optimizer.step(grad_scale=scale)
        return loss
