"""Synthetic clip generator — stands in for src/datasets/* (decord decode + augmentation), which needs data files.

Only the BATCH CONTRACT of the reference loaders is reproduced:
  stage 1  (videos [B,3,T,224,224] fp32 ImageNet-normalised, mask placeholder, label)     mae.py:220-222, build.py:63-69
  stage 2  (clip, label, index, {})                                                       kinetics_sparse.py:159
  stage 3  target (vid, vid_aug, label, name)                                             kinetics_sparse.py:174-180
Batches live in PINNED host memory (like DataLoader(pin_memory=True), run_stage1.py:700-708) and are seeded by
`seed + rank` (run_stage1.py:613).  `noise` is the Exp(1) draw the attention mask sampler consumes
(torch.multinomial's internal noise made explicit, run_stage1.py:382).
"""
import torch


class SyntheticStage1Loader:
    def __init__(self, batch_size, num_frames=8, img_size=224, steps=10, seed=0, rank=0, n_distinct=2, num_classes=12,
                 frames_per_token=1, pin=True, uint8=False):
        """uint8=True: batches carry DECODED frames uint8 [B,T,H,W,3] (what decord / NVDEC hand over, kinetics_sparse.py:
        loadvideo_decord) instead of the normalised fp32 clip; the engine normalises on the device (ops.patchify_u8)."""
        g = torch.Generator().manual_seed(seed + rank)
        self.steps = steps
        self.batches = []
        pin = pin and torch.cuda.is_available()
        HW = (img_size // 16) ** 2
        for _ in range(n_distinct):
            if uint8:
                v = torch.randint(0, 256, (batch_size, num_frames, img_size, img_size, 3), generator=g, dtype=torch.uint8)
            else:
                v = torch.randn(batch_size, 3, num_frames, img_size, img_size, generator=g)
            q = torch.empty(batch_size * (num_frames // frames_per_token), HW).exponential_(1, generator=g)
            y = torch.randint(0, num_classes, (batch_size,), generator=g)
            if pin:
                v, q = v.pin_memory(), q.pin_memory()
            self.batches.append((v, -1, y, q))

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            yield self.batches[i % len(self.batches)]
