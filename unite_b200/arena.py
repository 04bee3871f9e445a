"""Flat parameter arena: every parameter of a model lives in ONE fp32 device buffer.

Layout `[decay params | no-decay params]` (the two AdamW groups of src/optim_factory.py:76-118: a parameter
is "no-decay" when it is 1-D, its name ends in ".bias", or it is in the model's no_weight_decay() set), each
group ordered by BACKWARD COMPLETION (decoders, final norm, last block ... first block, patch embed) so that
gradient ranges become final front-to-back and can be all-reduced while backward is still running
(unite_b200/ddp.py).  Parallel buffers: `grads` (fp32, what wgrad GEMMs red.add into), `w16` (bf16 shadow the
GEMMs read), and — owned by the fused optimizer — exp_avg / exp_avg_sq.

nn.Parameters keep working as usual: `p.data` and `p.grad` are views into the arenas, so state_dict(),
load_state_dict(), torch optimizers and checkpoints see ordinary tensors with the reference's key names.
"""
from typing import Callable, Dict, List, Tuple

import torch

ALIGN = 8  # elements: 32 B in fp32, 16 B in bf16 (TMA base alignment)


def _round_up(x, a):
    return (x + a - 1) // a * a


class ParamArena:
    def __init__(self, module: torch.nn.Module, device, order_key: Callable[[str], Tuple], no_decay: Callable[[str, torch.Tensor], bool],
                 gap_after: Callable[[str, torch.Tensor], int] = lambda n, p: 0):
        """gap_after(name, p) -> number of always-zero elements reserved right after that parameter (used to make
        [q_bias | 0 | v_bias] one contiguous 3*D vector, modeling_finetune.py:104)."""
        named = [(n, p) for n, p in module.named_parameters()]
        decay = sorted([(n, p) for n, p in named if not no_decay(n, p)], key=lambda t: order_key(t[0]))
        nodecay = sorted([(n, p) for n, p in named if no_decay(n, p)], key=lambda t: order_key(t[0]))
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        for n, p in decay:
            self.offsets[n] = (off, p.numel())
            off = _round_up(off + p.numel(), ALIGN)
        self.n_decay = off
        for n, p in nodecay:
            self.offsets[n] = (off, p.numel())
            off = _round_up(off + p.numel() + gap_after(n, p), ALIGN)
        off = _round_up(off, 4 * ALIGN)
        self.numel = off
        self.device = torch.device(device)
        self.params = torch.zeros(off, device=device, dtype=torch.float32)
        self.grads = torch.zeros(off, device=device, dtype=torch.float32)
        self.w16 = torch.zeros(off, device=device, dtype=torch.bfloat16)
        self.names: List[str] = [n for n, _ in decay + nodecay]
        self._params = {n: p for n, p in named}
        with torch.no_grad():
            for n, p in named:
                o, k = self.offsets[n]
                view = self.params[o:o + k].view(p.shape)
                view.copy_(p.detach().to(device))
                p.data = view
                p.grad = None
        self._grads_attached = False
        self.w16_version = -1

    # ---- views ------------------------------------------------------------------------------
    def p32(self, name):
        o, k = self.offsets[name]
        return self.params[o:o + k].view(self._params[name].shape)

    def g32(self, name):
        o, k = self.offsets[name]
        return self.grads[o:o + k].view(self._params[name].shape)

    def b16(self, name):
        o, k = self.offsets[name]
        return self.w16[o:o + k].view(self._params[name].shape)

    def range_of(self, names):
        """[lo, hi) element range covering the given parameter names (they must be contiguous in the arena)."""
        lo = min(self.offsets[n][0] for n in names)
        hi = max(_round_up(self.offsets[n][0] + self.offsets[n][1], ALIGN) for n in names)
        return lo, hi

    # ---- gradient attachment: p.grad becomes a view of the grad arena ---------------------
    def attach_grads(self):
        """Called by backward.  If the user reset grads (zero_grad(set_to_none=True)) the arena is cleared and
        re-attached; otherwise gradients keep ACCUMULATING like torch's `p.grad += g`."""
        params = self._params
        if all(p.grad is None for p in params.values()):
            self.grads.zero_()
        for n, p in params.items():
            if p.grad is None or p.grad.data_ptr() != self.grads.data_ptr() + 4 * self.offsets[n][0]:
                p.grad = self.g32(n)

    def params_version(self):
        return sum(p._version for p in self._params.values())
