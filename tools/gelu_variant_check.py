"""GELU / GELU' epilogues of ub_gemm_bf16 against the exact erf form in fp64 (modeling_finetune.py:56, nn.GELU).

    python tools/gelu_variant_check.py                       # default build: hardware tanh.approx form, |err| <= 6e-4
    UB_LIB_VARIANT=erf python tools/gelu_variant_check.py    # -DUB_GELU_ERF build: Abramowitz-Stegun erf, |err| <= 2e-5
The forward is checked through the fp32-output epilogue (K = 64, bf16-exact products), so that the output rounding does not
hide the difference between the two forms; the derivative goes through the bf16 DGELU epilogue on an identity matrix.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unite_b200 import ops  # noqa: E402

erf = os.environ.get("UB_LIB_VARIANT") == "erf"
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
M, N, K = 512, 256, 64
a = (torch.randn(M, K, device=dev, generator=g) * 0.5).bfloat16()
w = (torch.randn(N, K, device=dev, generator=g) * 0.5).bfloat16()
bias = torch.randn(N, device=dev, generator=g)
out = torch.empty(M, N, device=dev)
ops.gemm(a, w, out, bias=bias, act=ops.UB_ACT_GELU)
pre = a.double() @ w.double().t() + bias.double()
ref = torch.nn.functional.gelu(pre)
err_f = float((out.double() - ref).abs().max())
# derivative: d = I * gelu'(pre)  (A = identity rows, B = identity -> accumulator 1 on the diagonal), bf16 out
n = 256
eye = torch.eye(n, device=dev).bfloat16()
x = torch.linspace(-6, 6, n * n, device=dev).view(n, n).bfloat16()
d = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
ops.gemm(eye, eye, d, b_t=True, act=ops.UB_ACT_DGELU, aux_in=x)
xd = x.double().requires_grad_()
torch.nn.functional.gelu(xd).sum().backward()
err_d = float((d.double().diagonal() - xd.grad.diagonal()).abs().max())
tol_f, tol_d = (2e-5, 6e-3) if erf else (6e-4, 8e-3)
print(f"gelu variant {'erf' if erf else 'tanh.approx'}: forward max abs err {err_f:.3e} (tol {tol_f:.0e}), derivative {err_d:.3e} (tol {tol_d:.0e}, bf16 out)")
ok = err_f <= tol_f and err_d <= tol_d and (erf or err_f > 2e-5)      # the default build must really be the approximate form
print("GELU VARIANT " + ("OK" if ok else "MISMATCH"))
sys.exit(0 if ok else 1)
