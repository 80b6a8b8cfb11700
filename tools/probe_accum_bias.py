"""Does tcgen05 accumulate fp32 with round-toward-zero?  C = A B^T with positive bf16 operands through
ctk_gemm_bf16_splitk at several split-K factors: a truncating accumulator loses ~0.5 ulp per MMA step, always downwards,
so the signed relative error is negative and proportional to the steps per accumulator (K / 16 / splits)."""
import os
import sys
from ctypes import c_int

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "torch-unet_b200"))
from ctk import _lib as L   # noqa: E402

torch.manual_seed(0)
M = N = 128
for K in (4096, 16384, 65536):
    a = (torch.rand(M, K) + 0.5).to(torch.bfloat16)
    b = (torch.rand(N, K) + 0.5).to(torch.bfloat16)
    ref = a.double() @ b.double().t()
    ad, bd = a.cuda(), b.cuda()
    for splits in (1, 4, 16, 64):
        part = torch.empty(splits, M, N, device="cuda")
        L.call("ctk_gemm_bf16_splitk", L.ptr(ad), L.ptr(bd), c_int(M), c_int(N), c_int(K), c_int(splits), L.ptr(part), L.stream())
        got = part.double().sum(0).cpu()
        rel = (got - ref) / ref
        print(f"K {K:6d} splits {splits:3d} steps/accumulator {K // 16 // splits:5d}: signed rel err mean {rel.mean().item():+.3e} "
              f"max |rel| {rel.abs().max().item():.3e}   (steps x 2^-24 = {K // 16 // splits * 2 ** -24:.3e})")
# mixed-sign operands (what the network sees): the bias acts on partial sums of either sign, the error is smaller but not RN-like
a = torch.randn(M, 16384).to(torch.bfloat16)
b = torch.randn(N, 16384).to(torch.bfloat16)
ref = a.double() @ b.double().t()
for splits in (1, 16):
    part = torch.empty(splits, M, N, device="cuda")
    L.call("ctk_gemm_bf16_splitk", L.ptr(a.cuda()), L.ptr(b.cuda()), c_int(M), c_int(N), c_int(16384), c_int(splits), L.ptr(part), L.stream())
    got = part.double().sum(0).cpu()
    scale = (a.double().abs() @ b.double().abs().t())
    print(f"randn K 16384 splits {splits}: rms err / rms ref {((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item():.3e}; "
          f"fp32 sequential RN sum for comparison {(((a.float().cuda() @ b.float().cuda().t()).double().cpu() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item():.3e}")
