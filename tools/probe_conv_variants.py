"""GPU probe of ctk_conv3x3_tc_eval against a CPU fp32 reference, one subprocess per case (a protocol bug
traps the kernel).  Round 1 used it to settle the UMMA descriptor question: with the halo buffer written
by TMA in SWIZZLE_128B, shifted tap views work with base-offset 0 and a dense 10-pixel row pitch (the
swizzle is a function of absolute shared-memory address bits); setting the base-offset field to
(addr>>7)&7 gives garbage.  Exploratory tool, not a test.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-unet_b200"))


def run_variant(flags: int, cin: int, cout: int, hw: int, n: int) -> None:
    import torch
    import torch.nn.functional as F
    from ctypes import c_float, c_int
    from ctk import _lib
    from ctk._lib import call, ptr, stream
    torch.manual_seed(1)
    x = torch.randn(n, hw, hw, cin).to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3) / (3.0 * cin ** 0.5)).to(torch.bfloat16).float()
    scale = 0.5 + torch.rand(cout)
    scale[::3] *= -1.0
    shift = 0.1 * torch.randn(cout)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    ref = F.max_pool2d(F.leaky_relu(ref, 0.01), 2).permute(0, 2, 3, 1).contiguous()
    xd, wd = x.cuda(), w.cuda()
    wp = torch.empty(9, cout, cin, device="cuda", dtype=torch.bfloat16)
    call("ctk_pack_conv_weight_bf16", ptr(wd), c_int(cout), c_int(cin), ptr(wp), stream())
    out = torch.zeros(n, hw // 2, hw // 2, cout, device="cuda", dtype=torch.bfloat16)
    sc, sh = scale.cuda(), shift.cuda()
    call("ctk_conv3x3_tc_eval", ptr(xd), c_int(n), c_int(hw), c_int(hw), c_int(cin), ptr(wp), c_int(cout), ptr(sc),
         ptr(sh), c_float(0.01), ptr(out), c_int(cout), c_int(0), c_int(flags), stream())
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs()
    print(f"flags={flags} cin={cin} cout={cout} hw={hw} n={n}: max_err={err.max():.4f} mean_err={err.mean():.5f} "
          f"ref_absmax={ref.abs().max():.3f} frac_bad={(err > 0.05 + 0.02 * ref.abs()).float().mean():.4f}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_variant(*[int(a) for a in sys.argv[1:6]])
    else:
        for flags in (0,):
            for (cin, cout, hw, n) in ((64, 128, 32, 2), (128, 256, 16, 3)):
                r = subprocess.run([sys.executable, __file__, str(flags), str(cin), str(cout), str(hw), str(n)],
                                   capture_output=True, text=True, timeout=300)
                print(r.stdout.strip() or f"flags={flags}: no output", flush=True)
                if r.returncode != 0:
                    print(f"flags={flags}: exit {r.returncode}: {r.stderr.strip()[-400:]}", flush=True)
