"""Measures, on the GPU, every gradient of the hot path against the oracle's fp64 autograd, beside the
reference's own fp32 noise floor (oracle fp32 GPU-eager vs oracle fp64).  Output: JSON on stdout / --out.

Test tooling (imports oracle/): sets the bars asserted in tests/test_gpu_parity.py.
For each case:  mine = max|mine - ref64| / max|ref64|,  floor = max|oracle32 - ref64| / max|ref64|,
and the element-wise worst |d| / (1e-5 |ref| + 1e-6 max|ref|) ("bar units": <= 1 passes north_star's 1e-5).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq                                   # noqa: E402
from oracle import alignq_oracle as O                      # noqa: E402

DEV = "cuda"


def errs(mine, ref64, orc32):
    ref64 = ref64.double()
    mx = float(ref64.abs().max())
    tol = 1e-5 * ref64.abs() + 1e-6 * mx

    def one(a):
        d = (a.double() - ref64).abs()
        return {"max_over_max": float(d.max()) / mx, "relnorm": float(d.norm() / ref64.norm()),
                "bar_units": float((d / tol).max())}
    return {"mine": one(mine), "floor": one(orc32)}


def weight_cases(out):
    shapes = [(16, 3, 3, 3), (16, 16, 3, 3), (64, 64, 3, 3), (32, 1, 3, 3), (64, 32, 1, 1), (2048, 512, 1, 1),
              (64, 3, 7, 7), (5, 7)]
    for variant in ("A", "B"):
        for k in (4, 8):
            torch.manual_seed(4)
            aq.set_args(variant=variant, bitW=k)
            for shape in shapes:
                fan = int(np.prod(shape[1:]))
                w = (torch.randn(*shape) * (2.0 / fan) ** 0.5).to(DEV)
                gup = torch.randn_like(w)
                wr = w.clone().requires_grad_(True)
                (aq.weight_quantize_fn(k, "second")(wr) * gup).sum().backward()
                wo = w.clone().requires_grad_(True)
                (O.weight_quantize(wo, k, variant)[0] * gup).sum().backward()
                w64 = w.double().clone().requires_grad_(True)
                (O.weight_quantize(w64, k, variant)[0] * gup.double()).sum().backward()
                out[f"gw {variant} k{k} {shape}"] = errs(wr.grad, w64.grad, wo.grad)


def fused_cases(out):
    cases = [("B", 128, (16, 16, 16)), ("B", 128, (64, 8, 8)), ("C", 28, (64, 14, 14)), ("B", 100, (32, 16, 16)),
             ("B", 128, (16, 32, 32)), ("C", 28, (256, 14, 14)), ("B", 100, (3, 11, 13)), ("C", 5, (3, 7, 5))]
    for variant, B, shape in cases:
        for pure in (False, True):
            torch.manual_seed(8)
            dim = 128
            x0 = torch.randn(B, *shape, device=DEV)
            gy = torch.randn_like(x0)
            Z0 = torch.rand(dim, dim, device=DEV)
            U0 = torch.rand(dim, dim, device=DEV)
            Fn = aq.activation_quantize_fn if variant == "B" else aq.activation_quantize_fn2

            def oracle(dt):
                xo = x0.to(dt).clone().requires_grad_(True)
                Zo = Z0.to(dt).clone().requires_grad_(True)
                Uo = U0.to(dt).clone().requires_grad_(True)
                yo, lo, Do = O.activation_quantize_admm(xo, 8, Zo, Uo, "second", variant, 2.0)
                (lo if pure else (yo * gy.to(dt)).sum() + lo).backward()
                return xo.grad, Zo.grad, Uo.grad, lo.detach(), Do.detach()
            g64 = oracle(torch.float64)
            g32 = oracle(torch.float32)
            for mode in ("fp32", "tf32x3", "bf16"):
                aq.set_args(variant=variant, act_range=2, method="ours", gram_mode=mode, admm_param_grads=True)
                admm = aq.ADMM(dim).to(DEV)
                with torch.no_grad():
                    admm.alterD.copy_(Z0)
                    admm.gamma.copy_(U0)
                x = x0.clone().requires_grad_(True)
                y, loss = Fn(8, "second", admm)(x)
                (loss if pure else (y * gy).sum() + loss).backward()
                tag = f"{'pure' if pure else 'full'} {mode} {variant} B={B} F={x0[0].numel()}"
                out[f"gx {tag}"] = errs(x.grad, g64[0], g32[0])
                if not pure and mode != "bf16":
                    out[f"gZ {tag}"] = errs(admm.alterD.grad, g64[1], g32[1])
                    out[f"gU {tag}"] = errs(admm.gamma.grad, g64[2], g32[2])
                out[f"loss {tag}"] = {"mine": abs(float(loss) - float(g64[3])) / abs(float(g64[3])),
                                      "floor": abs(float(g32[3]) - float(g64[3])) / abs(float(g64[3]))}
                dmax = float(g64[4].abs().max())
                out[f"D {tag}"] = {"mine_over_maxD": float((admm.D.double() - g64[4]).abs().max()) / dmax,
                                   "floor_over_maxD": float((g32[4].double() - g64[4]).abs().max()) / dmax,
                                   "maxD": dmax}


def bn_cases(out):
    import copy
    import torch.nn as nn
    import torch.nn.functional as F
    from alignq_b200.model.fused import bn_act
    for variant in ("A", "B"):
        for shape, relu in [((128, 16, 32, 32), True), ((128, 32, 16, 16), True), ((128, 64, 8, 8), True),
                            ((64, 64, 8, 8), False), ((4, 456, 8, 8), True)]:
            torch.manual_seed(0)
            aq.set_args(variant=variant, act_range=2, abitW=8, fuse_bn_act=True, method="none")
            C = shape[1]
            x0 = (torch.randn(shape, device=DEV) * 1.5 + 0.3).contiguous(memory_format=torch.channels_last)
            gy = torch.randn(shape, device=DEV).contiguous(memory_format=torch.channels_last)
            bn = nn.BatchNorm2d(C).to(DEV).train()
            with torch.no_grad():
                bn.weight.copy_(1 + 0.2 * torch.randn(C))
                bn.bias.copy_(0.1 * torch.randn(C))
            bn32, bn64 = copy.deepcopy(bn), copy.deepcopy(bn).double()
            actq = aq.activation_quantize_fn(8, "second")
            x = x0.clone().requires_grad_(True)
            y = bn_act(bn, actq, x, relu)
            (y * gy).sum().backward()

            def ref(bnm, dt):
                xr = x0.to(dt).clone().requires_grad_(True)
                yr = O.activation_quantize(bnm(xr), 8, "second", variant, 2.0)
                yr = F.relu(yr) if relu else yr
                (yr * gy.to(dt)).sum().backward()
                return yr.detach(), xr.grad
            y32, g32 = ref(bn32, torch.float32)
            y64, g64 = ref(bn64, torch.float64)
            n = y.numel()
            tag = f"bn_act {variant} {shape} relu={relu}"
            out[tag] = {"code_mismatch_frac_vs_cudnn32": float(((y - y32).abs() > 1e-6).sum()) / n,
                        "code_mismatch_frac_vs_fp64": float(((y.double() - y64).abs() > 1e-6).sum()) / n,
                        "floor_cudnn32_vs_fp64": float(((y32.double() - y64).abs() > 1e-6).sum()) / n,
                        "gx": errs(x.grad, g64, g32)}
            # gradients on the elements whose codes agree (a flipped code moves the ReLU mask of that element)
            same = ((y.double() - y64).abs() <= 1e-6) & ((y32.double() - y64).abs() <= 1e-6)
            d_m = ((x.grad.double() - g64).abs() * same).max() / g64.abs().max()
            d_f = ((g32.double() - g64).abs() * same).max() / g64.abs().max()
            out[tag]["gx_same_codes_max_over_max"] = {"mine": float(d_m), "floor": float(d_f)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = {}
    weight_cases(res)
    fused_cases(res)
    bn_cases(res)
    txt = json.dumps(res, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            f.write(txt)
    for k, v in res.items():
        print(k, json.dumps(v))
