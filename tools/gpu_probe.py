"""GPU-side probe: ATen arithmetic facts the bit-exact kernels rely on, and first micro-timings.
Writes gpurun_out/probe.json.  Development tool, not part of the product or the tests."""
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq  # noqa: E402
from alignq_b200 import _lib as L  # noqa: E402
from oracle import alignq_oracle as O  # noqa: E402

out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
dev = "cuda"
torch.manual_seed(0)

# ---- ATen facts ---------------------------------------------------------------------------
q = torch.arange(0, 256, device=dev, dtype=torch.float32)
for n in (3, 15, 255, 65535):
    qq = torch.arange(0, n + 1, device=dev, dtype=torch.float32)
    div = qq / n
    mul32 = qq * torch.tensor(1.0, dtype=torch.float32).div(n).item()
    out[f"div_by_{n}_equals_mul_f32_recip"] = bool(torch.equal(div, qq * (torch.ones(1, device=dev) / n)))
    out[f"div_by_{n}_equals_true_div"] = bool(torch.equal(div, qq / torch.full((1,), float(n), device=dev)))
x = torch.randn(1 << 22, device=dev)
out["div_sqrt2_equals_mul_recip"] = bool(torch.equal(x / math.sqrt(2), x * 0.70710677))
out["div_sqrt2_equals_true_div"] = bool(torch.equal(x / math.sqrt(2), x / torch.full((1,), math.sqrt(2), device=dev)))
for variant in ("A", "B"):
    cg = O.activation_codes(x, 8, variant, 2.0)
    cc = O.activation_codes(x.cpu(), 8, variant, 2.0)
    out[f"gpu_vs_cpu_eager_code_mismatch_{variant}"] = int((cg.cpu() != cc).sum())


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


lib = L.load()
# ---- activation quantizer stream (inputs >> L2) ----------------------------------------------
n = 256 * (1 << 20)
x = torch.randn(n, device=dev)
gy = torch.randn(n, device=dev)
y = torch.empty_like(x)
gx = torch.empty_like(x)
s = L.stream_ptr()
for variant in (0, 1):
    for k in (4, 8):
        tf = timeit(lambda: lib.alignq_act_fwd(x.data_ptr(), y.data_ptr(), 0, n, k, 2.0, variant, 0, s))
        tb = timeit(lambda: lib.alignq_act_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), n, k, 2.0, variant, 0, s))
        out[f"act_v{variant}_k{k}"] = {"fwd_ms": tf * 1e3, "bwd_ms": tb * 1e3, "fwd_gbs": 8 * n / tf / 1e9,
                                       "bwd_gbs": 12 * n / tb / 1e9, "fwdbwd_gbs": 20 * n / (tf + tb) / 1e9}
tc = timeit(lambda: y.copy_(x))
out["torch_copy_gbs"] = 8 * n / tc / 1e9
# reference GPU-eager on a smaller slice (memory heavy)
xs = x[: 1 << 24].clone().requires_grad_(True)
gys = gy[: 1 << 24]


def ref_fb():
    xs.grad = None
    yy = O.activation_quantize(xs, 8, "second", "A", 2.0)
    yy.backward(gys)


tr = timeit(ref_fb, iters=5, warm=2)
out["ref_gpu_eager_act_fwdbwd_gbs_at_20B"] = 20 * (1 << 24) / tr / 1e9
del x, gy, y, gx, xs, gys
torch.cuda.empty_cache()

# ---- small activation shapes (launch-bound regime) ---------------------------------------------
for shape in [(128, 16, 32, 32), (128, 64, 8, 8), (256, 144, 32, 32), (28, 256, 56, 56)]:
    x = torch.randn(*shape, device=dev)
    gy = torch.randn_like(x)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    nn_ = x.numel()
    tf = timeit(lambda: lib.alignq_act_fwd(x.data_ptr(), y.data_ptr(), 0, nn_, 8, 2.0, 0, 0, s), iters=50)
    tb = timeit(lambda: lib.alignq_act_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), nn_, 8, 2.0, 0, 0, s), iters=50)
    out[f"act_small_{'x'.join(map(str, shape))}"] = {"fwd_us": tf * 1e6, "bwd_us": tb * 1e6,
                                                     "fwdbwd_gbs": 20 * nn_ / (tf + tb) / 1e9}

# ---- weight quantizer --------------------------------------------------------------------------
aq.set_args(variant="A", bitW=8)
for shape in [(16, 16, 3, 3), (64, 64, 3, 3), (2048, 512, 1, 1), (512, 512, 3, 3)]:
    w = (torch.randn(*shape, device=dev) * 0.05).requires_grad_(True)
    g = torch.randn(*shape, device=dev)
    mod = aq.weight_quantize_fn(8, "second")

    def fb():
        w.grad = None
        mod(w).backward(g)
    tw = timeit(fb, iters=30)
    out[f"weight_{'x'.join(map(str, shape))}"] = {"fwdbwd_us": tw * 1e6, "gbs_at_28B": 28 * w.numel() / tw / 1e9}

# ---- fused act + ADMM (fp32 FFMA mode) -----------------------------------------------------------
aq.set_args(variant="B", method="ours", gram_mode="fp32")
for B, shape in [(128, (16, 32, 32)), (128, (32, 16, 16)), (128, (64, 8, 8)), (28, (256, 56, 56))]:
    admm = aq.ADMM(B).to(dev)
    fn = aq.activation_quantize_fn(8, "second", admm)
    x = torch.randn(B, *shape, device=dev, requires_grad=True)
    gy = torch.randn(B, *shape, device=dev)

    def fwd():
        with torch.no_grad():
            return fn(x)

    def fb():
        x.grad = None
        yy, ll = fn(x)
        torch.autograd.backward([yy, ll], [gy, torch.ones_like(ll)])
    t1 = timeit(fwd, iters=10)
    t2 = timeit(fb, iters=10)
    Fd = x.numel() // B
    out[f"fused_admm_B{B}_F{Fd}"] = {"fwd_us": t1 * 1e6, "fwdbwd_us": t2 * 1e6,
                                      "fwd_tflops": 4 * B * B * Fd / t1 / 1e12,
                                      "fwdbwd_gbs_at_20B": 20 * x.numel() / t2 / 1e9}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
print(json.dumps(out, indent=1))
