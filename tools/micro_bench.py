"""SURVEY.md 8(d) micro-benchmarks: every kernel of the hot path on the shapes of the BASELINE configs, timed on
the GPU (CUDA-graph replay of the C-ABI call, so Python/ctypes launch overhead is excluded), with the reference's
GPU-eager path (the oracle restatement run with device=cuda: the same ATen kernels the reference launches)
timed beside it with CUDA events.  Writes gpurun_out/micro_bench.json.  Dev/measurement tool."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq  # noqa: E402
from alignq_b200 import _lib as L  # noqa: E402
from alignq_b200.utils.weight_bank import WeightBank  # noqa: E402
from oracle import alignq_oracle as O  # noqa: E402
from tools.tc_probe_util import graph_time  # noqa: E402

dev = "cuda"
lib = L.load()
HBM = 6549.1
out = {"gpu": torch.cuda.get_device_name(0), "hbm_peak_gbs": HBM,
       "timing": "ours: CUDA-graph replay of the C-ABI call; reference: oracle restatement in GPU eager mode, CUDA events"}


def eager_time(fn, iters=10, warm=3):
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


# ---- 1. activation quantizer, every distinct activation shape of the configs ---------------------------------
torch.manual_seed(0)
ACT_SHAPES = [(128, 16, 32, 32), (128, 32, 16, 16), (128, 64, 8, 8), (256, 144, 32, 32), (128, 456, 8, 8),
              (28, 256, 56, 56), (256, 1 << 20)]
for shape in ACT_SHAPES:
    x = torch.randn(*shape, device=dev)
    gy = torch.randn_like(x)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    n = x.numel()
    for variant, vid in (("A", 0), ("B", 1)):
        for k in (8, 4):
            if k == 4 and variant == "B":
                continue
            tf = graph_time(lambda: L.check(lib.alignq_act_fwd(x.data_ptr(), y.data_ptr(), 0, n, k, 2.0, vid, 0, L.stream_ptr()), "act_fwd"), reps=5, iters=5)
            tb = graph_time(lambda: L.check(lib.alignq_act_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr(), n, k, 2.0, vid, 0, L.stream_ptr()), "act_bwd"), reps=5, iters=5)
            rec = {"fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "fwdbwd_gbs_at_20B": 20 * n / (tf + tb) / 1e9,
                   "frac_of_hbm_peak": 20 * n / (tf + tb) / 1e9 / HBM}
            if n <= (1 << 25) and k == 8:
                xr = x.clone().requires_grad_(True)

                def ref():
                    xr.grad = None
                    O.activation_quantize(xr, k, "second", variant, 2.0).backward(gy)
                tr = eager_time(ref)
                rec["reference_gpu_eager_fwdbwd_us"] = tr * 1e6
                rec["speedup_vs_reference_gpu_eager"] = tr / (tf + tb)
            out[f"act_{variant}_k{k}_{'x'.join(map(str, shape))}"] = rec
            print(f"act_{variant}_k{k}_{shape}", rec, flush=True)
    del x, gy, y, gx
torch.cuda.empty_cache()

# ---- 2. weight quantizer: the whole-model multi-tensor call of every config -----------------------------------
from alignq_b200.model.resnet import resnet20_quant, resnet56_quant  # noqa: E402
from alignq_b200.model.mobilenetV2 import mobile_v2  # noqa: E402
from alignq_b200.model.densenet import densenet_40_quant  # noqa: E402
from alignq_b200.model.dann import resnet50_dann  # noqa: E402

MODELS = [("cfg1_resnet20_W8", "A", lambda: resnet20_quant(8, 8, "second")),
          ("cfg2_resnet56_W8", "B", lambda: resnet56_quant(8, 8, "second")),
          ("cfg3_mobilenetv2_W4", "A", lambda: mobile_v2(4, 4, "second")),
          ("cfg4_densenet40_W8", "A", lambda: densenet_40_quant(8, 8, "second")),
          ("cfg5_resnet50_dann_W8", "C", lambda: resnet50_dann(8, 8, "second"))]
for name, variant, ctor in MODELS:
    aq.reset_args()
    aq.set_args(variant=variant, bitW=8, abitW=8, act_range=2, method="ours", train_batch_size=32)
    torch.manual_seed(0)
    model = ctor().to(dev)
    bank = WeightBank(model)
    g = torch.randn_like(bank.flat)
    gw = torch.empty_like(bank.flat)
    nel = bank.flat.numel()
    tf = graph_time(bank.quantize_all, reps=5, iters=5)

    def bwd():
        L.check(lib.alignq_wq_backward(bank.flat.data_ptr(), g.data_ptr(), None, bank.seg_off.data_ptr(),
                                       bank.chunk_seg.data_ptr(), bank.seg_chunk0.data_ptr(), len(bank.params), bank.nchunks,
                                       bank.w_bit, bank.stats.data_ptr(), gw.data_ptr(), 0, bank.bwd_ws.data_ptr(),
                                       L.stream_ptr()), "wq_backward")
    tb = graph_time(bwd, reps=5, iters=5)
    ws = [p.detach().clone().requires_grad_(True) for p in bank.params]
    gs = [torch.randn_like(w) for w in ws]

    def ref():
        for w, gg in zip(ws, gs):
            w.grad = None
            O.weight_quantize(w, bank.w_bit, variant)[0].backward(gg)
    tr = eager_time(ref, iters=3, warm=1)
    rec = {"tensors": len(bank.params), "elements": nel, "fwd_us": tf * 1e6, "bwd_us": tb * 1e6,
           "fwdbwd_gbs_at_20B": 20 * nel / (tf + tb) / 1e9, "launches": 4,
           "reference_gpu_eager_fwdbwd_us": tr * 1e6, "speedup_vs_reference_gpu_eager": tr / (tf + tb)}
    out[f"weight_{name}"] = rec
    print(f"weight_{name}", rec, flush=True)
    bank.release()
    del model, bank, ws, gs, g, gw
torch.cuda.empty_cache()

# ---- 3. fused act-quant + Gram/ADMM layer (forward incl. loss and dL/dD, backward) ------------------------------
MODES = {"fp32": 0, "tf32x3": 1, "bf16": 2}
for B, Fd, eps in [(128, 16384, 0.0), (128, 8192, 0.0), (128, 4096, 0.0), (28, 802816, 1e-5), (28, 100352, 1e-5),
                   (224, 802816, 1e-5)]:
    x = torch.randn(B, Fd, device=dev)
    gy = torch.randn_like(x)
    y, gx = torch.empty_like(x), torch.empty_like(x)
    D = torch.empty(B, B, device=dev); dL = torch.empty(B, B, device=dev); loss = torch.empty((), device=dev)
    gl = torch.ones((), device=dev)
    Z = torch.rand(B, B, device=dev); U = torch.rand(B, B, device=dev)
    ws = torch.empty(int(lib.alignq_gram_ws_bytes(B, Fd)), dtype=torch.uint8, device=dev)
    n = x.numel()
    for mode, mid in MODES.items():
        if B > 128 and mode != "fp32":
            continue                                         # tcgen05 paths cover B <= 128 (DESIGN.md gaps)
        if mode == "fp32" and n > (1 << 26) and B <= 128:
            reps = 1
        else:
            reps = 3
        fwd = lambda: L.check(lib.alignq_act_admm_fwd(x.data_ptr(), B, Fd, 8, 2.0, eps, Z.data_ptr(), U.data_ptr(), B, 0.2, 0.3,
                              y.data_ptr(), D.data_ptr(), loss.data_ptr(), dL.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "fwd")
        bwd = lambda: L.check(lib.alignq_act_admm_bwd(x.data_ptr(), gy.data_ptr(), dL.data_ptr(), gl.data_ptr(), B, Fd, 8, 2.0, eps,
                              gx.data_ptr(), ws.data_ptr(), ws.numel(), mid, L.stream_ptr()), "bwd")
        tf = graph_time(fwd, reps=reps, iters=3)
        tb = graph_time(bwd, reps=reps, iters=3)
        rec = {"fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "fwdbwd_gbs_at_20B": 20 * n / (tf + tb) / 1e9,
               "frac_of_hbm_peak": 20 * n / (tf + tb) / 1e9 / HBM, "tflops_at_8B2F": 8.0 * B * B * Fd / (tf + tb) / 1e12}
        out[f"admm_layer_{mode}_B{B}_F{Fd}"] = rec
        print(f"admm_layer_{mode}_B{B}_F{Fd}", rec, flush=True)
    if n <= (1 << 25):
        xr = x.clone().requires_grad_(True)
        variant = "B" if eps == 0.0 else "C"

        def ref():
            xr.grad = None
            yy, ll, _ = O.activation_quantize_admm(xr, 8, Z, U, "second", variant, 2.0)
            torch.autograd.backward([yy, ll], [gy, torch.ones_like(ll)])
        tr = eager_time(ref, iters=5, warm=2)
        out[f"admm_layer_reference_gpu_eager_B{B}_F{Fd}"] = {"fwdbwd_us": tr * 1e6}
        print(f"admm_layer_reference_gpu_eager_B{B}_F{Fd}", tr * 1e6, flush=True)
    del x, gy, y, gx, ws
    torch.cuda.empty_cache()

# ---- 4. bf16 Gram micro-shape (tensor roofline) ---------------------------------------------------------------------
xb = torch.randn(256, 1 << 20, device=dev).to(torch.bfloat16)
G = torch.empty(256, 256, device=dev)
ws2 = torch.empty(int(lib.alignq_gram_bf16_ws_bytes(256)), dtype=torch.uint8, device=dev)
t = graph_time(lambda: L.check(lib.alignq_gram_bf16(xb.data_ptr(), 256, 1 << 20, 1, G.data_ptr(), ws2.data_ptr(), ws2.numel(), L.stream_ptr()), "g16"),
               reps=5, iters=5)
tm = eager_time(lambda: torch.matmul(xb, xb.t()), iters=10)
out["gram_bf16_B256_F1048576"] = {"us": t * 1e6, "tflops": 2 * 256 * 256 * (1 << 20) / t / 1e12,
                                  "frac_of_burst_peak_1637": 2 * 256 * 256 * (1 << 20) / t / 1e12 / 1637.2,
                                  "torch_matmul_bf16_us": tm * 1e6, "torch_matmul_tflops": 2 * 256 * 256 * (1 << 20) / tm / 1e12}
print("gram_bf16", out["gram_bf16_B256_F1048576"], flush=True)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/micro_bench.json", "w"), indent=1)
print("done")
