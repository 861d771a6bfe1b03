"""One small invocation of every kernel family, for compute-sanitizer --tool memcheck / racecheck."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq
from alignq_b200.model.fused import bn_act
from alignq_b200.model.resnet import resnet20_quant
from alignq_b200.utils.train import QATStep
dev = "cuda"
torch.manual_seed(0)
aq.set_args(variant="A", bitW=8, abitW=8, act_range=2, train_batch_size=8, fuse_bn_act=True)
x = torch.randn(8, 16, 8, 8, device=dev, requires_grad=True)
y = aq.activation_quantize_fn(8, "second")(x); y.sum().backward()
w = torch.randn(16, 16, 3, 3, device=dev, requires_grad=True)
aq.weight_quantize_fn(8, "second")(w).sum().backward()
bn = torch.nn.BatchNorm2d(16).to(dev)
xc = torch.randn(8, 16, 8, 8, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
bn_act(bn, aq.activation_quantize_fn(8, "second"), xc, True, residual=xc.detach().clone()).sum().backward()
m = resnet20_quant(8, 8, "second").to(dev).train()
st = QATStep(m, channels_last=True)
xi = torch.randn(8, 3, 32, 32, device=dev).contiguous(memory_format=torch.channels_last); t = torch.randint(0, 10, (8,), device=dev)
for _ in range(2): st.step(xi, t)
for mode in ("fp32", "tf32x3", "bf16"):
    aq.set_args(variant="B", method="ours", gram_mode=mode, admm_param_grads=True)
    admm = aq.ADMM(8).to(dev)
    xa = torch.randn(8, 4, 6, 6, device=dev, requires_grad=True)
    ya, la = aq.activation_quantize_fn(8, "second", admm)(xa)
    (ya.sum() + la).backward()
    aq.ADMM_OPT([admm.alterD, admm.gamma]).step([0], [1], [admm.D], [admm.alterD], [admm.gamma], [0.2], [0.3])
    aq.corr(xa.detach().view(8, -1), xa.detach().view(8, -1))
from alignq_b200 import _lib as L
lib = L.load()
xb = torch.randn(40, 256, device=dev).to(torch.bfloat16); G = torch.empty(40, 40, device=dev)
ws = torch.empty(int(lib.alignq_gram_bf16_ws_bytes(40)), dtype=torch.uint8, device=dev)
L.check(lib.alignq_gram_bf16(xb.data_ptr(), 40, 256, 1, G.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "g16")
torch.cuda.synchronize()
print("sanitize_mini ok")
