"""How fast do two runs of the SAME oracle training loop diverge when one weight set is perturbed by
~1 ulp?  Sets the tolerance band for the product-vs-oracle multi-iteration tests."""
import sys, os, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alignq_b200 as aq
from alignq_b200.model import resnet
from alignq_b200.utils.train import QATStep
from oracle import models_oracle as MO
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
out = {}
for variant in ("A", "B"):
    B = 16
    aq.set_args(variant=variant, train_batch_size=B, bitW=8, abitW=8, act_range=2, method="ours")
    torch.manual_seed(0)
    x = torch.randn(B, 3, 32, 32, device=dev)
    t = torch.randint(0, 10, (B,), device=dev)
    prod = resnet.resnet20_quant(8, 8, "second")
    sd = MO.deterministic_fill(prod.state_dict(), seed=3)
    prod.load_state_dict(sd); prod.to(dev).train()
    o1 = MO.OracleResNet([3, 3, 3], 8, 8, variant, 2.0, dim=B); o1.load_state_dict(sd); o1.to(dev).train()
    o2 = MO.OracleResNet([3, 3, 3], 8, 8, variant, 2.0, dim=B); o2.load_state_dict(sd); o2.to(dev).train()
    with torch.no_grad():
        for p in o2.parameters():
            p.mul_(1.0 + 1e-7)
    step = QATStep(prod); t1 = MO.OracleTrainer(o1); t2 = MO.OracleTrainer(o2)
    rows = []
    for it in range(6):
        lp = float(step.step(x, t)); l1 = float(t1.step(x, t)[0]); l2 = float(t2.step(x, t)[0])
        def worst(a, b):
            return max(float((p.detach().double() - q.detach().double()).norm() / (q.detach().double().norm() + 1e-30))
                       for p, q in zip(a.parameters(), b.parameters()))
        rows.append({"it": it, "loss_prod": lp, "loss_oracle": l1, "loss_oracle_perturbed": l2,
                     "param_err_prod_vs_oracle": worst(prod, o1), "param_err_oracle_vs_perturbed": worst(o2, o1)})
    out[variant] = rows
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/chaos.json", "w"), indent=1)
