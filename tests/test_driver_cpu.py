"""CPU: host logic of the training driver -- the reference's LR schedule semantics, checkpoint keys and
file layout, AverageMeter / accuracy -- without launching a kernel (the GPU run of the same driver is
tests/test_gpu_models.py::test_driver_trains_checkpoints_and_resumes)."""
import os

import torch
from torch.optim.lr_scheduler import MultiStepLR

import alignq_b200 as aq
from alignq_b200.utils import common as C
from alignq_b200.utils.driver import scheduler_step_epoch


def test_average_meter_and_accuracy_match_the_reference_definitions():
    m = C.AverageMeter()
    m.update(2.0, 3)
    m.update(4.0, 1)
    assert m.val == 4.0 and m.sum == 10.0 and m.count == 4 and m.avg == 2.5
    logits = torch.tensor([[0.1, 0.9, 0.0, 0.0, 0.0, 0.0], [0.8, 0.05, 0.05, 0.04, 0.03, 0.03], [0.0, 0.1, 0.2, 0.3, 0.4, 0.5]])
    tgt = torch.tensor([1, 2, 0])
    p1, p5 = C.accuracy(logits, tgt, topk=(1, 5))
    assert abs(float(p1) - 100.0 / 3) < 1e-4 and abs(float(p5) - 200.0 / 3) < 1e-4


def test_scheduler_step_epoch_is_the_reference_closed_form():
    """main.py:127-128 calls s.step(epoch) at the top of every epoch: lr = lr0 * gamma ** (#milestones <= epoch)."""
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=0.04)
    sch = MultiStepLR(opt, [80, 150], gamma=0.1)
    for epoch, want in [(0, 0.04), (79, 0.04), (80, 0.004), (149, 0.004), (150, 0.0004), (199, 0.0004), (3, 0.04)]:
        scheduler_step_epoch(sch, epoch)
        assert abs(opt.param_groups[0]["lr"] - want) < 1e-12 and sch.last_epoch == epoch
    sd = sch.state_dict()
    assert sd["last_epoch"] == 3 and sorted(sd["milestones"].elements()) == [80, 150]


def test_checkpoint_directory_layout_and_keys(tmp_path):
    from types import SimpleNamespace
    args = SimpleNamespace(job_dir=str(tmp_path / "exp"), reset=False, bitW=8)
    ck = C.checkpoint(args)
    assert os.path.isdir(ck.ckpt_dir) and os.path.isdir(ck.run_dir) and "bitW: 8" in open(ck.job_dir / "config.txt").read()
    state = {"state_dict_t": {"w": torch.ones(2)}, "best_prec1": 12.5, "best_prec5": 50.0, "optimizer_t": {}, "scheduler_t": {},
             "epoch": 3}
    ck.save_model(state, 3, is_best=True)
    assert os.path.isfile(ck.ckpt_dir / "model_3.pt") and os.path.isfile(ck.ckpt_dir / "model_best.pt")
    back = torch.load(ck.ckpt_dir / "model_best.pt")
    assert set(back) == {"state_dict_t", "best_prec1", "best_prec5", "optimizer_t", "scheduler_t", "epoch"} and back["epoch"] == 3


def test_sgd_state_dict_is_the_reference_format():
    """'optimizer_t' in the checkpoint: per-parameter state is exactly {'momentum_buffer'} (no private bookkeeping)."""
    ps = [torch.nn.Parameter(torch.randn(3)), torch.nn.Parameter(torch.randn(2))]
    opt = aq.SGD(ps, lr=0.04, momentum=0.9, weight_decay=1e-4)
    opt.state[ps[0]]["momentum_buffer"] = torch.ones(3)
    opt.state[ps[0]]["alignq_first"] = False
    opt.state[ps[1]]["momentum_buffer"] = torch.empty(2)        # allocated, never stepped
    opt.state[ps[1]]["alignq_first"] = True
    sd = opt.state_dict()
    assert list(sd["state"].keys()) == [0] and set(sd["state"][0].keys()) == {"momentum_buffer"}
    assert sd["param_groups"][0]["lr"] == 0.04 and sd["param_groups"][0]["momentum"] == 0.9
    opt2 = aq.SGD([torch.nn.Parameter(torch.randn(3)), torch.nn.Parameter(torch.randn(2))], lr=0.1, momentum=0.9)
    opt2.load_state_dict(sd)
    st = opt2.state[opt2.param_groups[0]["params"][0]]
    assert torch.equal(st["momentum_buffer"], torch.ones(3)) and st["alignq_first"] is False
    assert opt2.param_groups[0]["lr"] == 0.04
