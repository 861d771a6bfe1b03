"""Training driver with the reference's epoch structure, learning-rate schedule and checkpoint format:
``main()`` / ``train()`` / ``test()`` of cdf_alignment/resnet-20-cifar-10/main.py:95-153,229-337 and
cdf_alignment_admm/resnet-56-cifar-10/main.py:86-165,241-444, on the drop-in modules, with every iteration executed by
``QATStep`` (CUDA kernels; optionally one CUDA graph per step).

Kept from the reference: ``MultiStepLR(optimizer_t, lr_decay_steps, gamma=lr_gamma)`` stepped with the EPOCH NUMBER at
the top of every epoch (``s.step(epoch)``, main.py:127-128: lr = lr0 * gamma^(#milestones <= epoch)); per-epoch
``test()`` in eval mode; ``best_prec1/5`` tracking; the checkpoint dict keys ``state_dict_t``, ``best_prec1``,
``best_prec5``, ``optimizer_t`` [, ``optimizer_admm``], ``scheduler_t``, ``epoch`` saved as
``<job_dir>/checkpoint/model_<epoch>.pt`` and ``model_best.pt``; ``--resume`` restores all of them (main.py:101-113).
Not kept (out of scope): dataset download / augmentation (any iterable of (inputs, targets) batches is a
loader), tensorboardX writers, ptflops.  The reference passes ``lr_decay_steps`` through ``argparse type=list`` which
explodes the string into characters; here it is a list of ints.
"""
from __future__ import annotations

import logging
from bisect import bisect_right
from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch.optim.lr_scheduler import MultiStepLR

from . import common as utils
from .options import args as qargs
from .train import QATStep

log = logging.getLogger("alignq_b200")


def scheduler_step_epoch(scheduler: MultiStepLR, epoch: int):
    """``scheduler.step(epoch)`` as the reference calls it (main.py:127-128): the closed form
    lr = base_lr * gamma ** (#milestones <= epoch), without the deprecation path of recent torch versions."""
    milestones = sorted(scheduler.milestones.elements())
    scheduler.last_epoch = epoch
    for group, base in zip(scheduler.optimizer.param_groups, scheduler.base_lrs):
        group["lr"] = base * scheduler.gamma ** bisect_right(milestones, epoch)
    scheduler._last_lr = [g["lr"] for g in scheduler.optimizer.param_groups]


class Trainer:
    """``Trainer(model, cfg).fit(loader_train, loader_test)``.  ``cfg`` fields (reference option names,
    utils/options.py:31-91): job_dir, num_epochs, lr, momentum, weight_decay, lr_decay_steps, lr_gamma, lam, lam2,
    print_freq, resume, reset, plus graph=True to replay each training step as a CUDA graph."""

    def __init__(self, model, cfg):
        d = dict(job_dir="experiment/ours/t_0/", num_epochs=200, lr=qargs.lr, momentum=qargs.momentum,
                 weight_decay=qargs.weight_decay, lr_decay_steps=[80, 150], lr_gamma=0.1, lam=qargs.lam, lam2=qargs.lam2,
                 print_freq=200, resume=None, reset=False, graph=False, channels_last=False, world_size=1, process_group=None)
        d.update(cfg if isinstance(cfg, dict) else vars(cfg))
        self.cfg = SimpleNamespace(**d)
        c = self.cfg
        self.model = model
        self.step = QATStep(model, lr=c.lr, momentum=c.momentum, weight_decay=c.weight_decay, lam=c.lam, lam2=c.lam2,
                            channels_last=c.channels_last, world_size=c.world_size, process_group=c.process_group,
                            keep_logits=True)
        self.optimizer_t = self.step.opt
        self.optimizer_admm = self.step.opt_admm
        self.scheduler_t = MultiStepLR(self.optimizer_t, [int(m) for m in c.lr_decay_steps], gamma=c.lr_gamma)
        self.checkpoint = utils.checkpoint(c)
        self.best_prec1 = self.best_prec5 = 0.0
        self.start_epoch = 0
        if c.resume:
            self.load(c.resume)

    # ---- checkpoint format of the reference (main.py:101-113, 136-151) ---------------------------------------
    def state(self, epoch):
        st = {"state_dict_t": self.model.state_dict(), "best_prec1": self.best_prec1, "best_prec5": self.best_prec5,
              "optimizer_t": self.optimizer_t.state_dict(), "scheduler_t": self.scheduler_t.state_dict(), "epoch": epoch + 1}
        if self.optimizer_admm is not None:
            st["optimizer_admm"] = self.optimizer_admm.state_dict()
        return st

    def load(self, path):
        dev = next(self.model.parameters()).device
        ckpt = torch.load(path, map_location=dev, weights_only=False)
        self.best_prec1 = ckpt["best_prec1"]
        self.best_prec5 = ckpt.get("best_prec5", 0.0)
        self.start_epoch = ckpt["epoch"]
        # in place: the parameters may be views of the weight / ADMM banks' flat buffers
        self.model.load_state_dict(ckpt["state_dict_t"])
        self.optimizer_t.load_state_dict(ckpt["optimizer_t"])
        if self.optimizer_admm is not None and "optimizer_admm" in ckpt:
            self.optimizer_admm.load_state_dict(ckpt["optimizer_admm"])
        self.scheduler_t.load_state_dict(ckpt["scheduler_t"])
        if self.step.bank is not None:
            self.step.bank.fresh = False
        self.step.graph = None                       # a captured graph holds the old learning rate table: re-capture
        return ckpt

    # ---- one epoch of train() (main.py:229-337) --------------------------------------------------------------
    def train_epoch(self, loader_train, epoch):
        c = self.cfg
        losses_t, top1, top5 = utils.AverageMeter(), utils.AverageMeter(), utils.AverageMeter()
        self.model.train()
        dev = next(self.model.parameters()).device
        n_it = len(loader_train) if hasattr(loader_train, "__len__") else None
        for i, (inputs, targets) in enumerate(loader_train, 1):
            inputs, targets = inputs.to(dev, non_blocking=True), targets.to(dev, non_blocking=True)
            if c.graph and self.step.graph is None and inputs.shape[0] == qargs.train_batch_size:
                self.step.capture(inputs, targets, warmup=0)
            if self.step.graph is not None and inputs.shape != self.step.static_x.shape:
                loss = self.step._iteration(inputs, targets)          # ragged last batch: eager
            else:
                loss = self.step.step(inputs, targets)
            logits = self.step.logits
            prec1, prec5 = utils.accuracy(logits, targets, topk=(1, min(5, logits.shape[1])))
            losses_t.update(float(loss), inputs.size(0))
            top1.update(float(prec1[0]), inputs.size(0))
            top5.update(float(prec5[0]), inputs.size(0))
            if c.print_freq and i % c.print_freq == 0:
                log.info("Epoch[%d](%d/%s): Train_loss: %.4f (%.4f) Prec@1 %.3f (%.3f), Prec@5 %.3f (%.3f)", epoch, i, n_it,
                         losses_t.val, losses_t.avg, top1.val, top1.avg, top5.val, top5.avg)
        return losses_t.avg, top1.avg, top5.avg

    # ---- test() (main.py:341-383) ------------------------------------------------------------------------------
    @torch.no_grad()
    def test(self, loader_test, epoch=0):
        losses, top1, top5 = utils.AverageMeter(), utils.AverageMeter(), utils.AverageMeter()
        self.model.eval()
        if self.step.bank is not None:
            self.step.bank.fresh = False             # eval re-quantizes per layer from the current weights
        dev = next(self.model.parameters()).device
        for inputs, targets in loader_test:
            inputs, targets = inputs.to(dev), targets.to(dev)
            if self.cfg.channels_last:
                inputs = inputs.contiguous(memory_format=torch.channels_last)
            out = self.model(inputs)
            logits = out[0] if isinstance(out, tuple) else out
            loss = F.cross_entropy(logits, targets)
            prec1, prec5 = utils.accuracy(logits, targets, topk=(1, min(5, logits.shape[1])))
            losses.update(float(loss), inputs.size(0))
            top1.update(float(prec1[0]), inputs.size(0))
            top5.update(float(prec5[0]), inputs.size(0))
        log.info("Prec@1 %.3f Prec@5 %.3f", top1.avg, top5.avg)
        return top1.avg, top5.avg

    # ---- main() loop (main.py:125-153) -------------------------------------------------------------------------
    def fit(self, loader_train, loader_test):
        c = self.cfg
        history = []
        for epoch in range(self.start_epoch, c.num_epochs):
            scheduler_step_epoch(self.scheduler_t, epoch)
            tr = self.train_epoch(loader_train, epoch)
            test_prec1, test_prec5 = self.test(loader_test, epoch)
            is_best = self.best_prec1 < test_prec1
            self.best_prec1 = max(test_prec1, self.best_prec1)
            self.best_prec5 = max(test_prec5, self.best_prec5)
            path = self.checkpoint.save_model(self.state(epoch), epoch + 1, is_best)
            history.append(dict(epoch=epoch, lr=self.optimizer_t.param_groups[0]["lr"], train_loss=tr[0], train_prec1=tr[1],
                                test_prec1=test_prec1, test_prec5=test_prec5, checkpoint=path))
        log.info("Best @prec1: %.3f @prec5: %.3f", self.best_prec1, self.best_prec5)
        return history
