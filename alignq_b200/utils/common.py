"""Run bookkeeping of the reference's training scripts (``utils/common.py``): running averages, top-k accuracy and the
checkpoint directory with the reference's file names and keys, so a run can be resumed by either code base.

Mirrors cdf_alignment/resnet-20-cifar-10/utils/common.py:12-98 (``AverageMeter`` :12-27, ``checkpoint`` :29-61,
``accuracy`` :83-98).  The only arithmetic here is ``accuracy`` (a top-k over [B, classes] logits, once per step,
off the quantization hot path): it stays a handful of torch calls exactly like the reference's.
"""
from __future__ import annotations

import datetime
import os
import shutil
from pathlib import Path

import torch


class AverageMeter(object):
    """Computes and stores the average and current value (common.py:12-27)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0.0
        self.avg = 0.0
        self.sum = 0.0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class checkpoint():
    """``<job_dir>/checkpoint/model_<epoch>.pt`` (+ ``model_best.pt``), ``<job_dir>/run/``, ``<job_dir>/config.txt``
    (common.py:29-61).  ``args`` is any namespace with ``job_dir`` (and optionally ``reset``)."""

    def __init__(self, args):
        now = datetime.datetime.now().strftime('%Y-%m-%d-%H:%M:%S')
        self.args = args
        self.job_dir = Path(args.job_dir)
        self.ckpt_dir = self.job_dir / 'checkpoint'
        self.run_dir = self.job_dir / 'run'
        if getattr(args, 'reset', False) and self.job_dir.exists():
            shutil.rmtree(self.job_dir)                    # the reference shells out to `rm -rf` (common.py:40)
        for d in (self.job_dir, self.ckpt_dir, self.run_dir):
            os.makedirs(d, exist_ok=True)
        with open(self.job_dir / 'config.txt', 'w') as f:
            f.write(now + '\n\n')
            for arg in vars(args):
                f.write('{}: {}\n'.format(arg, getattr(args, arg)))
            f.write('\n')

    def save_model(self, state, epoch, is_best):
        save_path = f'{self.ckpt_dir}/model_{epoch}.pt'
        torch.save(state, save_path)
        if is_best:
            shutil.copyfile(save_path, f'{self.ckpt_dir}/model_best.pt')
        return save_path


def accuracy(output, target, topk=(1,)):
    """Computes the precision@k for the specified values of k (common.py:83-98)."""
    with torch.no_grad():
        maxk = max(topk)
        batch_size = target.size(0)
        _, pred = output.topk(maxk, 1, True, True)
        pred = pred.t()
        correct = pred.eq(target.reshape(1, -1).expand_as(pred))
        res = []
        for k in topk:
            correct_k = correct[:k].reshape(-1).float().sum(0, keepdim=True)
            res.append(correct_k.mul_(100.0 / batch_size))
        return res
