"""Configuration namespace of the hot path.

The reference parses ``argparse`` flags at import time into a module-global ``args``
(``utils/options.py:5-98``, ``utils/options_office.py``) and the L1 code reads it at call time
(``model/quantization.py:12,98``; ``utils/optimizer.py:13,231``).  This module keeps the same
global object and field names but does NOT parse ``sys.argv`` on import; use ``set_args(...)``
(or ``parse_args(argv)`` for the reference's flag spellings).  Defaults are the reference's.

Extra fields (no reference counterpart): ``variant`` selects which of the three quantizer
variants the math follows ('A' = cdf_alignment/*, 'B' = cdf_alignment_admm/resnet-*,
'C' = cdf_alignment_admm/{dann,dsan}_office); ``gram_mode`` picks the Gram numerics
('fp32' FFMA parity mode, 'tf32x3' / 'bf16' tcgen05 modes); ``fuse_bn_act`` lets the model files run
BatchNorm -> activation quantizer -> ReLU as one fused kernel pair on channels_last inputs;
``admm_param_grads=False`` skips d trans_loss / d(alterD, gamma), which ``ADMM_OPT`` never reads
(it applies closed-form Z/U updates, optimizer.py:97-124).  Data-parallel (one process per GPU): ``dp_gram``
= 'replica' (every rank its own [b, b] Gram and ADMM(b): the reference at train_batch_size = b) or 'feature'
(global-batch Gram: all-to-all to feature slices, partial sums, one all-reduce -- utils/dp_gram.py);
``sync_bn`` (False | True / "nccl" | "peer") gives the fused BatchNorm -> act-quant kernels the batch statistics of
the GLOBAL batch: their fp64 (sum, sum of squares) accumulators are all-reduced with NCCL between the two launches,
or -- "peer" -- exchanged inside the kernels through NVLink peer-mapped memory (model/fused.py).
``own_conv`` ("off" | "tf32" | "tf32x3") runs the 3x3 / stride-1 / Cin == Cout quantized convolutions on the
hand-written tcgen05 kernels (model/conv_tc.py) instead of the library convolution, for the channel counts in
``own_conv_channels`` (default (16,): the wide, shallow layers where they beat cuDNN by 1.4-2.5x; at C = 32 / 64 the
kernels are correct but cuDNN's tiles are still faster -- profiles/r02_conv_bench.json); ``own_conv_stem`` (default
True, effective unless ``own_conv`` is "off") runs the first-layer convolution (3 -> 16 / 32 channels, 3x3, stride 1) on the
direct fp32 kernels of csrc/conv_stem.cu.
``own_wgrad_channels`` / ``own_dgrad_channels`` add channel counts whose WEIGHT / DATA gradient alone runs on the own kernel
(the rest stays with the library).
``fuse_dgrad_bn`` (default True; effective where an own C = 16 convolution follows a fused bn-act layer) moves the reduce
pass of that bn-act layer's backward into the epilogue of the convolution's data-gradient kernel (model/conv_tc.py).
``fused_head`` lets ``QATStep`` run the average pool -> classifier -> cross-entropy tail of models that offer ``forward_ce``
on the two kernels of csrc/head_ce.cu (model/fused.py:avgpool_linear_ce).
"""
from __future__ import annotations

import argparse
from types import SimpleNamespace

_DEFAULTS = dict(
    gpus=[0], bitW=2, abitW=2, act_range=2, lam=1.0, lam2=4.0, method="ours", stage="second",
    train_batch_size=128, eval_batch_size=100, lr=0.04, momentum=0.9, weight_decay=1e-4,
    variant="A", gram_mode="fp32", store_weight_attrs=True, fuse_bn_act=False, admm_param_grads=True,
    dp_gram="replica", sync_bn=False, own_conv="off", own_conv_channels=(16,), own_wgrad_channels=(), own_dgrad_channels=(), own_conv_stem=True, fused_head=False, async_wgrad=False, fuse_dgrad_bn=True,
)

args = SimpleNamespace(**_DEFAULTS)


def set_args(**kw):
    """Update the global config; unknown keys are rejected so typos fail loudly."""
    for k, v in kw.items():
        if k not in _DEFAULTS:
            raise KeyError(f"unknown alignq_b200 option {k!r}; known: {sorted(_DEFAULTS)}")
        if k == "variant" and v not in ("A", "B", "C"):
            raise ValueError("variant must be 'A', 'B' or 'C'")
        if k == "gram_mode" and v not in ("fp32", "tf32x3", "bf16"):
            raise ValueError("gram_mode must be 'fp32', 'tf32x3' or 'bf16'")
        if k == "own_conv" and v not in ("off", "tf32", "tf32x3"):
            raise ValueError("own_conv must be 'off', 'tf32' or 'tf32x3'")
        if k == "dp_gram" and v not in ("replica", "feature"):
            raise ValueError("dp_gram must be 'replica' (per-rank [b,b] Gram) or 'feature' (global-batch Gram)")
        setattr(args, k, v)
    return args


def reset_args():
    for k, v in _DEFAULTS.items():
        setattr(args, k, list(v) if isinstance(v, list) else v)
    return args


def parse_args(argv=None):
    """Accept the reference's command-line spellings (utils/options.py:31-91)."""
    p = argparse.ArgumentParser(description="alignq_b200")
    p.add_argument("--gpus", type=int, nargs="+", default=[0])
    p.add_argument("--bitW", type=int, default=_DEFAULTS["bitW"])
    p.add_argument("--abitW", type=int, default=_DEFAULTS["abitW"])
    p.add_argument("--act_range", type=float, default=_DEFAULTS["act_range"])
    p.add_argument("--lam", type=float, default=_DEFAULTS["lam"])
    p.add_argument("--lam2", type=float, default=_DEFAULTS["lam2"])
    p.add_argument("--method", type=str, default=_DEFAULTS["method"])
    p.add_argument("--stage", type=str, default=_DEFAULTS["stage"])
    p.add_argument("--train_batch_size", type=int, default=_DEFAULTS["train_batch_size"])
    p.add_argument("--eval_batch_size", type=int, default=_DEFAULTS["eval_batch_size"])
    p.add_argument("--lr", type=float, default=_DEFAULTS["lr"])
    p.add_argument("--momentum", type=float, default=_DEFAULTS["momentum"])
    p.add_argument("--weight_decay", type=float, default=_DEFAULTS["weight_decay"])
    p.add_argument("--variant", type=str, default=_DEFAULTS["variant"], choices=["A", "B", "C"])
    p.add_argument("--gram_mode", type=str, default=_DEFAULTS["gram_mode"], choices=["fp32", "tf32x3", "bf16"])
    ns, _ = p.parse_known_args(argv)
    return set_args(**vars(ns))
