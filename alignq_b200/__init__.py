"""alignq_b200 -- B200-native implementation of AlignQ's per-layer quantization hot path.

Layout mirrors one reference experiment directory (``model/quantization.py``, ``model/resnet.py``,
``utils/admm.py``, ``utils/optimizer.py``, ``utils/options.py``) so that a user switches with
``from alignq_b200.model.quantization import *`` etc.  All arithmetic runs in the hand-written
sm_100a kernels of ``csrc/`` behind the C ABI of ``include/alignq_b200.h``.
"""
from .utils.options import args, set_args, reset_args, parse_args  # noqa: F401
from ._lib import AlignQError, LIB_PATH, load as load_library  # noqa: F401
from .model.quantization import (uniform_quantize, cdf, weight_quantize_fn, activation_quantize_fn,  # noqa: F401
                                 activation_quantize_fn2, corr, conv2d_Q_fn, linear_Q_fn)
from .utils.admm import ADMM  # noqa: F401
from .utils.optimizer import SGD, ADMM_OPT  # noqa: F401

__version__ = "0.1.0"
