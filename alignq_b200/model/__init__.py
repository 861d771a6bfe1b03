from .quantization import *  # noqa: F401,F403
