from .options import args, set_args, reset_args, parse_args  # noqa: F401
