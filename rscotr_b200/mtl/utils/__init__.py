from .optimizer import build_optimizer  # noqa: F401
