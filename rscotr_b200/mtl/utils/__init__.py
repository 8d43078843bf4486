from .optimizer import build_optimizer, build_optimizer_constructor, MTLOptimizerConstructor  # noqa: F401
