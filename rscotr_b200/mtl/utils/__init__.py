from .optimizer import build_optimizer, build_optimizer_constructor, MTLOptimizerConstructor  # noqa: F401
from .misc import init_random_seed, set_random_seed  # noqa: F401
