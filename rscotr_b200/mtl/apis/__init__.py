from .train import train_model  # noqa: F401
from .test import test_model  # noqa: F401
