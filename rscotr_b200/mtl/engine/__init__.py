from .step import StepEngine  # noqa: F401
from .test import single_gpu_test, multi_gpu_test  # noqa: F401
