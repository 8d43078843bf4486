from .iter_runner import IterBasedRunner  # noqa: F401
from .hooks import MultiDatasetsEvalHook  # noqa: F401
