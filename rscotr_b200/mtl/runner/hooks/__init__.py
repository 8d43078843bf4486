from .evaluation import MultiDatasetsEvalHook  # noqa: F401
from .checkpoint import CheckpointHook  # noqa: F401
