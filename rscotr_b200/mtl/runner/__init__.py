from .iter_runner import IterBasedRunner  # noqa: F401
from .hooks import MultiDatasetsEvalHook, CheckpointHook  # noqa: F401
