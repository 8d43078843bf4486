from .evaluation import MultiDatasetsEvalHook  # noqa: F401
