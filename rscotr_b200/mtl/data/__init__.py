from .build import (load_data_cfg, build_datasets, build_dataloaders, build_iteration_strategy,  # noqa: F401
                    build_multidataloader, strategies_map)
from .multi_data_loader import MultiDataLoader  # noqa: F401
