from .train import train_model, train_model_without_det_eval  # noqa: F401
from .test import test_model  # noqa: F401
from .inference import inference_one_img  # noqa: F401
