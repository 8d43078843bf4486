from .build import build_backbone, build_neck, build_head, build_transformer, build_transformer_layer_sequence  # noqa
