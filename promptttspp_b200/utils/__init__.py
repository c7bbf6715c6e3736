from .model import remove_weight_norm_, sequence_mask  # noqa: F401
