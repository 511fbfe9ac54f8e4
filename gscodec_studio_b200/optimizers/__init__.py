from .selective_adam import GroupedAdam, SelectiveAdam  # noqa: F401
