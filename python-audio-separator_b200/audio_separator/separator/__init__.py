from .separator import Separator  # noqa: F401
