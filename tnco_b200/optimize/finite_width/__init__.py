from .optimizer import Optimizer  # noqa: F401
