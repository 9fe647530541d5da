from . import paths  # noqa: F401
