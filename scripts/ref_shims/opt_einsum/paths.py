class PathOptimizer:
    """Empty shell: the reference subclasses it at import time."""
