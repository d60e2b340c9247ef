from .phase1_mvp import Phase1MVP  # noqa: F401
