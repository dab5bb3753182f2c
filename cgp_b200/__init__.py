"""cgp_b200 — B200-native per-cluster GP hot path of gptune/cGP."""
from .cgp import CGPRegressor  # noqa: F401
from .gpr import GaussianProcessRegressor, KNeighborsClassifier  # noqa: F401

__version__ = "0.1.0"
