"""Drop-in steps (same class names as igm/steps/__init__.py:4,9 exports)."""
from .ActivationDistanceStep import ActivationDistanceStep
from .HicEvaluationStep import HicEvaluationStep

__all__ = ["ActivationDistanceStep", "HicEvaluationStep"]
