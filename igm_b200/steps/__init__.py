"""Drop-in steps (same class names as igm/steps/__init__.py:4,5,9,10 exports)."""
from .ActivationDistanceStep import ActivationDistanceStep
from .HicEvaluationStep import HicEvaluationStep
from .DamidActivationDistanceStep import DamidActivationDistanceStep, NuclDamidActivationDistanceStep

__all__ = ["ActivationDistanceStep", "HicEvaluationStep", "DamidActivationDistanceStep",
           "NuclDamidActivationDistanceStep"]
