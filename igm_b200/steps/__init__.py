"""Drop-in steps (same class names as igm/steps/__init__.py:4-11 exports)."""
from .ActivationDistanceStep import ActivationDistanceStep
from .HicEvaluationStep import HicEvaluationStep
from .DamidActivationDistanceStep import DamidActivationDistanceStep, NuclDamidActivationDistanceStep
from .FishAssignmentStep import FishAssignmentStep
from .PolymerAssignmentStep import PolymerAssignmentStep
from .SpriteAssignmentStep import SpriteAssignmentStep

__all__ = ["ActivationDistanceStep", "HicEvaluationStep", "DamidActivationDistanceStep",
           "NuclDamidActivationDistanceStep", "FishAssignmentStep", "PolymerAssignmentStep", "SpriteAssignmentStep"]
