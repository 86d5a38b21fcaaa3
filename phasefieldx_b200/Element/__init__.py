from . import Elasticity, Phase_Field, Phase_Field_Fracture  # noqa: F401

__all__ = ["Elasticity", "Phase_Field", "Phase_Field_Fracture"]
