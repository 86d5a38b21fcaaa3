from . import Phase_Field_Fracture  # noqa: F401

__all__ = ["Phase_Field_Fracture"]
