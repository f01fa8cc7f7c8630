from .ngp import NGPRadianceField, trunc_exp

__all__ = ["NGPRadianceField", "trunc_exp"]
