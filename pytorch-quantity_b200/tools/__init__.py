"""Drop-in replacement of the reference's ``tools`` package (quantity/tools/__init__.py:1-3)."""
from .pytorch_quantizer import Quantity
from .reconstruction import Reconstruction
from .rewriter import BiasReWriter

__all__ = ["Quantity", "Reconstruction", "BiasReWriter"]
