"""Drop-in replacement of the reference's ``common.quantity`` package
(quantity/common/quantity/__init__.py:1-6): the same 21 public names, backed by the
hand-written sm_100a kernels of libpq_sm100.so (see include/pq_sm100.h).  No CPU fallback."""
from .distribution_collector import DistributionCollector
from .quantizer import Quantizer
from .bit_reader import BitReader
from .utils import merge_bn, walk_dirs, tid
from .fabu_layer import Eltwise, Concat, Identity, View
from .new_quantity_op import (RightShift, Sp, BiasAdd, NewConv2d, NewAdd, NewLinear, QuanDequan,
                              TestConv, TestLinear, Quantity, DeQuantity)

from .int8_pipeline import enable_int8_pipeline, QTensor, GraphedForward   # extensions, not reference names

__all__ = ["DistributionCollector", "Quantizer", "BitReader", "merge_bn", "walk_dirs", "tid",
           "Eltwise", "Concat", "Identity", "View", "RightShift", "Sp", "BiasAdd", "NewConv2d",
           "NewAdd", "NewLinear", "QuanDequan", "TestConv", "TestLinear", "Quantity", "DeQuantity"]
