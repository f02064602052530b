"""Model reconstruction (reference: quantity/tools/reconstruction.py:93-333).

``Reconstruction(model)`` reads feat.table / weight.table, derives the per-layer
``{weight,bias,input,output}_bit`` and swaps Conv2d / Linear / Eltwise for the simulation
operators of ``common.quantity``:
  ReconModel -> NewConv2d / NewLinear / NewAdd   (integer simulation, tcgen05 int8 kernels)
  ReconTest  -> TestConv / TestLinear / NewAdd   (fake-quant, fused QuanDequan kernel)
``Concat`` is left untouched, as in the reference (:219-238).  The rebuilt model is saved whole
with ``torch.save`` (:240, :323), so the module definitions must be importable when loading.
"""
from collections import OrderedDict

import torch

from common.quantity import (BitReader, NewAdd, NewConv2d, NewLinear, TestConv, TestLinear, merge_bn)

from ._config import load_tool_config


def _swap(root, dotted, new_module):
    parent = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        parent = getattr(parent, p)
    parent.add_module(parts[-1], new_module)


class Reconstruction(object):

    def __init__(self, model, config=None):
        self.model = model
        self._config_arg = config
        self.load_configs()

    def load_configs(self):
        self.config = load_tool_config(self._config_arg)

    def get_quantity_information(self):
        """layer name -> {weight_bit, bias_bit, output_bit, input_bit[, layer, layer_type]}
        (reference :107-172).  ``bias_bit`` is the OUTPUT bit whatever weight.table says (:138)."""
        cared = self.config["SETTINGS"]["CARE_OP_TYPE"]
        reader = BitReader(feat_table=self.config["OUTPUT"]["FEAT_BIT_TABLE"],
                           weight_table=self.config["OUTPUT"]["WEIGHT_BIT_TABLE"])
        weight_bits, bias_bits = reader.get_weight_info()
        feat_bits, infeat_bits = reader.get_feat_info()
        info = OrderedDict()
        for name, wbit in weight_bits.items():
            assert name in feat_bits, "{} not in {}".format(name, feat_bits)
            assert name in infeat_bits, "{} not in {}".format(name, infeat_bits)
            out_bit, in_bit = feat_bits[name], int(infeat_bits[name][0])
            print("name: {} weight:{} bias:{} in:{} out:{}".format(name, wbit, bias_bits[name], in_bit, out_bit))
            info.setdefault(name, {"weight_bit": wbit, "bias_bit": out_bit,
                                   "output_bit": out_bit, "input_bit": in_bit})
        for name, bit in feat_bits.items():          # parameter-free cared layers: eltwise, concat, image
            if name in info:
                continue
            info[name] = {"weight_bit": None, "bias_bit": None, "output_bit": bit,
                          "input_bit": None if name == "image" else int(infeat_bits[name][0])}
        for name, module in self.model.named_modules():
            kind = type(module).__name__
            if name in info and kind in cared:
                info[name]["layer"] = module
                info[name]["layer_type"] = kind
        return info

    def _rebuild(self, all_quantize_infor, new_model_path, make_conv, make_linear):
        for name, module in list(self.model.named_modules()):
            kind = type(module).__name__
            if kind not in ("Conv2d", "Linear", "Eltwise"):
                continue
            assert all_quantize_infor[name]["layer_type"] == kind, "layer type wrong"
            if kind == "Conv2d":
                new = make_conv(name, module, all_quantize_infor[name])
            elif kind == "Linear":
                new = make_linear(name, module, all_quantize_infor[name])
            else:
                new = NewAdd()
                new.output_bit = all_quantize_infor[name]["output_bit"]   # feat bit of the Eltwise (int8 pipeline)
            _swap(self.model, name, new)
            print("The layer change: {} ==>{} ".format(name, type(new).__name__))
        print("Model reconstruction successfully !")
        if new_model_path:
            torch.save(self.model, new_model_path)
        return self.model

    def ReconModel(self, all_quantize_infor, new_model_path):
        """Integer simulation model (reference :175-241)."""
        return self._rebuild(all_quantize_infor, new_model_path,
                             lambda n, m, q: NewConv2d(m, q), lambda n, m, q: NewLinear(m, q))

    def ReconTest(self, all_quantize_infor, new_model_path):
        """Fake-quant model: w -> qw -> dqw, bias likewise, outputs likewise (reference :243-324)."""
        return self._rebuild(all_quantize_infor, new_model_path,
                             lambda n, m, q: TestConv(n, m, q, new_model_path),
                             lambda n, m, q: TestLinear(n, m, q, new_model_path))

    def merge_bn(self):
        return merge_bn(self.model)
