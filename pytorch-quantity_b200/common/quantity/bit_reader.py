"""Parser of ``feat.table`` / ``weight.table`` (reference: quantity/common/quantity/bit_reader.py:20-64).

File formats (the boundary between calibration, the re-writer and the simulators):
  feat.table    ``<module_name> <out_bit> <in_bit>...`` per line, first line ``image <bit>``
  weight.table  ``<param_name> <bit>`` per line, param_name ends in ``.weight`` or ``.bias``
"""
from collections import OrderedDict


def _bit(text):
    # the reference evaluates the text (bit_reader.py:31,49); a plain integer parse accepts
    # everything the writers emit and nothing else
    return int(text)


class BitReader:

    def __init__(self, feat_table=None, weight_table=None):
        self._feat_table = feat_table
        self._weight_table = weight_table

    def get_feat_info(self):
        """-> (feat_bits {layer: out_bit}, infeat_bits {layer: [in_bit strings]})."""
        assert self._feat_table
        feat_bits, infeat_bits = {}, {}
        with open(self._feat_table, "r") as f:
            for raw in f:
                fields = raw.strip().split(" ")
                if len(fields) < 2:
                    continue
                feat_bits[fields[0]] = _bit(fields[1])
                infeat_bits[fields[0]] = fields[2:]
        print("feat count:", len(feat_bits))
        return feat_bits, infeat_bits

    def get_weight_info(self):
        """-> (weight_bits, bias_bits), ordered as in the file, keyed by layer name."""
        assert self._weight_table
        weight_bits, bias_bits = OrderedDict(), OrderedDict()
        with open(self._weight_table, "r") as f:
            for raw in f:
                raw = raw.strip()
                if not raw:
                    continue
                name, text = raw.split(" ")
                if name.endswith(".weight"):
                    weight_bits[name[:-len(".weight")]] = _bit(text)
                elif name.endswith(".bias"):
                    bias_bits[name[:-len(".bias")]] = _bit(text)
                else:
                    print("Unknow layer name {}".format(name))
        print("weight count:", len(weight_bits))
        print("bias count:", len(bias_bits))
        return weight_bits, bias_bits
