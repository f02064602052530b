"""Table / JSON re-writer (reference: quantity/tools/rewriter.py:11-140).

Host-only file post-processing that finalises weight.table:
  * bias bit := output (feat) bit of the layer, bias JSON rescaled            (:38-73)
  * weight bit capped so that weight_bit + input_bit - output_bit <= MAX_SHIFT (:75-140)
The rescale is ``around(v / 2^old * 2^new).astype(int8)`` -- it WRAPS rather than saturates
(200 -> -56), exactly like the reference (quirk Q8), because the files are the contract."""
import json
import os.path as osp

import numpy as np

from common.quantity import BitReader, walk_dirs

from ._jsonio import dump_int_array, written_array


def _rescale(values, old_bit, new_bit):
    v = np.array(values, dtype=np.float32)
    v = v / 2 ** old_bit * 2 ** new_bit
    return np.around(v).astype(np.int8)


class BiasReWriter:

    def __init__(self, weight_dir, bias_dir, output_weight_dir, output_bias_dir, weight_file,
                 feat_file, max_shift_limit=None):
        self._weight_dir = weight_dir
        self._bias_dir = bias_dir
        self._output_weight_dir = output_weight_dir
        self._output_bias_dir = output_bias_dir
        self._weight_file = weight_file
        self._max_shift_limit = max_shift_limit
        self._bit_reader = BitReader(feat_table=feat_file, weight_table=weight_file)

    def get_weight_info(self):
        return self._bit_reader.get_weight_info()

    def get_feat_info(self):
        return self._bit_reader.get_feat_info()

    def _rewrite_dir(self, src_dir, dst_dir, suffix, old_bits, new_bits):
        for path in walk_dirs(src_dir, file_type=".json"):
            assert path.endswith(suffix), path
            layer = osp.basename(path)[:-len(suffix)]
            if layer not in new_bits:
                print("Can't find {} in weight table, but json file exists.".format(layer))
                continue
            values = written_array(path)                   # ours, unchanged since we wrote it: no need to parse it
            if values is None:
                with open(path, "r") as f:
                    values = json.load(f)
            dump_int_array(_rescale(values, old_bits[layer], new_bits[layer]),
                           osp.join(dst_dir, osp.basename(path)))

    def _rewrite_table(self, strip, old_bits, new_bits):
        with open(self._weight_file, "r") as f:
            lines = [ln.strip().split(" ")[:2] for ln in f if ln.strip()]
        with open(self._weight_file, "w") as f:
            for name, bit in lines:
                if name[:-strip] in old_bits:
                    bit = str(new_bits[name[:-strip]])
                f.write("{} {}\n".format(name, bit))

    def rewrite_bias_dir(self, old_bias_bits, new_bias_bits):
        self._rewrite_dir(self._bias_dir, self._output_bias_dir, ".bias.json", old_bias_bits, new_bias_bits)

    def rewrite_bias_table(self, old_bias_bits, new_bias_bits):
        self._rewrite_table(len(".bias"), old_bias_bits, new_bias_bits)

    def max_shift_limit_weight(self, feat_bits, infeat_bits, weight_bits):
        if self._max_shift_limit is None:
            return True, {}
        need_rewrite, new_weight_bits = False, {}
        for name, wbit in weight_bits.items():
            assert name in feat_bits, "{} not in {}".format(name, feat_bits)
            assert name in infeat_bits, "{} not in {}".format(name, infeat_bits)
            assert len(set(infeat_bits[name])) == 1, infeat_bits[name]
            shift = wbit + int(infeat_bits[name][0]) - feat_bits[name]
            if shift > self._max_shift_limit:
                capped = wbit - (shift - self._max_shift_limit)
                print("weight bit: {} => {}".format(wbit, capped))
                wbit, need_rewrite = capped, True
            new_weight_bits[name] = wbit
        if not need_rewrite:
            print("Nothing needs to change.")
        return need_rewrite, new_weight_bits

    def rewrite_weight_dir(self, old_weight_bits, new_weight_bits):
        self._rewrite_dir(self._weight_dir, self._output_weight_dir, ".weight.json",
                          old_weight_bits, new_weight_bits)

    def rewrite_weight_table(self, old_weight_bits, new_weight_bits):
        self._rewrite_table(len(".weight"), old_weight_bits, new_weight_bits)
