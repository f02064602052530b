"""Byte-identical, faster equivalent of ``json.dump(nested_int_list, f, indent=4)`` -- the
format of the weight / bias JSON files (pytorch_quantizer.py:663-669, rewriter.py:58-59,125-126)."""
import numpy as np


def dumps_int_array(arr, indent=4):
    arr = np.asarray(arr)
    if arr.ndim == 0:
        return str(int(arr))

    def rec(a, level):
        pad_in = " " * (indent * (level + 1))
        pad_out = " " * (indent * level)
        if a.shape[0] == 0:
            return "[]"
        sep = ",\n" + pad_in
        if a.ndim == 1:
            body = sep.join(map(str, a.tolist()))
        else:
            body = sep.join(rec(sub, level + 1) for sub in a)
        return "[\n" + pad_in + body + "\n" + pad_out + "]"

    return rec(arr, 0)


def dump_int_array(arr, path, indent=4):
    with open(path, "w") as f:
        f.write(dumps_int_array(arr, indent))
