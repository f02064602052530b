"""Byte-identical, faster equivalent of ``json.dump(nested_int_list, f, indent=4)`` -- the
format of the weight / bias JSON files (pytorch_quantizer.py:663-669, rewriter.py:58-59,125-126).

``json.dump(..., indent=4)`` runs the pure-Python encoder (the C encoder does not indent): ~1.5 us per number, 14 s
for the 11.7 M parameters of ResNet-18 -- more than the whole GPU calibration.  The text is perfectly regular, so it
is assembled here as ONE fixed-width byte matrix with numpy and the padding bytes are dropped at the end:

  * the trailing dimensions are merged into "blocks" of >= 32 numbers; the text of a block is the same template for
    every block (brackets, commas, indentation) with a 6-byte slot ``-ddddd`` per number, unused positions zero;
  * a block is preceded / followed by one of a handful of strings (how many enclosing arrays open before / close
    after it), chosen per block from a small table by the block's index pattern;
  * the matrix [blocks][open | template | close] is flattened and its zero bytes removed.
"""
import collections
import os
import threading

import numpy as np

_MAX_ABS = 99999          # 5 digits; wider values take the plain path
_LUT = None               # tokens of -32768 .. 32767


def _dumps_plain(arr, indent):
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


def _digits(v):
    """int64 [n] with |v| <= 99999 -> uint8 [n][6]: sign, five digit positions; unused positions are 0."""
    neg = v < 0
    a = np.abs(v)
    tok = np.zeros((v.shape[0], 6), dtype=np.uint8)
    tok[:, 0] = np.where(neg, ord("-"), 0)
    for i, p in enumerate((10000, 1000, 100, 10, 1)):
        digit = (a // p) % 10
        tok[:, 1 + i] = np.where((a >= p) | (p == 1), digit + ord("0"), 0)
    return tok


def _tokens(v):
    global _LUT
    if v.size and -32768 <= int(v.min()) and int(v.max()) <= 32767:
        if _LUT is None:
            _LUT = _digits(np.arange(-32768, 32768, dtype=np.int64))
        return _LUT[v + 32768]
    return _digits(v)


def _table(strings):
    width = max(len(s) for s in strings)
    t = np.zeros((len(strings), width), dtype=np.uint8)
    for i, s in enumerate(strings):
        t[i, :len(s)] = np.frombuffer(s.encode(), dtype=np.uint8)
    return t


def _block_template(shape, depth, indent):
    """Text of one block (an array of `shape` whose brackets sit at nesting depth `depth`) as the strings that
    precede each of its numbers, plus the string that follows the last one."""
    pad = lambda d: " " * (indent * d)                            # noqa: E731
    n = int(np.prod(shape))
    m = len(shape)
    pre = []
    idx = np.array(np.unravel_index(np.arange(n), shape)).T if n else np.zeros((0, m), dtype=np.int64)
    for j in range(n):
        z = 0                                                      # trailing zero indices: arrays opening here
        while z < m and idx[j, m - 1 - z] == 0:
            z += 1
        first = m - z                                              # shallowest (relative) level that opens at j
        if z == m:                                                 # first number of the block: all m levels open
            s = "".join("[\n" + pad(depth + d + 1) for d in range(m))
        elif z == 0:
            s = ",\n" + pad(depth + m)
        else:                                                      # close z arrays, comma, open z arrays
            s = "".join("\n" + pad(depth + d) + "]" for d in range(m - 1, first - 1, -1))
            s += ",\n" + pad(depth + first) + "".join("[\n" + pad(depth + d + 1) for d in range(first, m))
        pre.append(s)
    post = "".join("\n" + pad(depth + d) + "]" for d in range(m - 1, -1, -1))
    return pre, post


def dumps_int_array_bytes(arr, indent=4):
    arr = np.asarray(arr)
    if arr.ndim == 0:
        return str(int(arr)).encode()
    if arr.size == 0 or arr.dtype.kind not in "iuf" or indent <= 0:
        return _dumps_plain(arr, indent).encode()
    v = arr.astype(np.int64).ravel()
    if int(v.max()) > _MAX_ABS or int(v.min()) < -_MAX_ABS:
        return _dumps_plain(arr, indent).encode()
    nd = arr.ndim
    m = 1
    while m < nd and int(np.prod(arr.shape[-m:])) < 32:
        m += 1
    outer, block = arr.shape[:nd - m], arr.shape[nd - m:]
    S = int(np.prod(block))
    if S > (1 << 16):                                              # (the template is built number by number)
        return _dumps_plain(arr, indent).encode()
    R = v.size // S
    depth = nd - m                                                 # nesting depth of a block's own brackets
    pad = lambda d: " " * (indent * d)                            # noqa: E731
    # a block whose last `o` outer indices are 0 opens o enclosing arrays before itself; one whose last `c` outer
    # indices are at their maximum closes c enclosing arrays after itself
    opens, closes = [], []
    for o in range(len(outer) + 1):
        first = depth - o
        s = "" if o == len(outer) else ",\n" + pad(first)          # (o == len(outer): the very first block)
        opens.append(s + "".join("[\n" + pad(d + 1) for d in range(first, depth)))
    for c in range(len(outer) + 1):
        closes.append("".join("\n" + pad(d) + "]" for d in range(depth - 1, depth - 1 - c, -1)))
    o_cls = np.zeros(R, dtype=np.int64)
    c_cls = np.zeros(R, dtype=np.int64)
    if outer:
        idx = np.unravel_index(np.arange(R), outer)
        run_o = np.ones(R, dtype=bool)
        run_c = np.ones(R, dtype=bool)
        for ax in range(len(outer) - 1, -1, -1):
            run_o &= idx[ax] == 0
            run_c &= idx[ax] == outer[ax] - 1
            o_cls += run_o
            c_cls += run_c
    pre, post = _block_template(block, depth, indent)
    t_open, t_close = _table(opens), _table(closes)
    # one row of the matrix: [open][pre_0][slot_0][pre_1][slot_1]...[post][close]
    template = [np.zeros(t_open.shape[1], dtype=np.uint8)]
    slot_at = np.empty(S, dtype=np.int64)
    at = t_open.shape[1]
    for j, s in enumerate(pre):
        template.append(np.frombuffer(s.encode(), dtype=np.uint8))
        template.append(np.zeros(6, dtype=np.uint8))
        slot_at[j] = at + len(s)
        at += len(s) + 6
    template.append(np.frombuffer(post.encode(), dtype=np.uint8))
    at += len(post)
    template.append(np.zeros(t_close.shape[1], dtype=np.uint8))
    template = np.concatenate(template)
    buf = np.empty((R, template.size), dtype=np.uint8)
    buf[:] = template
    buf[:, :t_open.shape[1]] = t_open[o_cls]
    buf[:, at:] = t_close[c_cls]
    cols = (slot_at[:, None] + np.arange(6)[None, :]).ravel()
    buf[:, cols] = _tokens(v).reshape(R, S * 6)
    flat = buf.ravel()
    return flat[flat != 0].tobytes()


def dumps_int_array(arr, indent=4):
    return dumps_int_array_bytes(arr, indent).decode()


# Arrays this process has written, so that BiasReWriter does not have to json.load the 10^7 numbers it wrote a moment
# ago (0.3 us per number, the largest part of weight_quantize after the writer above): path -> (mtime_ns, size, array).
# An entry is used only while the file still has the recorded mtime and size; at most _CACHE_LIMIT bytes are kept.
_CACHE = collections.OrderedDict()
_CACHE_LIMIT = 512 << 20
_cache_lock = threading.Lock()


def dump_int_array(arr, path, indent=4):
    arr = np.asarray(arr)
    with open(path, "wb") as f:
        f.write(dumps_int_array_bytes(arr, indent))
    st = os.stat(path)
    with _cache_lock:
        _CACHE.pop(os.path.abspath(path), None)
        if arr.nbytes <= _CACHE_LIMIT:
            _CACHE[os.path.abspath(path)] = (st.st_mtime_ns, st.st_size, arr.copy())
            total = sum(e[2].nbytes for e in _CACHE.values())
            while total > _CACHE_LIMIT:
                _, old = _CACHE.popitem(last=False)
                total -= old[2].nbytes


def written_array(path):
    """The array dump_int_array wrote to `path` in this process, or None when the file is not ours any more (or never
    was): the caller then parses the file."""
    try:
        st = os.stat(path)
    except OSError:
        return None
    with _cache_lock:
        e = _CACHE.get(os.path.abspath(path))
    if e is not None and e[0] == st.st_mtime_ns and e[1] == st.st_size:
        return e[2]
    return None

