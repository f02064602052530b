"""KL-divergence threshold search on the GPU (subsystem 2).

Drop-in for the reference class (quantity/common/quantity/quantizer.py:7-176): same
constructor, ``quantize(distributions, distribution_intervals)``, ``bits`` and
``threshold_value``.  All tensors' histograms are searched by one launch sequence of
pq_kl_search_f64 (one CTA per candidate threshold per tensor) instead of 3.3 s of Python
loops per tensor; the threshold -> fractional-bit step stays on the host with the
reference's own Python expression (quantizer.py:86-90), because ``math.log(x, 2)`` is part
of the contract.
"""
import math

import numpy as np
import torch

from . import _native


class Quantizer:

    def __init__(self, tensor_list, worker_num=1, debug=False, device=None, keep_curves=False):
        self._tensor_list = tensor_list
        self._worker_num = worker_num          # compatibility only
        self._debug = debug
        self._device = device
        self._keep_curves = keep_curves
        self._bits = {}
        self._threshold_value = {}
        self._threshold_bin = {}
        self._quantized_flag = False
        self.kl_curves = None                  # fp64 CUDA [n][1920] when keep_curves

    @property
    def bits(self):
        assert self._quantized_flag, "Please use quantize() first."
        return self._bits

    @property
    def threshold_value(self):
        assert self._quantized_flag, "Please use quantize() first."
        return self._threshold_value

    @property
    def threshold_bin(self):
        assert self._quantized_flag, "Please use quantize() first."
        return self._threshold_bin

    def _stack_counts(self, distributions):
        rows = [distributions[name] for name in self._tensor_list]
        if all(isinstance(r, torch.Tensor) and r.is_cuda for r in rows):
            return torch.stack([r.to(torch.float64) for r in rows])
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: Quantizer has no CPU fallback")
        # counts are integers (or float64 group sums of integers): exact in float64
        host = np.stack([np.asarray(r.cpu() if isinstance(r, torch.Tensor) else r, dtype=np.float64)
                         for r in rows])
        return torch.from_numpy(host).to(self._device or "cuda")

    def quantize(self, distributions, distribution_intervals):
        if self._debug and self._quantized_flag:
            return
        self._quantized_flag = True
        if not self._tensor_list:
            return
        counts = self._stack_counts(distributions)
        thr, curves = _native.kl_search(counts, want_curves=self._keep_curves)
        self.kl_curves = curves
        threshold_bins = thr.cpu().tolist()                                   # one D2H, n ints
        for name, threshold_bin in zip(self._tensor_list, threshold_bins):
            # quantizer.py:86-90, verbatim arithmetic: fp32 product when the interval is np.float32
            threshold_bias = (threshold_bin + 0.5) * distribution_intervals[name]
            bit_int_d = math.ceil(math.log(threshold_bias, 2))
            bit_bra_d_8 = int(8 - 1 - bit_int_d)
            self._bits[name] = bit_bra_d_8
            self._threshold_value[name] = threshold_bias
            self._threshold_bin[name] = threshold_bin
            print("{} ".format(name), "bit:", bit_bra_d_8)
