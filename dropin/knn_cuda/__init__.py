"""Drop-in for knn_cuda (unlimblue/KNN_CUDA v0.2): KNN(k, transpose_mode).forward(ref, query) -> (dist, idx)
as used at models/dvae.py:23,68 (k=4, transpose_mode=False) and :159,172 (k=32, True)."""
import torch
import torch.nn as nn

from act_b200 import ops as _ops

__version__ = "0.2"


class KNN(nn.Module):
    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        assert ref.size(0) == query.size(0), "ref.shape={} != query.shape={}".format(ref.shape, query.shape)
        with torch.no_grad():
            if not self._t:            # [B,3,N] layout: the kernel wants points-major
                ref, query = ref.transpose(1, 2), query.transpose(1, 2)
            d, i, _ = _ops.knn(ref.float().contiguous(), query.float().contiguous(), self.k)
            if not self._t:
                d, i = d.transpose(1, 2).contiguous(), i.transpose(1, 2).contiguous()
        return d, i
