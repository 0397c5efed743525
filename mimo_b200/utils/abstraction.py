"""Statistics: the tuple-with-arithmetic the conjugate updates are written in
(mirrors mimo/utils/abstraction.py:7-24: ``prior.nat_param + stats``, shard sums,
scalar scaling for SVI)."""
import numpy as np


def _is_seq(v):
    return isinstance(v, (list, tuple)) and not isinstance(v, np.ndarray)


class Statistics(tuple):

    def __new__(cls, items):
        return tuple.__new__(cls, items)

    @staticmethod
    def _zip(a, b, op):
        if _is_seq(a) and _is_seq(b):
            return [op(u, v) for u, v in zip(a, b)]
        return op(a, b)

    def __add__(self, other):
        return Statistics(self._zip(a, b, lambda u, v: u + v) for a, b in zip(self, other))

    def __sub__(self, other):
        return Statistics(self._zip(a, b, lambda u, v: u - v) for a, b in zip(self, other))

    def __mul__(self, scalar):
        return Statistics(scalar * item for item in self)

    __rmul__ = __mul__
