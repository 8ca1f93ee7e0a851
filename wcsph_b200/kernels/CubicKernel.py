"""CubicKernel (kernels/CubicKernel.py:12-54 of the reference): the constants are the product
surface -- the @ti.func bodies are inlined into the CUDA sweeps (csrc/engine.cuh cubic_W /
cubic_gradW).  The methods below are host-side float32 evaluations of the same formulas, for
inspection and the closed-form identity tests; they are not on the hot path."""
import math

import numpy as np


class CubicKernel:
    def __init__(self, searchR):
        self.searchR = searchR
        self.h3 = 1.0 / (searchR * searchR * searchR)
        self.m_k = 8.0 / (math.pi)
        self.m_l = 48.0 / (math.pi)

    def Cubic_W_P(self, q):
        q = np.float32(q)
        res = np.float32(0.0)
        if q <= 1.0:
            if q <= 0.5:
                qq = q * q
                res = np.float32(6.0) * qq * q - np.float32(6.0) * qq + np.float32(1.0)
            else:
                f = np.float32(1.0) - q
                res = np.float32(2.0) * f * f * f
        return res

    def Cubic_W_norm(self, v):
        return self.Cubic_W_P(np.float32(v) / np.float32(self.searchR)) * np.float32(self.m_k) * np.float32(self.h3)

    def Cubic_W(self, v):
        v = np.asarray(v, dtype=np.float32)
        return self.Cubic_W_norm(np.sqrt(np.float32(np.dot(v, v))))

    def CubicGradW(self, r):
        r = np.asarray(r, dtype=np.float32)
        res = np.zeros(3, dtype=np.float32)
        rl = np.sqrt(np.float32(np.dot(r, r)))
        q = rl / np.float32(self.searchR)
        if rl > 1.0e-5 and q <= 1.0:
            gradq = r / (rl * np.float32(self.searchR))
            c = np.float32(self.m_l * self.h3)
            if q <= 0.5:
                res = c * q * (np.float32(3.0) * q - np.float32(2.0)) * gradq
            else:
                f = np.float32(1.0) - q
                res = -c * (f * f) * gradq
        return res
