"""AdhesionKernel (kernels/AdhesionKernel.py:12-33): Akinci-2013 adhesion spline A(r).
Constants are the product surface; the device formula lives in csrc/dfsph.cu adh_W."""
import math

import numpy as np


class AdhesionKernel:
    def __init__(self, searchR):
        self.searchR = searchR
        self.m_k = 0.007 / math.pow(searchR, 3.25)

    def Cubic_W_norm(self, r):
        r = np.float32(r)
        h = np.float32(self.searchR)
        res = np.float32(0.0)
        if r * r <= h * h and r > np.float32(0.5) * h:
            res = np.float32(self.m_k) * np.power(np.float32(-4.0) * r * r / h + np.float32(6.0) * r - np.float32(2.0) * h, np.float32(0.25))
        return res

    def Cubic_W(self, r):
        r = np.asarray(r, dtype=np.float32)
        return self.Cubic_W_norm(np.sqrt(np.float32(np.dot(r, r))))
