"""CohesionKernel (kernels/CohesionKernel.py:12-33): Akinci-2013 cohesion spline C(r).
Constants are the product surface; the device formula lives in csrc/dfsph.cu coh_W."""
import math

import numpy as np


class CohesionKernel:
    def __init__(self, searchR):
        self.searchR = searchR
        self.m_k = 32.0 / (math.pi * math.pow(searchR, 9.0))
        self.m_c = math.pow(searchR, 6.0) / 64.0

    def Cubic_W_norm(self, r):
        r = np.float32(r)
        h = np.float32(self.searchR)
        res = np.float32(0.0)
        if r * r <= h * h:
            r3 = r * r * r
            if r > np.float32(0.5) * h:
                res = np.float32(self.m_k) * np.power(h - r, np.float32(3.0)) * r3
            else:
                res = np.float32(self.m_k) * np.float32(2.0) * np.power(h - r, np.float32(3.0)) * r3 - np.float32(self.m_c)
        return res

    def Cubic_W(self, r):
        r = np.asarray(r, dtype=np.float32)
        return self.Cubic_W_norm(np.sqrt(np.float32(np.dot(r, r))))
