import sys
sys.path.insert(0, '.')
from tests.mgpu_check import slab_parity
r = slab_parity(1, 0, nz_per_rank=32, verbose=True)
