import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is not None and np.isscalar(low):
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = dtype
