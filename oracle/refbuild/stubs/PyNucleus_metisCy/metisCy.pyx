import numpy as np
idx = np.int32
real = np.float32
