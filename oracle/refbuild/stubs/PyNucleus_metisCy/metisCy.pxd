cimport numpy as np
ctypedef np.int32_t idx_t
ctypedef float real_t
