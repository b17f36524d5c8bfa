from .cuda_lib import lib, check, stream_ptr, PRO_RMSNORM, PRO_PLAIN, EPI_STORE, EPI_RESIDUAL, EPI_SWIGLU
from .q40 import DeviceQ40, repack_q40, gemv_q40
