"""quadrotorilqr_b200 -- B200-native batched iLQR for the quadrotor-on-SE(3) problem of
nitishthatte/QuadrotorILQR: hand-written sm_100a CUDA kernels behind a C ABI
(``include/qilqr.h``), with a host-side mirror of the reference's ``ILQR`` interface."""
from .options import ConvergenceCriteria, ILQROptions, LineSearchParams
from .solver import BatchILQR, QilqrError, RESULT_DTYPE
from . import problems

__all__ = ["BatchILQR", "QilqrError", "RESULT_DTYPE", "ILQROptions", "LineSearchParams",
           "ConvergenceCriteria", "problems"]
