import numpy as _np
import scipy.linalg as _sla
from ..numpy import _wrap


def cho_solve(c_and_lower, b, **k):
    c, lower = c_and_lower
    return _wrap(_sla.cho_solve((_np.asarray(c), lower), _np.asarray(b)))


def solve_triangular(a, b, trans=0, lower=False, unit_diagonal=False, **k):
    return _wrap(_sla.solve_triangular(_np.asarray(a), _np.asarray(b), trans=trans, lower=lower, unit_diagonal=unit_diagonal))
