import numpy as _np
import scipy.special as _sp
from ..numpy import _wrap


def multigammaln(a, d):
    return _wrap(_np.asarray(_sp.multigammaln(float(_np.asarray(a)), int(d))))
