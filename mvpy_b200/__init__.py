"""mvpy_b200 -- B200-native drop-in for MVPy's cross-validated ridge hot path.

Same sklearn-style surface as the reference for that path (``estimators.{RidgeCV, TimeDelayed,
RidgeEncoder, RidgeDecoder, Sliding}``, ``preprocessing.Scaler``, ``crossvalidation.{KFold, Validator,
cross_val_score}``, ``model_selection.{Hierarchical, Shapley, hierarchical_score, shapley_score}``,
``metrics``, ``math.{r2, pearsonr}``, ``dataset.make_meeg_continuous``); the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C-ABI in ``include/mvb.h`` (``libmvb.so``).
"""
from . import _lib  # noqa: F401
from . import math  # noqa: F401
from . import metrics  # noqa: F401
from . import preprocessing  # noqa: F401
from . import estimators  # noqa: F401
from . import crossvalidation  # noqa: F401
from . import model_selection  # noqa: F401
from . import dataset  # noqa: F401

__version__ = '0.1.0'
