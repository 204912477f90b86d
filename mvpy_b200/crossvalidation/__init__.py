from .kfold import KFold
from .validator import Validator
from .cross_val_score import cross_val_score

__all__ = ['KFold', 'Validator', 'cross_val_score']
