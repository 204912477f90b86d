from .kfold import KFold
from .repeatedkfold import RepeatedKFold
from .stratifiedkfold import StratifiedKFold
from .repeatedstratifiedkfold import RepeatedStratifiedKFold
from .lgo import LeaveGroupsOut
from .validator import Validator
from .cross_val_score import cross_val_score

__all__ = ['KFold', 'RepeatedKFold', 'StratifiedKFold', 'RepeatedStratifiedKFold', 'LeaveGroupsOut', 'Validator', 'cross_val_score']
