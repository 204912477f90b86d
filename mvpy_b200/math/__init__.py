from .r2 import r2
from .pearsonr import pearsonr
from .rank import rank
from .spearmanr import spearmanr, spearmanr_d
from .accuracy import accuracy
from .roc_auc import roc_auc

__all__ = ['r2', 'pearsonr', 'rank', 'spearmanr', 'spearmanr_d', 'accuracy', 'roc_auc']
