from .metric import Metric
from .r2 import R2, r2
from .pearsonr import Pearsonr, pearsonr
from .spearmanr import Spearmanr, spearmanr
from .score import score, reduce_

__all__ = ['Metric', 'R2', 'r2', 'Pearsonr', 'pearsonr', 'Spearmanr', 'spearmanr', 'score', 'reduce_']
