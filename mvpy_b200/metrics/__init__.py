from .metric import Metric
from .r2 import R2, r2
from .pearsonr import Pearsonr, pearsonr
from .spearmanr import Spearmanr, spearmanr
from .accuracy import Accuracy, accuracy
from .roc_auc import Roc_auc, roc_auc
from .score import score, reduce_

__all__ = ['Metric', 'R2', 'r2', 'Pearsonr', 'pearsonr', 'Spearmanr', 'spearmanr', 'Accuracy', 'accuracy', 'Roc_auc', 'roc_auc', 'score', 'reduce_']
