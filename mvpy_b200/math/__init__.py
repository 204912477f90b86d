from .r2 import r2
from .pearsonr import pearsonr

__all__ = ['r2', 'pearsonr']
