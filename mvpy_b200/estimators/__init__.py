from .ridgecv import RidgeCV
from .timedelayed import TimeDelayed
from .ridgeencoder import RidgeEncoder
from .ridgedecoder import RidgeDecoder
from .sliding import Sliding

__all__ = ['RidgeCV', 'TimeDelayed', 'RidgeEncoder', 'RidgeDecoder', 'Sliding']
