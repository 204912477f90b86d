from .ridgecv import RidgeCV
from .timedelayed import TimeDelayed
from .ridgeencoder import RidgeEncoder
from .ridgedecoder import RidgeDecoder
from .sliding import Sliding
from .b2b import B2B
from .receptivefield import ReceptiveField
from .ridgeclassifier import RidgeClassifier
from .classifier import Classifier

__all__ = ['RidgeCV', 'TimeDelayed', 'RidgeEncoder', 'RidgeDecoder', 'Sliding', 'B2B', 'ReceptiveField', 'RidgeClassifier', 'Classifier']
