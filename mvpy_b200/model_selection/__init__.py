from .hierarchical import Hierarchical
from .shapley import Shapley
from .hierarchical_score import hierarchical_score
from .shapley_score import shapley_score

__all__ = ['Hierarchical', 'Shapley', 'hierarchical_score', 'shapley_score']
