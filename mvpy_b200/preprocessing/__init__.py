from .scaler import Scaler
from .labelbinariser import LabelBinariser

__all__ = ['Scaler', 'LabelBinariser']
