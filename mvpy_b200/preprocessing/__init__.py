from .scaler import Scaler

__all__ = ['Scaler']
