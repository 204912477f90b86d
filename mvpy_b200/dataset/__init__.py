from .make_meeg_continuous import make_meeg_continuous, make_meeg_layout

__all__ = ['make_meeg_continuous', 'make_meeg_layout']
