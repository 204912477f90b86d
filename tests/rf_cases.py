"""ReceptiveField fixture cases shared by the oracle and GPU tests (same table as oracle/gen_golden.py RF_CASES)."""
import torch

RF_CASES = {
    'single': ((-0.04, 0.06, 100), dict(alpha=1.0)),
    'kfold': ((-0.04, 0.06, 100), dict(alpha=torch.logspace(-1, 5, 7), reg_cv=5)),
    'loo': ((-0.04, 0.06, 100), dict(alpha=torch.logspace(-1, 5, 7), reg_cv='LOO')),
    'laplacian': ((-0.03, 0.04, 100), dict(alpha=[0.1, 10.0, 1000.0], reg_type='laplacian', reg_cv=3, patterns=True)),
    'noedge': ((-0.05, 0.03, 100), dict(alpha=2.0, edge_correction=False, fit_intercept=False, reg_type=['laplacian', 'ridge'])),
    'causal': ((0.02, 0.08, 100), dict(alpha=0.5, patterns=True)),
    'anticausal': ((-0.08, -0.02, 100), dict(alpha=3.0, reg_type=('ridge', 'laplacian'))),
}

RF_REG_CASES = [(3, 4, 'laplacian'), (3, 4, ('ridge', 'laplacian')), (3, 4, ('laplacian', 'ridge')), (1, 5, 'laplacian'),
                (4, 1, 'laplacian'), (2, 2, 'laplacian'), (3, 4, 'ridge'), (5, 3, ['laplacian', 'laplacian'])]
