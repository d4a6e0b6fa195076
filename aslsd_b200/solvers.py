"""First-order solvers driving the fit loop where the reference package is not
installed (GPU box, benchmarks).  O(n_param) host arithmetic; not part of the
accelerated path.

  * learning-rate schedules   optimize/solvers/solver.py:35-73
  * ADAM.iterate              optimize/solvers/adam.py:51-68
"""
import numpy as np


def make_learning_rate(learning_rate_type='constant', **kw):
    if learning_rate_type == 'constant':
        rate_0 = kw.get('initial_learning_rate', 10 ** -2)
        return lambda t: rate_0
    if learning_rate_type == 'exponential':
        rate_0 = kw.get('initial_learning_rate', 10 ** -1)
        divider = kw.get('learning_rate_divider', 2)
        period = kw.get('learning_rate_period', 200)
        cutoff = kw.get('learning_rate_cutoff', 1. * 10 ** -3)
        return lambda t: max(rate_0 / float(divider ** (t // period)), cutoff)
    if learning_rate_type == 'linear':
        rate_0 = kw.get('initial_learning_rate', 1.)
        offset = kw.get('learning_rate_offset', 1.)
        slope = kw.get('learning_rate_slope', .1)
        return lambda t: rate_0 / (offset + slope * t)
    if learning_rate_type == 'custom':
        func = kw.get('learning_rate_custom', None)
        if func is None:
            raise ValueError("a rate function must be given in 'learning_rate_custom'")
        return func
    raise ValueError('unknown learning_rate_type %r' % (learning_rate_type,))


class ADAM:
    """x <- x - rate(t) * m_hat / (sqrt(v_hat) + epsilon) with bias-corrected
    first / second gradient moments (adam.py:51-68); no gradient clipping."""

    def __init__(self, **kwargs):
        self.momentum_1 = kwargs.get('momentum_1', 0.9)
        self.momentum_2 = kwargs.get('momentum_2', 0.999)
        self.epsilon = kwargs.get('epsilon', 10 ** -8)
        self.momentum_term_1 = 0.
        self.momentum_term_2 = 0.
        self.t = 0
        if kwargs.get('clipping', False):
            # the reference clips the gradient against a window of its past magnitudes
            # (solver.py:78-95); a fit configured that way must use the reference's own solver
            raise NotImplementedError('gradient clipping is not mirrored: pass the reference\'s ADAM')
        self.learning_rate = make_learning_rate(**kwargs)

    def iterate(self, t, x, grad):
        m = self.momentum_1 * self.momentum_term_1 + (1. - self.momentum_1) * grad
        v = self.momentum_2 * self.momentum_term_2 + (1. - self.momentum_2) * grad ** 2
        self.momentum_term_1, self.momentum_term_2 = m, v
        m_hat = m / (1. - self.momentum_1 ** (t + 1))
        v_hat = v / (1. - self.momentum_2 ** (t + 1))
        rate = self.learning_rate(t)
        self.t += 1
        return x - rate * m_hat / (np.sqrt(v_hat) + self.epsilon)
