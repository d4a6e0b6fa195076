"""Stratification state of the ASLSD estimator (host side, picklable).

Mirrors of the reference's stratification classes - same constructor keywords,
same attributes, same default geometry - so that the estimators of
``aslsd_b200.estimators`` can be used where the reference package is not
installed.  Where it is, its own objects can be passed instead through the
``single_strfs`` / ``double_strfs`` keywords of ``initialize``: the estimators
only read / write the attributes listed below.

  * ``GeneralStratification(n, **kw)``      optimize/stratifications/general_stratification.py:10-157
  * ``DoubleIxStratification(h_max, **kw)`` optimize/stratifications/double_ix_stratification.py:8-110
  * ``partition_space``                     utilities/useful_functions.py:255-270

Attributes consumed by the estimators: ``n_exact``, ``exact_domain``,
``n_adaptive_strata``, ``adaptive_strata``, ``adaptive_strata_sizes``,
``n_samples_adaptive``, ``n_adaptive_iters``, ``ema_weight``,
``adaptive_rel_alloc``, ``guess_rel_alloc``, ``n_nonadaptive_strata``,
``nonadaptive_strata``, ``nonadaptive_strata_sizes``, ``nonadaptive_abs_alloc``.
"""
import pickle

import numpy as np


def partition_space(l_bound, r_bound, n_strata, disc_type='linear'):
    """Cut the integer range [l_bound, r_bound] into n_strata consecutive
    [lo, hi] strata whose boundaries are the integer parts of a linear (or
    logarithmic) grid (useful_functions.py:255-270)."""
    if disc_type == 'log':
        grid = np.logspace(l_bound, r_bound, n_strata + 1)
    else:
        grid = np.linspace(l_bound, r_bound, n_strata + 1)
    grid = grid.astype(int)
    strata = [[grid[q], grid[q + 1] - 1] for q in range(n_strata)]
    strata[0] = [l_bound, grid[1] - 1]
    strata[-1] = [grid[-2], r_bound]
    return strata


def _sizes(strata):
    return np.array([hi - lo + 1 for lo, hi in strata], dtype=int)


class _Picklable:
    def save(self, file, **kwargs):
        with open(file, 'wb', **kwargs) as f:
            pickle.dump(self, f)

    def load(self, file, **kwargs):
        with open(file, 'rb') as f:
            other = pickle.load(f)
        self.__dict__.update(vars(other))


class GeneralStratification(_Picklable):
    """Strata of a single-index sum over ``n`` elements: an exactly evaluated
    tail, an adaptively sampled block next to it and a non-adaptive remainder
    (general_stratification.py:31-157)."""

    def __init__(self, n=0, **kwargs):
        self.n = n
        if n > 0:
            self.initialize_strata(**kwargs)

    def initialize_strata(self, **kw):
        n = self.n
        reverse = self.start_exact_reverse = kw.get('start_exact_reverse', True)
        # exact domain: the last (or first) n_exact indices
        self.exact_domain = kw.get('exact_domain', None)
        if self.exact_domain is not None:
            self.n_exact = len(self.exact_domain)
        else:
            self.n_exact = min(n, kw.get('n_exact', 1000))
            if self.n_exact > 0:
                first = n - self.n_exact if reverse else 0
                self.exact_domain = first + np.arange(self.n_exact)
        # adaptive strata
        self.n_nonadaptive_strata = kw.get('n_nonadaptive_strata', None)
        self.adaptive_strata = kw.get('adaptive_strata', None)
        if n == self.n_exact:
            self.n_adaptive_strata = 0
        elif self.adaptive_strata is not None:
            self.n_adaptive_strata = len(self.adaptive_strata)
        else:
            self.n_adaptive_strata = min(kw.get('n_adaptive_strata', 10), n - self.n_exact)
        if self.n_adaptive_strata > 0:
            n_samples_adaptive = kw.get('n_samples_adaptive', 10 ** 3)
            if isinstance(n_samples_adaptive, (list, np.ndarray)):
                # the reference raises here too (it reads the attribute before it
                # is assigned): multi-stage sampling is unreachable
                raise AttributeError("'GeneralStratification' object has no attribute "
                                     "'n_samples_adaptive'")
            self.n_adaptive_iters = kw.get('n_adaptive_iters', 1)
            if n_samples_adaptive is not None:
                n_samples_adaptive = n_samples_adaptive * np.ones(self.n_adaptive_iters)
            if self.adaptive_strata is None:
                width = 10 ** 3 * self.n_adaptive_strata
                if reverse:
                    hi = n - self.n_exact - 1
                    lo = 0 if self.n_nonadaptive_strata == 0 else max(0, hi - width)
                else:
                    lo = self.n_exact
                    hi = n - 1 if self.n_nonadaptive_strata == 0 else min(n - 1, lo + width)
                self.adaptive_strata_type = kw.get('adaptive_strata_type', 'linear')
                self.adaptive_strata = partition_space(lo, hi, self.n_adaptive_strata,
                                                       self.adaptive_strata_type)
            self.adaptive_strata_sizes = _sizes(self.adaptive_strata)
            self.total_size_adaptive = np.sum(self.adaptive_strata_sizes)
            self.n_samples_adaptive = np.minimum(n_samples_adaptive,
                                                 self.total_size_adaptive).astype(int)
            self.adaptive_sample_size = np.sum(self.n_samples_adaptive)
            self.adaptive_strata_fpop = np.array([np.sqrt(s / float(s - 1.))
                                                  for s in self.adaptive_strata_sizes])
            self.ema_weight = kw.get('ema_weight', .6)
            alloc = self.get_default_rel_alloc(
                self.n_adaptive_strata,
                alloc_type=kw.get('default_ada_rel_alloc_type', 'uniform'),
                strata_sizes=self.adaptive_strata_sizes)
            self.adaptive_rel_alloc = alloc.copy()
            self.guess_rel_alloc = alloc.copy()
        else:
            self.total_size_adaptive = 0
            self.adaptive_sample_size = 0
        # non-adaptive strata: whatever is left
        self.nonadaptive_strata = kw.get('nonadaptive_strata', None)
        sample_size = kw.get('nonadaptive_sample_size', 10 ** 4)
        self.total_size_nonadaptive = n - self.n_exact - self.total_size_adaptive
        if self.total_size_nonadaptive == 0 or sample_size == 0:
            self.n_nonadaptive_strata = 0
        elif self.nonadaptive_strata is None:
            self.n_nonadaptive_strata = min(10, self.total_size_nonadaptive)
        else:
            self.n_nonadaptive_strata = len(self.nonadaptive_strata)
        if self.n_nonadaptive_strata > 0:
            if self.nonadaptive_strata is None:
                if reverse:
                    lo = 0
                    hi = (self.adaptive_strata[0][0] - 1 if self.n_adaptive_strata > 0
                          else n - self.n_exact - 1)
                else:
                    lo = (self.adaptive_strata[-1][1] + 1 if self.n_adaptive_strata > 0
                          else self.n_exact)
                    hi = n - 1
                self.nonadaptive_strata_type = kw.get('nonadaptive_strata_type', 'linear')
                self.nonadaptive_strata = partition_space(lo, hi, self.n_nonadaptive_strata,
                                                          self.nonadaptive_strata_type)
            self.nonadaptive_strata_sizes = _sizes(self.nonadaptive_strata)
            abs_alloc = kw.get('nonadaptive_abs_alloc', None)
            if abs_alloc is None:
                per_stratum = max(1, int(sample_size / self.n_nonadaptive_strata))
                abs_alloc = per_stratum * np.ones(self.n_nonadaptive_strata, dtype=int)
            self.nonadaptive_abs_alloc = np.minimum(abs_alloc, self.nonadaptive_strata_sizes)
            self.nonadaptive_sample_size = np.sum(self.nonadaptive_abs_alloc)
        else:
            self.total_size_nonadaptive = 0
            self.nonadaptive_sample_size = 0
        self.total_eval_size = (self.n_exact + self.adaptive_sample_size
                                + self.nonadaptive_sample_size)

    def get_stratification_size(self, strata):
        return _sizes(strata)

    def get_default_rel_alloc(self, n_strata, alloc_type='uniform', strata_sizes=None, **kwargs):
        alloc = np.zeros(n_strata)
        if alloc_type == 'uniform':
            alloc[:-1] = 1. / n_strata
            alloc[-1] = 1. - np.sum(alloc)
        elif alloc_type == 'proportional':
            sizes = np.array(strata_sizes, dtype=float)
            alloc = sizes / np.sum(sizes)
            alloc[-1] = 1. - np.sum(alloc[:-1])
        return alloc


class DoubleIxStratification(_Picklable):
    """Strata of a double-index sum by index lag h: single-lag adaptive strata
    [h, h] for h < max_lag_adaptive and log-spaced non-adaptive lag ranges up to
    h_max (double_ix_stratification.py:8-110)."""

    def __init__(self, h_max=0, **kwargs):
        self.h_max = h_max
        self.n_exact = 0
        if h_max > 0:
            self.initialize_strata(**kwargs)

    def initialize_strata(self, **kw):
        self.n_adaptive_iters = kw.get('n_adaptive_iters', 1)
        self.n_samples_adaptive = kw.get('n_samples_adaptive', 10 ** 4)
        self.adaptive_rel_alloc_0 = kw.get('adaptive_rel_alloc_0', None)
        self.adaptive_rel_alloc = kw.get('adaptive_rel_alloc', None)
        self.ema_weight = kw.get('ema_weight', 0.6)
        if isinstance(self.n_samples_adaptive, (list, np.ndarray)):
            self.n_adaptive_iters = len(self.n_samples_adaptive)
        else:
            self.n_samples_adaptive = np.array([self.n_samples_adaptive])
            self.n_adaptive_iters = 1
        strata = kw.get('adaptive_strata', None)
        if strata is None:
            strata = self.get_default_adaptive_strata(kw.get('max_lag_adaptive', 30))
        self.adaptive_strata = strata
        self.n_adaptive_strata = len(strata)
        if self.adaptive_rel_alloc_0 is None:
            self.adaptive_rel_alloc_0 = self.get_default_adaptive_rel_alloc_0()
        elif len(self.adaptive_rel_alloc_0) != self.n_adaptive_strata:
            raise ValueError('Incompatible lengths between adaptive stratification and '
                             'absolute allocation')
        self.guess_rel_alloc = 0 + self.adaptive_rel_alloc_0
        self.adaptive_rel_alloc = 0 + self.adaptive_rel_alloc_0
        self.nonadaptive_strata = kw.get('nonadaptive_strata', None)
        if self.nonadaptive_strata is None:
            self.nonadaptive_strata = self.get_default_nonadaptive_strata()
        self.n_nonadaptive_strata = len(self.nonadaptive_strata)
        self.nonadaptive_abs_alloc = kw.get('nonadaptive_abs_alloc', None)
        self.nonadaptive_sample_size = kw.get('nonadaptive_sample_size', 10 ** 4)
        if (self.nonadaptive_abs_alloc is not None
                and len(self.nonadaptive_abs_alloc) != self.n_nonadaptive_strata):
            raise ValueError('Incompatible lengths between non adaptive stratification and '
                             'absolute allocation')

    def get_lag_strata_sizes(self, strata, lag_sizes, i, j):
        return [np.sum(lag_sizes[i][j][lo:hi + 1]) for lo, hi in strata]

    def get_lag_strata_finitepop(self):
        return [np.sqrt(s / float(s - 1.)) for s in self.adaptive_strata_sizes]

    def get_stratification_size(self, lag_sizes, i, j):
        self.adaptive_strata_sizes = self.get_lag_strata_sizes(self.adaptive_strata,
                                                               lag_sizes, i, j)
        self.adaptive_strata_fpop = self.get_lag_strata_finitepop()
        total = sum(self.adaptive_strata_sizes)
        for q in range(self.n_adaptive_iters):
            self.n_samples_adaptive[q] = min(self.n_samples_adaptive[q], total)
        self.nonadaptive_strata_sizes = self.get_lag_strata_sizes(self.nonadaptive_strata,
                                                                  lag_sizes, i, j)
        abs_alloc = self.nonadaptive_abs_alloc
        if abs_alloc is None:
            per_stratum = max(1, int(self.nonadaptive_sample_size / self.n_nonadaptive_strata))
            abs_alloc = per_stratum * np.ones(self.n_nonadaptive_strata, dtype=int)
        self.nonadaptive_abs_alloc = np.minimum(abs_alloc, self.nonadaptive_strata_sizes)
        self.nonadaptive_sample_size = np.sum(self.nonadaptive_abs_alloc)

    def get_default_adaptive_strata(self, max_lag_adaptive):
        return np.array([[h, h] for h in range(1, min(max_lag_adaptive, self.h_max))])

    def get_default_nonadaptive_strata(self):
        h_min = self.adaptive_strata[-1][1] + 1
        edges = np.logspace(np.log10(h_min), np.log10(self.h_max + 1), num=10)
        edges[0] = h_min
        edges[-1] = self.h_max + 1
        edges = [int(x) for x in edges]
        return [[edges[q], edges[q + 1] - 1] for q in range(len(edges) - 1)]

    def get_default_adaptive_rel_alloc_0(self):
        return (1. / self.n_adaptive_strata) * np.ones(self.n_adaptive_strata)
