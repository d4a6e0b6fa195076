"""Host-side index sampling of the ASLSD estimator.

The stochastic estimates match the reference only if the numpy ``Generator`` is
consumed by exactly the same sequence of calls, so these functions issue the
same ``rng.choice`` calls, in the same order and with the same arguments, as

  * ``sample_single_ix_stratified``   utilities/useful_statistics.py:38-60
  * ``sample_double_ix_stratum``      utilities/useful_statistics.py:73-109
  * ``sample_double_ix_stratified``   utilities/useful_statistics.py:112-127

(one ``choice`` without replacement per single-lag stratum; for a lag range one
``choice`` with replacement for the lags, then one ``choice`` per distinct lag
in first-occurrence order).  Sampling stays on the host by design (BASELINE.json
north_star); the device consumes the flat (m, n) arrays these return.
"""
import numpy as np


def sample_single_ix_stratified(strata, n_samples, rng):
    """samples[s] = lo_s + choice(hi_s - lo_s + 1, n_samples[s], replace=False)."""
    out = []
    for (lo, hi), n in zip(strata, n_samples):
        if lo > hi:
            raise ValueError('m_inf should be smaller than m_sup')
        out.append(lo + rng.choice(hi - lo + 1, size=n, replace=False))
    return out


def sample_double_ix_stratum(i, j, h_inf, h_sup, n_samples, kappa, varpi, n_events, rng):
    """Pairs (m, n) of a type-i target and a type-j source at index lag h in
    [h_inf, h_sup]: m >= varpi[i][j][h], n = kappa[j][i][m] - h + 1."""
    kap = kappa[j][i]
    if h_inf == h_sup:
        first = varpi[i][j][h_inf]
        m = first + rng.choice(n_events[i] - first, n_samples, replace=False)
        return {'m': m, 'n': kap[m] - (h_inf - 1)}
    lags = h_inf + rng.choice(h_sup + 1 - h_inf, size=n_samples)
    # distinct lags in order of first occurrence, with their multiplicities
    distinct, first_pos, counts = np.unique(lags, return_index=True, return_counts=True)
    order = np.argsort(first_pos, kind='stable')
    list_m, list_n = [], []
    for h, n_h in zip(distinct[order], counts[order]):
        first = varpi[i][j][h]
        pop = n_events[i] - first
        m = first + rng.choice(pop, n_h, replace=bool(n_h > pop))
        list_m.append(m)
        list_n.append(kap[m] - h + 1)
    return {'m': np.concatenate(list_m), 'n': np.concatenate(list_n)}


def sample_double_ix_stratified(i, j, strata, n_samples, kappa, varpi, n_events, rng):
    return [sample_double_ix_stratum(i, j, h_inf, h_sup, n, kappa, varpi, n_events, rng)
            for (h_inf, h_sup), n in zip(strata, n_samples)]
