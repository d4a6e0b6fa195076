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
    hs, n_hs = distinct[order], counts[order]
    firsts = np.asarray(varpi[i][j])[hs]
    pops = n_events[i] - firsts
    m_all = np.empty(int(n_hs.sum()), dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(n_hs)])
    # A lag drawn once costs the reference one ``rng.choice(pop, 1, replace=False)``, which is a
    # single bounded draw (Floyd's algorithm with one element, nothing to shuffle:
    # numpy/random/_generator.pyx).  ``rng.integers(0, pops)`` with an array of bounds issues the
    # same bounded draws in the same order, so runs of such lags are sampled in one call and the
    # generator stays bit-identical to the reference's (tests/test_sampling_stream.py); lags
    # drawn several times take the reference's call one by one.
    multi = np.flatnonzero(n_hs != 1)
    start = 0
    for stop in list(multi) + [len(hs)]:
        if stop > start:
            m_all[offs[start:stop]] = firsts[start:stop] + rng.integers(0, pops[start:stop], dtype=np.int64)
        if stop < len(hs):
            n_h, pop = int(n_hs[stop]), int(pops[stop])
            m_all[offs[stop]:offs[stop + 1]] = firsts[stop] + rng.choice(pop, n_h, replace=bool(n_h > pop))
        start = stop + 1
    h_all = np.repeat(hs, n_hs)
    return {'m': m_all, 'n': kap[m_all] - h_all + 1}


def sample_double_ix_stratified(i, j, strata, n_samples, kappa, varpi, n_events, rng):
    return [sample_double_ix_stratum(i, j, h_inf, h_sup, n, kappa, varpi, n_events, rng)
            for (h_inf, h_sup), n in zip(strata, n_samples)]
