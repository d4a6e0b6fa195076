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
import ctypes

import numpy as np

from aslsd_b200 import capi

_MASK64 = (1 << 64) - 1


def _native_state(rng):
    """(library, aslsd_pcg64) for a numpy Generator on PCG64, else None.  The native sampler
    (csrc/sampler.cu) restates that generator bit for bit; other bit generators keep the numpy
    path below, which issues the reference's calls one by one."""
    try:
        state = rng.bit_generator.state
    except AttributeError:
        return None
    if state.get('bit_generator') != 'PCG64':
        return None
    try:
        lib = capi.load_library()
    except (RuntimeError, OSError):
        return None
    inner = state['state']
    g = capi.Pcg64(inner['state'] >> 64, inner['state'] & _MASK64, inner['inc'] >> 64,
                   inner['inc'] & _MASK64, int(state['has_uint32']), int(state['uinteger']))
    return lib, g


def _store_state(rng, g):
    rng.bit_generator.state = {
        'bit_generator': 'PCG64',
        'state': {'state': (int(g.state_hi) << 64) | int(g.state_lo),
                  'inc': (int(g.inc_hi) << 64) | int(g.inc_lo)},
        'has_uint32': int(g.has_uint32), 'uinteger': int(g.uinteger)}


def sample_single_ix_stratified(strata, n_samples, rng, flat=False):
    """samples[s] = lo_s + choice(hi_s - lo_s + 1, n_samples[s], replace=False).  With a PCG64
    generator all strata are drawn by one native call (aslsd_sample_single_ix: the restatement of
    ``Generator.choice``, same values and same final state, tests/test_sampling_stream.py) - numpy
    builds arange(population) for every large stratum, the native tail shuffle only touches the
    sampled positions.  ``flat``: return (all samples in one array, samples per stratum) instead of
    one array per stratum."""
    for lo, hi in strata:
        if lo > hi:
            raise ValueError('m_inf should be smaller than m_sup')
    native = _native_state(rng) if len(strata) else None
    if native is None:
        out = [lo + rng.choice(hi - lo + 1, size=n, replace=False) for (lo, hi), n in zip(strata, n_samples)]
        if not flat:
            return out
        counts = np.array([len(a) for a in out], dtype=np.int64)
        return (np.concatenate(out).astype(np.int64) if out else np.zeros(0, dtype=np.int64)), counts
    lib, g = native
    bounds = np.ascontiguousarray(strata, dtype=np.int64).reshape(len(strata), 2)
    lo, hi = np.ascontiguousarray(bounds[:, 0]), np.ascontiguousarray(bounds[:, 1])
    n_s = np.ascontiguousarray(n_samples, dtype=np.int64)
    if np.any(n_s > hi - lo + 1):
        raise ValueError('Cannot take a larger sample than population when replace is False')
    buf = np.empty(int(n_s.sum()), dtype=np.int64)
    rc = lib.aslsd_sample_single_ix(ctypes.byref(g), len(strata), capi.ptr_i64(lo), capi.ptr_i64(hi),
                                    capi.ptr_i64(n_s), capi.ptr_i64(buf))
    if rc != 0:
        raise ValueError('Cannot take a larger sample than population when replace is False')
    _store_state(rng, g)
    if flat:
        return buf, n_s
    offs = np.concatenate([[0], np.cumsum(n_s)])
    return [buf[offs[q]:offs[q + 1]] for q in range(len(strata))]


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


def sample_double_ix_stratified(i, j, strata, n_samples, kappa, varpi, n_events, rng, flat=False):
    """All strata of a term.  With a PCG64 generator one native call draws them all
    (aslsd_sample_double_ix: the same draws as the loop below, tests/test_sampling_stream.py).
    ``flat``: return (m, n, samples per stratum) with all strata in one pair of arrays."""
    native = _native_state(rng) if len(strata) else None
    if native is None:
        out = [sample_double_ix_stratum(i, j, h_inf, h_sup, n, kappa, varpi, n_events, rng)
               for (h_inf, h_sup), n in zip(strata, n_samples)]
        if not flat:
            return out
        counts = np.array([len(s['m']) for s in out], dtype=np.int64)
        if not out:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64), counts
        return (np.concatenate([s['m'] for s in out]).astype(np.int64),
                np.concatenate([s['n'] for s in out]).astype(np.int64), counts)
    lib, g = native
    h_inf = np.ascontiguousarray([s[0] for s in strata], dtype=np.int64)
    h_sup = np.ascontiguousarray([s[1] for s in strata], dtype=np.int64)
    n_s = np.ascontiguousarray(n_samples, dtype=np.int64)
    var = np.ascontiguousarray(varpi[i][j], dtype=np.int64)
    kap = np.ascontiguousarray(kappa[j][i], dtype=np.int64)
    total = int(n_s.sum())
    m = np.empty(total, dtype=np.int64)
    n = np.empty(total, dtype=np.int64)
    rc = lib.aslsd_sample_double_ix(ctypes.byref(g), len(strata), capi.ptr_i64(h_inf), capi.ptr_i64(h_sup),
                                    capi.ptr_i64(n_s), capi.ptr_i64(var), capi.ptr_i64(kap),
                                    int(n_events[i]), capi.ptr_i64(m), capi.ptr_i64(n))
    if rc != 0:
        raise ValueError('Cannot take a larger sample than population when replace is False')
    _store_state(rng, g)
    if flat:
        return m, n, n_s
    offs = np.concatenate([[0], np.cumsum(n_s)])
    return [{'m': m[offs[q]:offs[q + 1]], 'n': n[offs[q]:offs[q + 1]]} for q in range(len(strata))]
