"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# relative tolerance of the LSE path stated by BASELINE.json north_star
RTOL = 1e-10


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False))


def golden_times(g):
    d = int(g['d'])
    return [np.asarray(g['times/%d' % i]) for i in range(d)], float(g['T_f'])


def iter_sum_entries(sum_f, sum_diff_f):
    """Yield (label, value) for every scalar / vector of a per-term result."""
    for name in sum_f:
        yield 'f/' + name, np.float64(sum_f[name])
        dv = sum_diff_f[name]
        if isinstance(dv, list):
            yield 'df1/' + name, np.asarray(dv[1], float)
            yield 'df2/' + name, np.asarray(dv[2], float)
        else:
            yield 'df/' + name, np.asarray(dv, float)


def assert_sums_close(sum_f, sum_diff_f, golden, prefix, rtol=RTOL, atol=0.0):
    """Per-term parity gate: |got - ref| <= rtol * |ref| (+ atol)."""
    worst = 0.0
    for label, got in iter_sum_entries(sum_f, sum_diff_f):
        ref = golden['%s/%s' % (prefix, label)]
        err = np.abs(np.asarray(got) - ref)
        bound = rtol * np.abs(ref) + atol
        assert np.all(err <= bound), (
            '%s %s: got %r ref %r (rel err %.3e)' % (
                prefix, label, got, ref,
                float(np.max(err / np.maximum(np.abs(ref), 1e-300)))))
        worst = max(worst, float(np.max(err / np.maximum(np.abs(ref), 1e-300))))
    return worst


# ---- stratified golden traces (tests/golden/make_golden_strat.py) ---------------------------
def _flat_terms(est):
    f = np.array([est.sum_f[name] for name in sorted(est.sum_f)], dtype=float)
    parts = []
    for name in sorted(est.sum_diff_f):
        v = est.sum_diff_f[name]
        parts += [np.ravel(v[1]), np.ravel(v[2])] if isinstance(v, list) else [np.ravel(v)]
    return f, np.concatenate(parts).astype(float)


def _flat_allocs(est):
    out = [np.asarray(est.stratif[name].adaptive_rel_alloc, dtype=float)
           for name in sorted(est.stratif) if est.stratif[name].n_adaptive_strata > 0]
    return np.concatenate(out)


def replay_strat_golden(name, engine, mode='strict', free_running=False, rtol=RTOL, first_calls=None):
    """Teacher-forced replay of a reference ASLSD trace: at every iteration the
    estimator is evaluated at the reference's own iterate with the same seeded
    Generator; per-term sums, gradient, LSE and adaptive allocations must match.

    mode 'strict': nothing but the evaluation point is forced.  The RNG stream
        only stays aligned if every integer sample allocation is identical, which
        the final generator state certifies.
    mode 'forced': the allocation state (adaptive_rel_alloc / guess_rel_alloc) is
        also reset to the reference's before every call, so the samples are the
        reference's regardless of the allocations computed here; the computed
        allocations are compared only for groups whose strata all have a
        coefficient of variation above 1e-6.  (In a near-constant stratum the
        reference's std is the rounding residue of its summation: a last-bit
        difference in a per-sample value changes it by O(1), see DESIGN.md.)
    ``free_running``: additionally take the ADAM steps from the computed
    gradients and compare the trajectory (looser: errors feed back).
    ``first_calls``: replay only that many estimator calls of sub-problem 0 (a prefix of the trace:
    the final generator state is then not compared; the sums of a call can only match if every
    sample before it was the reference's)."""
    import aslsd_b200
    from cases import lookup_case
    g = load_golden(name)
    case = str(g['case'])
    model = lookup_case(case)[0]
    d = int(g['d'])
    times = [np.asarray(g['times/%d' % i]) for i in range(d)]
    T_f = float(g['T_f'])
    kw = {str(k): int(v) for k, v in zip(g['kw_names'], g['kw_values'])}
    n_iter = int(g['n_iter'])
    grad_alloc = bool(g['grad_alloc'])
    mirror = aslsd_b200.MHP(model)
    path = aslsd_b200.ProcessPath(times, T_f, engine=engine) if hasattr(engine, 'index_maps') \
        else _oracle_path(times, T_f)
    rng = np.random.default_rng(int(g['seed']))
    worst = 0.
    n_checked = n_groups = 0
    n_replay = n_iter if first_calls is None else min(int(first_calls), n_iter)
    for k in range(d if first_calls is None else 1):
        est = aslsd_b200.MHPStratEstimB200(is_log_lse=True, engine=engine)
        est.initialize(k, n_iter, mirror, path, **kw)
        solver = aslsd_b200.ADAM()
        lower = mirror.lower_bounds(k)
        x_free = g['k%d/x' % k][0].copy()
        group_names = [nm for nm in sorted(est.stratif) if est.stratif[nm].n_adaptive_strata > 0]
        for t in range(n_replay):
            x_t = g['k%d/x' % k][t]
            grad = est.lse_k_grad_estimate(x_t, rng=rng, grad_alloc=grad_alloc)
            f, df = _flat_terms(est)
            ref_f, ref_df = g['k%d/sum_f' % k][t], g['k%d/sum_diff_f' % k][t]
            scale = np.abs(ref_df).max()
            np.testing.assert_allclose(f, ref_f, rtol=rtol, atol=0., err_msg='%s k%d t%d sum_f' % (name, k, t))
            np.testing.assert_allclose(df, ref_df, rtol=rtol, atol=1e-12 * scale,
                                       err_msg='%s k%d t%d sum_diff_f' % (name, k, t))
            np.testing.assert_allclose(grad, g['k%d/grad' % k][t], rtol=rtol, atol=rtol * scale / T_f,
                                       err_msg='%s k%d t%d grad' % (name, k, t))
            ref_alloc = g['k%d/allocs' % k][t]
            if mode == 'strict':
                np.testing.assert_allclose(_flat_allocs(est), ref_alloc, rtol=1e-9, atol=1e-15,
                                           err_msg='%s k%d t%d allocations' % (name, k, t))
            else:
                pos = 0
                for nm in group_names:
                    strf = est.stratif[nm]
                    n_a = strf.n_adaptive_strata
                    want = ref_alloc[pos:pos + n_a]
                    mean, std, counts = est.last_alloc_stat[nm]
                    n_groups += 1
                    # a single-sample stratum has std == 0 exactly on both sides
                    if np.all((std > 1e-6 * np.abs(mean)) | (mean == 0.) | (counts == 1)):
                        n_checked += 1
                        np.testing.assert_allclose(strf.adaptive_rel_alloc, want, rtol=1e-8, atol=1e-15,
                                                   err_msg='%s k%d t%d allocation %s' % (name, k, t, nm))
                    # force the reference's state for the next call (:566-570, :740)
                    strf.adaptive_rel_alloc = want.copy()
                    pos += n_a
            worst = max(worst, float(np.max(np.abs(f - ref_f) / np.maximum(np.abs(ref_f), 1e-300))))
            if free_running:
                x_free = np.clip(solver.iterate(t, x_free, grad), lower, np.inf)
                np.testing.assert_allclose(x_free, g['k%d/x' % k][t + 1], rtol=1e-7)
        np.testing.assert_allclose(est.logged_lse[:n_replay], g['k%d/logged_lse' % k][:n_replay], rtol=rtol)
    if first_calls is None:
        assert str(rng.bit_generator.state) == str(g['rng_final']), 'RNG stream diverged from the reference'
    if mode != 'strict':
        assert n_checked >= n_groups // 3, 'too few non-degenerate allocation groups compared'
    return worst


def _oracle_path(times, T_f):
    """ProcessPath-like container whose index maps come from the oracle (CPU tests)."""
    import types
    from oracle import time_ordering as OT
    kappa, varpi = OT.kappa_varpi_closed_form(times, T_f)
    return types.SimpleNamespace(
        list_times=times, list_marks=None, T_f=T_f, d=len(times),
        n_events=np.array([len(t) for t in times]), eta=np.array([len(t) for t in times]) / T_f,
        list_times2end=[T_f - t for t in times], kappa=kappa, varpi=varpi,
        lag_sizes=OT.lag_sizes(times, kappa, varpi))


# ---- the windowed oracle on several host cores (sub-problems are independent) ----------------
def _oracle_job(args):
    model, k, x, times, T_f, ranges = args
    from oracle import exact as OE
    f, df, n_pairs = OE.exact_sums_windowed(model, k, x, times, T_f, ranges=ranges)
    return {'k/%s' % label: val for label, val in iter_sum_entries(f, df)}, n_pairs


def oracle_sums_parallel(model, ks, xs, times, T_f, ranges=None, n_proc=None):
    """[(reference dict for assert_sums_close, n_pairs)] of ``OE.exact_sums_windowed`` for the
    sub-problems ``ks``, one forked process each (the children only run numpy: they never touch
    the parent's CUDA context)."""
    import multiprocessing as mp
    jobs = [(model, k, np.asarray(x), times, T_f, ranges) for k, x in zip(ks, xs)]
    n_proc = n_proc or max(1, min(len(jobs), os.cpu_count() or 1))
    if n_proc == 1:
        return [_oracle_job(j) for j in jobs]
    with mp.get_context('fork').Pool(n_proc) as pool:
        return pool.map(_oracle_job, jobs)
