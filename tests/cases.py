"""Small model / parameter cases shared by the golden generator and the tests.

Models are plain data (see oracle/functionals.py): ``{'d', 'kernels'}`` with
``kernels[k][i]`` a list of bases ``{'type', 'fixed'}``.  Parameter values
follow the configurations of BASELINE.md / SURVEY.md section 8d, at reduced T_f so
that the literal O(N^2) reference finishes in well under a second.
"""
import numpy as np


def exp_kernel():
    return [{'type': 'exp'}]


def gauss_kernel():
    return [{'type': 'gauss'}]


def sbf_gauss_kernel(n_basis, beta=0.5, spacing=1.0):
    """SBF kernel of fit_sbgauss1d10r.py:60-63: beta and delta fixed, omega free."""
    return [{'type': 'gauss', 'fixed': {1: beta, 2: spacing * r}}
            for r in range(n_basis)]


def square(d, make):
    return {'d': d, 'kernels': [[make() for _ in range(d)] for _ in range(d)]}


# name -> (model, mu, kernel_param[k][i], T_f)
CASES = {
    # C1: examples/fit_exp1d.ipynb cell 4
    'exp1d': (square(1, exp_kernel), [1.5], [[[0.3, 0.8]]], 150.),
    # C2: papers/grad_based_estim/experiments/fit_exp2d.py:34-51
    'exp2d': (square(2, exp_kernel), [1.5, 1.],
              [[[.2, 1.], [.6, 1.5]], [[.7, 2.], [.1, 1.3]]], 60.),
    # C3: papers/grad_based_estim/experiments/fit_gauss1d.py:31-34
    'gauss1d': (square(1, gauss_kernel), [1.5], [[[.5, .5, 3.]]], 100.),
    # all three Gaussian variables free in 2-D: exercises the cross beta/delta
    # branches (known-wrong in the reference)
    'gauss2d': (square(2, gauss_kernel), [.8, .6],
                [[[.2, .5, 1.], [.3, .7, 2.]], [[.25, .4, 1.5], [.1, .6, .5]]], 80.),
    # C4-like: sum of Gaussians with (beta, delta) fixed
    'sbf1d3': (square(1, lambda: sbf_gauss_kernel(3)), [.5],
               [[[.2, .3, .2]]], 150.),
    'sbf1d10': (square(1, lambda: sbf_gauss_kernel(10)), [.2],
                [[[.05, .1, .05, .1, .05, .1, .05, .1, .05, .1]]], 300.),
    # C5-like: multi-dimensional SBF
    'sbf2d3': (square(2, lambda: sbf_gauss_kernel(3)), [.5, .4],
               [[[.2, .1, .1], [.1, .2, .05]], [[.05, .1, .2], [.2, .1, .1]]], 100.),
    # two free exponentials per kernel (multi-basis with free beta)
    'exp1d2b': (square(1, lambda: [{'type': 'exp'}, {'type': 'exp'}]), [1.],
                [[[.2, .8, .15, 3.]]], 120.),
    # mixed fixed pattern: gaussian with only delta fixed, two bases
    'gauss1d2b': (square(1, lambda: [{'type': 'gauss', 'fixed': {2: 1.0}},
                                     {'type': 'gauss', 'fixed': {1: 0.4}}]), [.7],
                  [[[.2, .5, .25, 2.5]]], 120.),
    # mixed families inside one kernel (exponential + Gaussian basis): upsilon of unlike pairs
    # (basis_kernel_exponential.py:30-39, :89-156; basis_kernel_gaussian.py:40-50, :151-213)
    'mix1d': (square(1, lambda: [{'type': 'exp'}, {'type': 'gauss'}]), [.8],
              [[[.2, 1.2, .25, .5, 2.]]], 100.),
    'mix2d': (square(2, lambda: [{'type': 'exp'}, {'type': 'gauss', 'fixed': {1: 0.6}}]), [.6, .5],
              [[[.15, 1., .2, 1.5], [.1, 2., .15, 1.]], [[.2, .8, .1, 2.5], [.05, 1.5, .2, .5]]], 70.),
}


# Three event types: the sub-problem index k differs from both i and j in most upsilon_i_j terms,
# which never happens for d <= 2 and is the rule in the 10-dimensional configuration C5.
# Distinct parameters in every cell so that a swapped (k, i) / (k, j) slice cannot go unnoticed.
CASES_3D = {
    'exp3d': (square(3, exp_kernel), [.7, .5, .6],
              [[[.15, 1.], [.2, 1.6], [.1, 2.2]],
               [[.25, 1.3], [.1, .9], [.15, 1.9]],
               [[.05, 2.5], [.2, 1.1], [.2, 1.4]]], 45.),
    'sbf3d2': (square(3, lambda: sbf_gauss_kernel(2, beta=0.4, spacing=1.5)), [.6, .5, .7],
               [[[.15, .05], [.1, .1], [.05, .15]],
                [[.1, .15], [.2, .05], [.05, .05]],
                [[.05, .1], [.15, .1], [.1, .2]]], 45.),
    'gauss3d': (square(3, gauss_kernel), [.6, .7, .5],
                [[[.15, .5, 1.], [.2, .7, 2.], [.1, .4, 1.5]],
                 [[.2, .6, .5], [.1, .5, 2.5], [.15, .8, 1.]],
                 [[.1, .45, 2.], [.15, .65, .8], [.2, .55, 1.2]]], 40.),
}
CASES.update(CASES_3D)
NAMES_2D = sorted(set(CASES) - set(CASES_3D))   # the cases of the first golden set


# Marked models (MTLH): name -> (model incl. impacts, x_k per k, number of events per type, T_f).
# Constant baselines; linear impact 1 + <a, xi> on marks uniform in [0, 1]^q
# (basis_impact_linear.py:36-41); kernels as above.
def _linear(q=1):
    return {'type': 'linear', 'mark_dim': q}


def mtlh_model(d, make_kernel, make_impact):
    return {'d': d, 'kernels': [[make_kernel() for _ in range(d)] for _ in range(d)],
            'impacts': [[make_impact() for _ in range(d)] for _ in range(d)]}


MTLH_CASES = {
    'mtlh_exp2d': (mtlh_model(2, exp_kernel, _linear),
                   [[.8, .3, 1., .2, 1.5, .5, .7], [.6, .25, 2., .1, 1.3, .4, .9]], [60, 45], 50.),
    'mtlh_exp1d_m2': (mtlh_model(1, exp_kernel, lambda: _linear(2)),
                      [[1.1, .35, .9, .6, .3]], [80], 60.),
    'mtlh_gauss1d': (mtlh_model(1, gauss_kernel, _linear),
                     [[.9, .4, .5, 2., .8]], [70], 60.),
    'mtlh_sbf2d3': (mtlh_model(2, lambda: sbf_gauss_kernel(3), _linear),
                    [[.5, .2, .1, .1, .1, .2, .05, .6, .3], [.4, .05, .1, .2, .2, .1, .1, .2, .7]],
                    [55, 65], 60.),
    'mtlh_exp2d_const': (mtlh_model(2, exp_kernel, lambda: {'type': 'constant'}),
                         [[.8, .3, 1., .2, 1.5], [.6, .25, 2., .1, 1.3]], [50, 40], 40.),
}


# Baselines constant on time intervals (PieceConstBaseline, basis_baseline_piececonst.py): the one
# non-constant baseline the reference's exact estimator can evaluate.  x_k = [one value per interval,
# kernels, impacts].  Every sub-problem has the same breakpoints: the reference's K closures are all
# bound to the baseline of the last one (late binding, mtlh.py:1028-1077) and index its intervals
# with the caller's values.  In the two-type case the last breakpoint IS T_f, so that the interval of
# t + s = (T_f - s) + s depends on the rounding of that sum, event by event (2 of 60 and 1 of 45 events
# of this sample fall below T_f); in the one-type case the breakpoints lie inside the window and beyond it.
def _with_baselines(model, baselines):
    out = dict(model)
    out['baselines'] = baselines
    return out


MTLH_CASES_PC = {
    'mtlh_exp2d_pc': (_with_baselines(mtlh_model(2, exp_kernel, _linear),
                                      [{'type': 'piececonst', 'beta': [0., 20., 35., 50.1]},
                                       {'type': 'piececonst', 'beta': [0., 20., 35., 50.1]}]),
                      [[.9, .6, .75, .8, .3, 1., .2, 1.5, .5, .7],
                       [.5, .7, .6, .65, .25, 2., .1, 1.3, .4, .9]], [60, 45], 50.1),
    'mtlh_gauss1d_pc': (_with_baselines(mtlh_model(1, gauss_kernel, _linear),
                                        [{'type': 'piececonst', 'beta': [0., 10., 25., 40., 70.]}]),
                        [[.9, .7, 1.1, .8, .6, .4, .5, 2., .8]], [70], 60.),
}
MTLH_CASES.update(MTLH_CASES_PC)


# three event types, distinct parameters in every cell (see CASES_3D)
MTLH_CASES_3D = {
    'mtlh_exp3d': (mtlh_model(3, exp_kernel, _linear),
                   [[.7, .15, 1., .2, 1.6, .1, 2.2, .5, .7, .3],
                    [.5, .25, 1.3, .1, .9, .15, 1.9, .4, .9, .6],
                    [.6, .05, 2.5, .2, 1.1, .2, 1.4, .8, .2, .5]], [45, 38, 52], 40.),
}
MTLH_NAMES_2D = sorted(MTLH_CASES)
MTLH_CASES.update(MTLH_CASES_3D)


# stratified traces of marked models: name -> (case of MTLH_CASES, events per type, T_f,
# stratification keywords, iterations, RNG seed)
_MTLH_KW = dict(n_exact_single=20, n_samples_adaptive_single=200, nonadaptive_sample_size_single=100,
                n_samples_adaptive_double=300, nonadaptive_sample_size_double=200)
MTLH_STRAT = {
    'mtlh_exp2d': ('mtlh_exp2d', [900, 700], 400., _MTLH_KW, 4, 3),
    'mtlh_gauss1d': ('mtlh_gauss1d', [1200], 700., _MTLH_KW, 4, 5),
}


def mtlh_inputs(name):
    """Seeded event times and marks of an MTLH case."""
    model, xs, n_events, T_f = MTLH_CASES[name]
    rng = np.random.default_rng(321)
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in n_events]
    q = max(imp.get('mark_dim', 1) for row in model['impacts'] for imp in row)
    marks = [rng.uniform(0, 1, size=(n, q)) for n in n_events]
    return times, marks, T_f


def _bench_weights(n, coeff):
    """get_random_mixture_weights of the reference (utilities/useful_statistics.py:23-32), seed 1234:
    the mixture weights of bench.py's model (aslsd_b200/simulate.py)."""
    rng = np.random.default_rng(1234)
    w = rng.uniform(size=n)
    w /= np.sum(w)
    w *= coeff
    w[-1] = coeff - np.sum(w[:-1])
    return w


# The benchmark's own model (config C5: ten event types, every kernel a sum of ten Gaussians with beta
# and delta fixed, every adjacency row summing to 0.5).  Kept out of CASES: the device tests that run
# over CASES are sized for one to three event types; this one is replayed on the CPU
# (tests/test_stratified_golden.py) and has its own device test (tests/test_gpu_exact.py, c5_short).
BENCH_CASES = {
    'c5model': (square(10, lambda: sbf_gauss_kernel(10)), [.5] * 10,
                [[list(_bench_weights(10, 0.05)) for _ in range(10)] for _ in range(10)], 250.),
}


def lookup_case(name):
    return CASES[name] if name in CASES else BENCH_CASES[name]


def case_x(name):
    """Evaluation points: truth * (1 + 0.1 U(-1, 1)), default_rng(12)
    (SURVEY section 8d)."""
    model, mu, kp, T_f = lookup_case(name)
    rng = np.random.default_rng(12)
    xs = []
    for k in range(model['d']):
        x = np.concatenate([[mu[k]]] + [np.asarray(kp[k][i], float)
                                        for i in range(model['d'])])
        xs.append(x * (1 + 0.1 * rng.uniform(-1, 1, size=len(x))))
    return xs


def flatten_sums(sum_f, sum_diff_f, prefix):
    """{'phi_0': ...} -> flat dict of arrays for npz storage."""
    out = {}
    for name, val in sum_f.items():
        out['%s/f/%s' % (prefix, name)] = np.float64(val)
        dv = sum_diff_f[name]
        if isinstance(dv, list):
            out['%s/df1/%s' % (prefix, name)] = np.asarray(dv[1], dtype=float)
            out['%s/df2/%s' % (prefix, name)] = np.asarray(dv[2], dtype=float)
        else:
            out['%s/df/%s' % (prefix, name)] = np.asarray(dv, dtype=float)
    return out


def term_names(d):
    names = []
    for i in range(d):
        names += ['phi_%d' % i, 'upsilonzero_%d' % i, 'psi_%d' % i]
    for i in range(d):
        for j in range(d):
            names.append('upsilon_%d_%d' % (i, j))
    return names
