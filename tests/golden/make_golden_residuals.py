"""Golden vectors of the intensity at the jumps and of the residuals, from the live reference.

    python tests/golden/make_golden_residuals.py

For a few cases of tests/cases.py the reference simulates a path (as make_golden.py does) and
evaluates ``MHP.get_intensity_at_jumps`` and ``goodness_of_fit.get_residuals_mhp`` at the case's
true parameters.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_shim  # noqa: E402
from cases import CASES  # noqa: E402
from make_golden import kernel_param_array, simulate  # noqa: E402

NAMES = ['exp2d', 'gauss1d', 'sbf2d3', 'gauss2d']


def main():
    ref = ref_shim.import_reference()
    warnings.simplefilter('ignore')
    from aslsd.stats.residual_analysis import goodness_of_fit as gof
    for name in NAMES:
        model, mu, kp, T_f = CASES[name]
        d = model['d']
        mhp = ref_shim.build_reference_model(ref, model)
        path = simulate(ref, mhp, model, mu, kp, T_f)
        mu_arr = np.asarray(mu, dtype=float)
        kp_arr = kernel_param_array(kp, d)
        out = {'T_f': np.float64(T_f), 'd': np.int64(d)}
        intensity = mhp.get_intensity_at_jumps(path, mu=mu_arr, kernel_param=kp_arr)
        residuals = gof.get_residuals_mhp(path, mhp.psi, mu_arr, kp_arr)
        for i in range(d):
            out['times/%d' % i] = np.asarray(path.list_times[i], dtype=float)
            out['intensity/%d' % i] = np.asarray(intensity[i], dtype=float)
            out['residuals/%d' % i] = np.asarray(residuals[i], dtype=float)
        np.savez_compressed(os.path.join(HERE, 'residuals_%s.npz' % name), **out)
        print(name, [len(t) for t in path.list_times], [float(np.mean(r)) for r in residuals])


MTLH_NAMES = ['mtlh_exp2d', 'mtlh_gauss1d', 'mtlh_sbf2d3']


def expected_impact(d):
    """Fixed expected-impact matrix of the golden residuals (upstream it comes from the mark
    distributions, MTLH.make_expected_impact; here it is an input like the parameters)."""
    return np.array([[1. + 0.25 * (k + 1) + 0.1 * i for i in range(d)] for k in range(d)])


def main_mtlh():
    """``goodness_of_fit.get_residuals_mtlh`` on the marked cases of tests/cases.py, with the
    reference's own MTLH object providing mu_compensator and psi."""
    ref_shim.import_reference()
    warnings.simplefilter('ignore')
    from aslsd.stats.events.process_path import ProcessPath
    from aslsd.stats.residual_analysis import goodness_of_fit as gof
    from cases import MTLH_CASES, mtlh_inputs
    from make_golden_mtlh import build_reference_mtlh
    for name in MTLH_NAMES:
        model, xs, n_events, T_f = MTLH_CASES[name]
        d = model['d']
        mtlh = build_reference_mtlh(model)
        times, marks, T_f = mtlh_inputs(name)
        path = ProcessPath(times, T_f, list_marks=marks)
        params = [mtlh.xk2matrix_params(k, np.asarray(xs[k], dtype=float)) for k in range(d)]
        mu_param = [p[0] for p in params]
        kernel_param = [p[1] for p in params]
        impact_param = [p[2] for p in params]
        E = expected_impact(d)
        residuals = gof.get_residuals_mtlh(path, mtlh.mu_compensator, mtlh.psi, E, mu_param, kernel_param,
                                           impact_param)
        out = {'T_f': np.float64(T_f), 'd': np.int64(d), 'expected_impact': E}
        for k in range(d):
            out['times/%d' % k] = np.asarray(times[k], dtype=float)
            out['residuals/%d' % k] = np.asarray(residuals[k], dtype=float)
            out['mu/%d' % k] = np.asarray(mu_param[k], dtype=float)
            for i in range(d):
                out['kernel_param/%d/%d' % (k, i)] = np.asarray(kernel_param[k][i], dtype=float)
        np.savez_compressed(os.path.join(HERE, 'residuals_%s.npz' % name), **out)
        print(name, [len(t) for t in times], [float(np.mean(r)) for r in residuals])


if __name__ == '__main__':
    if sys.argv[1:] != ['mtlh']:
        main()
    main_mtlh()
