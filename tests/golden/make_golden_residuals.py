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


if __name__ == '__main__':
    main()
