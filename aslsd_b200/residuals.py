"""Intensity at the jump times and residuals of a fitted MHP, on the B200.

Same results as the two O(N^2) Python loops that follow a fit in the reference:
  * ``MHP.get_intensity_at_jumps(process_path, mu, kernel_param)``
        models/hawkes/linear/mhp.py:669-690
  * ``goodness_of_fit.get_residuals_mhp(process_path, psi, mu, kernel_param)`` (what
    ``MHP.get_residuals`` returns without sampling)
        stats/residual_analysis/goodness_of_fit.py:134-236
``model`` is the reference's ``MHP``, the mirror of ``aslsd_b200.model`` or a plain dict;
``mu`` / ``kernel_param`` are the vector of baselines and the matrix of kernel parameter
vectors, as in the reference.  All sums run on the device (``aslsd_point_sums``).
"""
import numpy as np


def _bind(model, process_path, engine):
    from aslsd_b200.events import default_engine
    engine = engine or getattr(process_path, 'engine', None) or default_engine()
    engine.bind(process_path, model)
    return engine


def _x_k(mu, kernel_param, k, d):
    return np.concatenate([[float(mu[k])]] + [np.asarray(kernel_param[k][i], dtype=np.float64)
                                              for i in range(d)])


def get_intensity_at_jumps(model, process_path, mu, kernel_param, engine=None):
    """intensity[i][m] = mu_i + sum_j sum_{n <= kappa[j][i][m]} phi_ij(t^i_m - t^j_n)."""
    engine = _bind(model, process_path, engine)
    d = engine.d
    return [float(mu[i]) + engine.point_sums(i, _x_k(mu, kernel_param, i, d), 'phi').sum(axis=0)
            for i in range(d)]


def get_residuals_k(model, process_path, k, mu, kernel_param, engine=None, reference_first_residual=True):
    """r^k_m = mu_k (t^k_{m+1} - t^k_m) + sum_i [Psi_ki(m + 1) - Psi_ki(m)] with
    Psi_ki(m) = sum_{n <= kappa[i][k][m]} psi_ki(t^k_m - t^i_n).

    ``reference_first_residual`` (default) reproduces a quirk of the reference: the compensator
    value at the first type-k event is only computed when varpi[k][i][1] > 0
    (goodness_of_fit.py:192-203), so type-i events that precede the very first type-k event are
    not subtracted from the first residual; ``False`` subtracts them."""
    engine = _bind(model, process_path, engine)
    d = engine.d
    t_k = np.asarray(process_path.list_times[k], dtype=np.float64)
    kernel_part = engine.point_sums(k, _x_k(mu, kernel_param, k, d), 'psi')
    if reference_first_residual and kernel_part.shape[1] > 0:
        kernel_part[:, 0] = 0.
    poisson = float(mu[k]) * (t_k[1:] - t_k[:-1])
    return poisson + (kernel_part[:, 1:] - kernel_part[:, :-1]).sum(axis=0)


def get_residuals(model, process_path, mu, kernel_param, engine=None):
    engine = _bind(model, process_path, engine)
    return [get_residuals_k(model, process_path, k, mu, kernel_param, engine=engine)
            for k in range(engine.d)]
