"""Intensity at the jump times and residuals of a fitted MHP (and residuals of a fitted marked
model with constant baselines), on the B200.

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


def get_residuals_k_mtlh(model, process_path, k, mu_k, kernel_param_k, expected_impact_k, engine=None):
    """Residuals of a marked model with constant baselines, as
    ``goodness_of_fit.get_residuals_k_mtlh`` (stats/residual_analysis/goodness_of_fit.py:245-301,
    what ``MTLH.get_residuals`` returns without sampling) computes them:
        r^k_m = mu_k t^k_{m+1} - mu_k t^k_m + sum_i e_ki [Psi_ki(m + 1) - Psi_ki(m)],
    with e_ki = ``expected_impact_matrix[k][i]`` the expected impact (the reference's
    ``MTLH.make_expected_impact``) - the marks of the individual events do not enter - and without
    the first-residual quirk of the MHP version (:271-276 always evaluate the first value).
    ``mu_k``: the constant baseline of type k; ``kernel_param_k[i]``: parameters of phi_ki."""
    engine = _bind(model, process_path, engine)
    d = engine.d
    t_k = np.asarray(process_path.list_times[k], dtype=np.float64)
    x_k = np.concatenate([[float(mu_k)]] + [np.asarray(kernel_param_k[i], dtype=np.float64) for i in range(d)])
    kernel_part = engine.point_sums(k, x_k, 'psi') * np.asarray(expected_impact_k, dtype=np.float64)[:, None]
    poisson = float(mu_k) * t_k[1:] - float(mu_k) * t_k[:-1]
    return poisson + (kernel_part[:, 1:] - kernel_part[:, :-1]).sum(axis=0)


def get_residuals_mtlh(model, process_path, mu, kernel_param, expected_impact_matrix, engine=None):
    """All event types: ``mu[k]`` the constant baselines, ``kernel_param[k][i]``,
    ``expected_impact_matrix[k][i]`` as in ``goodness_of_fit.get_residuals_mtlh`` (:304-322)."""
    engine = _bind(model, process_path, engine)
    return [get_residuals_k_mtlh(model, process_path, k, np.ravel(mu[k])[0], kernel_param[k],
                                 expected_impact_matrix[k], engine=engine) for k in range(engine.d)]


def get_residuals(model, process_path, mu, kernel_param, engine=None):
    engine = _bind(model, process_path, engine)
    return [get_residuals_k(model, process_path, k, mu, kernel_param, engine=engine)
            for k in range(engine.d)]
