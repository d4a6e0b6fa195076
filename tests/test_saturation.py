"""Saturated values of the single-index terms (aslsd_b200/saturation.py): the constant handed to
the sampled-lag kernel is, bit for bit, what the reference's own closures return for every sample
beyond the threshold; the mirror model's closures are the reference's."""
import numpy as np
import pytest

from cases import CASES, case_x

NAMES = ['exp1d', 'exp2d', 'gauss1d', 'gauss2d', 'sbf2d3', 'sbf1d10', 'exp1d2b', 'gauss1d2b', 'exp3d', 'sbf3d2']


@pytest.mark.parametrize('name', NAMES)
def test_mirror_closures_saturate_at_the_threshold(name):
    import aslsd_b200
    from aslsd_b200 import saturation
    model = aslsd_b200.MHP(CASES[name][0])
    desc = model.desc
    rng = np.random.default_rng(3)
    for k in range(desc.d):
        x = np.asarray(case_x(name)[k], dtype=float)
        for i in range(desc.d):
            lo, hi = desc.interval_map[k][i]
            for kind in (0, 1):
                sat = saturation.term_saturation(model, desc, k, i, kind, x)
                assert sat is not None
                sat_t, v_sat = sat
                t = sat_t * (1. + rng.uniform(0., 50., size=400))
                t[0] = sat_t
                if kind == 0:
                    vals = x[0] * model.psi[k][i](t, x[lo:hi])
                else:
                    vals = model.upsilon[i][i][k](t, 0., x[lo:hi], x[lo:hi])
                assert np.all(vals == v_sat), (name, k, i, kind)
                # ... and not yet well inside the threshold (the constant is not applied too early)
                near = model.psi[k][i](np.array([0.2 * sat_t]), x[lo:hi]) * x[0] if kind == 0 else \
                    model.upsilon[i][i][k](np.array([0.2 * sat_t]), 0., x[lo:hi], x[lo:hi])
                assert near[0] != v_sat


@pytest.mark.reference
@pytest.mark.parametrize('name', NAMES)
def test_closures_and_constants_equal_the_references(name):
    """Live reference: its psi / upsilon closures against the mirror's on random arguments (bitwise),
    and term_saturation fed with the reference model itself."""
    import ref_shim
    import aslsd_b200
    from aslsd_b200 import saturation
    ref = ref_shim.import_reference()
    plain = CASES[name][0]
    mhp = ref_shim.build_reference_model(ref, plain)
    mirror = aslsd_b200.MHP(plain)
    desc = mirror.desc
    rng = np.random.default_rng(5)
    t = np.concatenate([rng.uniform(0., 30., size=300), rng.uniform(30., 3000., size=100)])
    for k in range(desc.d):
        x = np.asarray(case_x(name)[k], dtype=float)
        for i in range(desc.d):
            lo, hi = desc.interval_map[k][i]
            np.testing.assert_array_equal(mirror.psi[k][i](t, x[lo:hi]), mhp.psi[k][i](t, x[lo:hi]))
            np.testing.assert_array_equal(mirror.upsilon[i][i][k](t, 0., x[lo:hi], x[lo:hi]),
                                          mhp.upsilon[i][i][k](t, 0., x[lo:hi], x[lo:hi]))
            for kind in (0, 1):
                a = saturation.term_saturation(mhp, desc, k, i, kind, x)
                b = saturation.term_saturation(mirror, desc, k, i, kind, x)
                assert a == b and a is not None
                far = a[0] * (1. + rng.uniform(0., 20., size=200))
                vals = x[0] * mhp.psi[k][i](far, x[lo:hi]) if kind == 0 else \
                    mhp.upsilon[i][i][k](far, 0., x[lo:hi], x[lo:hi])
                assert np.all(vals == a[1])
