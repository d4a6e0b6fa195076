"""Driver of tests/test_plan_and_assembly_host.py (subprocess, sanitizer runtimes preloaded, library
built against the do-nothing CUDA runtime).

Planning and assembly are host code: which parameters and event streams enter which row of the
raw accumulator array (csrc/plan.cu, exact.cu) and how the rows combine into the packed per-term
sums (csrc/plan.h emit_term).  Here the rows of the plan are read back (aslsd_exact_plan_rows),
their raw components are computed with numpy straight from the component definitions of
csrc/basis_math.cuh - the part the device kernels do - and handed to aslsd_exact_assemble_batch;
the packed sums must equal the oracle's (the literal restatement of the reference, pinned to its
golden vectors).  That covers what depends on the number of event types without a GPU: models
with one to five types."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402
from scipy.special import ndtr  # noqa: E402

from aslsd_b200 import capi  # noqa: E402

capi.LIB_PATH = sys.argv[1]
capi._lib = None
import aslsd_b200  # noqa: E402
from aslsd_b200.engine import unpack_terms  # noqa: E402
from cases import CASES, case_x  # noqa: E402
from oracle import exact as OE  # noqa: E402

(UPS_EXP, UPS_GAUSS, PHI_EXP, PHI_GAUSS, PSI_EXP, PSI_GAUSS, UPS0_EXP, UPS0_GAUSS, UPS_EG, UPS_GE, UPS0_EG,
 UPS0_GE) = range(12)
C_G, C_B1, C_D1, C_B2, C_D2, C_SW, C_SB, C_SD = range(8)
N_COMP = 8


def pairs(t_tgt, t_src, same):
    """All (target, source) pairs with the source strictly earlier (by index inside one type)."""
    if same:
        m, n = np.tril_indices(len(t_tgt), k=-1)
    else:
        hi = np.searchsorted(t_src, t_tgt, side='left')
        m = np.repeat(np.arange(len(t_tgt)), hi)
        n = np.concatenate([np.arange(h) for h in hi]) if len(hi) else np.zeros(0, dtype=int)
    return m, n


def pdf(x, mean, std):
    return np.exp(-0.5 * ((x - mean) / std) ** 2) / (std * np.sqrt(2. * np.pi))


def ups_components(family, t, s, b1, d1, b2, d2, bo, dlo, mask, w=1.):
    """Raw components of one upsilon basis pair at (t = time to the end, s = lag): basis_math.cuh
    eval_ups_exp / eval_ups_gauss."""
    out = np.zeros(N_COMP)
    S = lambda v: np.sum(w * v)          # w: weight of every pair (marked models), 1 otherwise
    if family in (UPS_EXP, UPS0_EXP):
        bsum = b1 + b2
        pr = b1 * b2 / bsum
        e, E = np.exp(-b2 * s), np.exp(-bsum * t)
        out[C_G] = S(e * (1. - E))
        if mask & (1 << C_B1):
            out[C_B1] = S(e * ((b2 / bsum) ** 2 * (1. - E) + pr * t * E))
        if mask & (1 << C_B2):
            out[C_B2] = S(e * ((b1 / bsum) ** 2 * (1. - E) + pr * ((t + s) * E - s)))
        eo, E2 = np.exp(-bo * s), np.exp(-2. * bo * t)
        if mask & (1 << C_SW):
            out[C_SW] = S(eo * (1. - E2))
        if mask & (1 << C_SB):
            out[C_SB] = S(0.5 * eo * (1. - bo * s + E2 * (bo * (2. * t + s) - 1.)))
        return out
    if family in (UPS_EG, UPS0_EG, UPS_GE, UPS0_GE):
        return mixed_components(family in (UPS_EG, UPS0_EG), t, s + 0. * t, b1, d1, b2, d2, mask, w)
    sig2 = b1 * b1 + b2 * b2
    sig = np.sqrt(sig2)
    dsv = (b2 * b2 * d1 + b1 * b1 * (d2 - s)) / sig2
    b_s = b1 * b2 / sig
    lo, hi = -dsv / b_s, (t - dsv) / b_s
    P, J = pdf(s, d2 - d1, sig), ndtr(hi) - ndtr(lo)
    p_lo, p_hi = pdf(lo, 0., 1.), pdf(hi, 0., 1.)
    out[C_G] = S(P * J)
    quad = (s - d2 + d1) ** 2 / sig2 - 1.
    # the beta / delta branches as the kernels evaluate them (the reference's closed forms,
    # basis_kernel_gaussian.py:65-145, known-wrong branches included)
    if mask & (1 << C_B1):
        c = 2. * b1 * (d2 - d1 - s) / sig2
        br = -(c - dsv / b1) * p_lo + (c + (t - dsv) / b1) * p_hi
        out[C_B1] = S(P * ((b1 / sig2) * quad * J - (b2 / (b1 * sig2)) * br))
    if mask & (1 << C_B2):
        c = 2. * b2 * (d1 - d2 + s) / sig2
        br = -(c + dsv / b2) * p_lo + (c - (t - dsv) / b2) * p_hi
        out[C_B2] = S(P * ((b2 / sig2) * quad * J - (b1 / (b2 * sig2)) * br))
    if mask & (1 << C_D1):
        out[C_D1] = S(P * (((d2 - s - d1) / sig2) * J - (b2 / (b1 * sig2)) * (p_hi - p_lo)))
    if mask & (1 << C_D2):
        out[C_D2] = S(P * (((d1 + s - d2) / sig2) * J - (b1 / (b2 * sig2)) * (p_hi - p_lo)))
    if mask & ((1 << C_SW) | (1 << C_SB) | (1 << C_SD)):
        # own-basis forms behind the derivatives of a same-type term (:502-525); d / d omega with
        # the reference's limits without the s / 2 shift (:507-508)
        r2 = np.sqrt(2.)
        P0 = pdf(s, 0., r2 * bo)
        if mask & (1 << C_SW):
            out[C_SW] = S(2. * P0 * (ndtr(r2 * (t - dlo) / bo) - ndtr(-r2 * dlo / bo)))
        a, e = -r2 * (dlo - 0.5 * s) / bo, r2 * (t - dlo + 0.5 * s) / bo
        I_ae, pdf_a, pdf_e = ndtr(e) - ndtr(a), pdf(a, 0., 1.), pdf(e, 0., 1.)
        if mask & (1 << C_SB):
            br = -(r2 * (-dlo + 0.5 * s) / bo ** 2) * pdf_a + (r2 * (t - dlo + 0.5 * s) / bo ** 2) * pdf_e
            out[C_SB] = S(P0 * (((s * s - 2. * bo * bo) / (2. * bo ** 3)) * I_ae - br))
        if mask & (1 << C_SD):
            out[C_SD] = S(-(r2 / bo) * P0 * (pdf_e - pdf_a))
    return out


def mixed_components(exp_first, t, s, b1, d1, b2, d2, mask, w=1.):
    """Exponential x Gaussian (exp_first) and Gaussian x exponential pairs: eval_ups_eg / eval_ups_ge.
    Slot C_SW holds the value under the reference's overflow guard (np.where(arg > 10**2, 0., ...)),
    which its derivatives apply sample by sample and its value does not."""
    out = np.zeros(N_COMP)
    S = lambda v: np.sum(w * v)
    std = lambda x: pdf(x, 0., 1.)
    if exp_first:
        be, bg, dg = b1, b2, d2
        d_s = dg - s - be * bg * bg
        arg = -be * (dg - s - 0.5 * be * bg * bg)
        e = np.exp(arg)
        lo, hi = -d_s / bg, (t - d_s) / bg
        J = ndtr(hi) - ndtr(lo)
        out[C_G] = S(e * J)
        keep = ~(arg > 100.)
        p_lo, p_hi, be_e = std(lo), std(hi), be * e
        if mask & (1 << C_SW):
            out[C_SW] = S(np.where(keep, e * J, 0.))
        if mask & (1 << C_B1):
            out[C_B1] = S(np.where(keep, be_e * ((1. / be - dg + s + 0.5 * be * bg * bg) * J + bg * (p_hi - p_lo)), 0.))
        if mask & (1 << C_B2):
            pdf_term = -(be + (dg - s) / bg ** 2) * p_lo + (be + (dg - s - t) / bg ** 2) * p_hi
            out[C_B2] = S(np.where(keep, be_e * (bg * be * be * J + pdf_term), 0.))
        if mask & (1 << C_D2):
            out[C_D2] = -S(np.where(keep, be_e * (be * J + (p_hi - p_lo) / bg), 0.))
        return out
    be, bg, dg = b2, b1, d1
    d0 = dg - be * bg * bg
    arg = -be * (dg + s - 0.5 * be * bg * bg)
    e = np.exp(arg)
    lo, hi = -d0 / bg, (t - d0) / bg
    J = ndtr(hi) - ndtr(lo)
    out[C_G] = S(e * J)
    if mask & (1 << C_SW):
        out[C_SW] = S(np.where(arg > 100., 0., e * J))
    be_e = be * e
    if mask & ((1 << C_B1) | (1 << C_D1)):
        d1s = dg + s - be * bg * bg        # the reference shifts these limits by s (:176, :186)
        lo1, hi1 = -d1s / bg, (t - d1s) / bg
        J1, p_lo1, p_hi1 = ndtr(hi1) - ndtr(lo1), std(lo1), std(hi1)
        if mask & (1 << C_B1):
            pdf_term = -(be + dg / bg ** 2) * p_lo1 + (be - (t - dg) / bg ** 2) * p_hi1
            out[C_B1] = S(be_e * (bg * be * be * J1 + pdf_term))
        if mask & (1 << C_D1):
            out[C_D1] = S(be_e * (-be * J1 - (p_hi1 - p_lo1) / bg))
    if mask & (1 << C_B2):
        out[C_B2] = S(be_e * ((1. / be - dg - s + 0.5 * be * bg * bg) * J + bg * (std(hi) - std(lo))))
    return out


FAR = 1e7   # stands for t = infinity: every e^{-c t}, pdf(upper limit) and 1 - cdf(upper limit) is exactly 0


def row_components(meta, par, times, T_f, with_derivatives, truncated=False, weights=None):
    """truncated: as the kernels traverse the row - only lags inside [s_lo, s_hi], and the limits
    at infinity for targets whose time to T_f is at least t_int.  weights (marked models): the
    vectors of aslsd_set_weights; the row sums w_a[m] w_b[n] x component with the vectors named in
    meta[5:7]."""
    family, tgt, src, same, mask = (int(v) for v in meta[:5])
    w_a = w_b = None
    if weights is not None:
        wa, wb = int(meta[5]), int(meta[6])
        assert (wa < 0 or weights[wa][0] == tgt) and (wb < 0 or weights[wb][0] == src)
        w_a = weights[wa][1] if wa >= 0 else None
        w_b = weights[wb][1] if wb >= 0 else None

    def pair_w(m, n=None):
        w = 1.
        if w_a is not None:
            w = w * w_a[m]
        if w_b is not None and n is not None:
            w = w * w_b[n]
        return w
    b1, d1, b2, d2, bo, dlo = par[:6]
    s_lo, s_hi, t_int = par[6:9]
    m_lo, m_hi = int(par[9]), int(par[10])
    t_tgt, t_src = times[tgt], times[src]
    out = np.zeros(N_COMP)
    block = np.zeros(len(t_tgt), dtype=bool)     # the targets the call was asked for
    block[m_lo:m_hi] = True

    def to_end(t_m):
        t = T_f - t_m
        return np.where(t >= t_int, FAR, t) if truncated else t

    def window(m, n):
        keep = block[m]
        if truncated:
            s = t_tgt[m] - t_src[n]
            keep &= (s >= s_lo) & (s <= s_hi)
        return m[keep], n[keep]
    if family in (UPS_EXP, UPS_GAUSS, UPS_EG, UPS_GE):
        m, n = window(*pairs(t_tgt, t_src, same))
        return ups_components(family, to_end(t_tgt[m]), t_tgt[m] - t_src[n], b1, d1, b2, d2, bo, dlo, mask,
                              pair_w(m, n))
    if family in (UPS0_EXP, UPS0_GAUSS, UPS0_EG, UPS0_GE):
        return ups_components(family, to_end(t_tgt[block]), 0., b1, d1, b2, d2, bo, dlo, mask,
                              pair_w(np.flatnonzero(block)))
    if family in (PHI_EXP, PHI_GAUSS):
        m, n = window(*pairs(t_tgt, t_src, same))
        s = t_tgt[m] - t_src[n]
        w = pair_w(m, n)
        if family == PHI_EXP:
            e = np.exp(-b1 * s)
            out[0] = np.sum(w * e)
            out[1] = np.sum(w * e * (1. - b1 * s))
        else:
            P, z = pdf(s, d1, b1), (s - d1) / b1
            out[0] = np.sum(w * P)
            out[1] = np.sum(w * P * (z * z - 1.))
            out[2] = np.sum(w * P * (s - d1) / (b1 * b1))
        return out
    if family in (PSI_EXP, PSI_GAUSS):
        t = to_end(t_tgt[block])
        w = pair_w(np.flatnonzero(block))
        if family == PSI_EXP:
            E = np.exp(-b1 * t)
            out[0] = np.sum(w * (1. - E))
            out[1] = np.sum(w * t * E)
        else:
            hi, lo = (t - d1) / b1, -d1 / b1
            out[0] = np.sum(w * (ndtr(hi) - ndtr(lo)))
            out[1] = np.sum(w * ((d1 - t) * pdf(hi, 0., 1.) - d1 * pdf(lo, 0., 1.)))
            out[2] = np.sum(w * (pdf(hi, 0., 1.) - pdf(lo, 0., 1.)))
        return out
    raise AssertionError('family %d is not covered by this check' % family)


def larger_model(name):
    """Models that are not among the golden cases: five event types (exponential, fixed-shape sums of
    Gaussians) and three / four types with an exponential and a Gaussian basis in every kernel;
    different parameters in every cell."""
    from cases import exp_kernel, sbf_gauss_kernel, square
    d = int(name[3])        # 'exp5d', 'sbf5d2', 'mix3d': the digit after the family
    make = {'exp': exp_kernel, 'sbf': lambda: sbf_gauss_kernel(2, beta=0.45, spacing=1.2),
            'mix': lambda: [{'type': 'exp'}, {'type': 'gauss'}]}[name[:3]]
    model = square(d, make)
    rng = np.random.default_rng(19)
    cell = {'exp': lambda: [rng.uniform(.02, .15), rng.uniform(.7, 2.5)],
            'sbf': lambda: [rng.uniform(.02, .1), rng.uniform(.02, .1)],
            'mix': lambda: [rng.uniform(.03, .12), rng.uniform(.8, 2.), rng.uniform(.03, .12), rng.uniform(.4, .8),
                            rng.uniform(.5, 2.)]}[name[:3]]
    xs = [np.array([rng.uniform(.3, .9)] + [v for _ in range(d) for v in cell()]) for _ in range(d)]
    return model, xs


def check(name, engine, with_derivatives):
    if name == 'C5':
        # the benchmark's model itself (ten event types, ten fixed-shape Gaussians per kernel) at its
        # own evaluation point, on a short uniform path
        from aslsd_b200 import simulate
        model, _, _, xs_given = simulate.config_c5(d=10, n_basis=10, T_f=5., seed=1234)
    else:
        model, xs_given = (CASES[name][0], None) if name in CASES else larger_model(name)
    d = model['d']
    rng = np.random.default_rng(7)
    T_f = 25.
    times = [np.sort(rng.uniform(0, T_f, size=(40 + 11 * i) if d <= 5 else 14 + i)) for i in range(d)]
    engine.set_events(times, T_f)
    engine.set_model(model)
    xs = case_x(name) if xs_given is None else xs_given
    ks = list(range(d))
    # plan the batch (the launches are no-ops under the stub runtime), read the rows back
    scratch = np.zeros(engine.exact_raw_len_batch(d))
    engine.exact_raw_device_batch(ks, xs, scratch.ctypes.data, True, with_derivatives)
    meta, par = engine.plan_rows()
    n_rows = len(meta)
    assert n_rows * N_COMP == engine.exact_raw_used(), (n_rows, engine.exact_raw_used())
    raw = np.zeros((n_rows, N_COMP))
    for r in range(n_rows):
        raw[r] = row_components(meta[r], par[r], times, T_f, with_derivatives)
    packed = engine.exact_assemble_batch(ks, xs, raw.ravel(), True, with_derivatives)
    # the linear map of the device assembly (CSR table of symbolic omega factors + omega tables,
    # exact.cu build_assembly_host) applied on the host in the kernel's order of operations: bit for
    # bit the packed sums of the host assembly
    mapped = engine.exact_assemble_map_batch(ks, xs, raw.ravel(), True, with_derivatives)
    for a, b in zip(packed, mapped):
        np.testing.assert_array_equal(b, a, err_msg=name + ' device map')
    worst = 0.
    for k in ks:
        sum_f, sum_diff_f = unpack_terms(engine.desc, k, packed[k])
        # (ten event types: the vectorised windowed form of the oracle, equal to the literal one to
        # 1e-13 on every golden case, instead of its per-target Python loops)
        oracle = OE.exact_sums_windowed if d > 5 else OE.exact_sums_literal
        o_f, o_df = oracle(model, k, xs[k], times, T_f)[:2]
        for term in o_f:
            assert abs(sum_f[term] - o_f[term]) <= 1e-11 * abs(o_f[term]), (name, k, term, sum_f[term], o_f[term])
            worst = max(worst, abs(sum_f[term] - o_f[term]) / abs(o_f[term]))
            if not with_derivatives:
                continue
            got, want = sum_diff_f[term], o_df[term]
            for a, b in (zip(got[1:], want[1:]) if isinstance(want, list) else [(got, want)]):
                np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12 * np.abs(b).max(),
                                           err_msg='%s k=%d %s' % (name, k, term))
    print('%-8s d=%d rows=%5d derivatives=%s  worst value error %.1e' % (name, d, n_rows, with_derivatives, worst))


def check_truncation(engine):
    """The truncation a plan prescribes (csrc/plan.cu: lag windows of 9 standard deviations / 42
    decay lengths, limits at infinity beyond 10 standard deviations / 60 decay lengths of T_f)
    does not change the packed sums: rows evaluated in full against rows evaluated the way the
    kernels traverse them, both assembled by the library, for parameters drawn over a wide range."""
    from cases import square
    rng = np.random.default_rng(23)
    T_f = 400.
    worst = 0.
    families = {'exp': lambda: [{'type': 'exp'}], 'gauss': lambda: [{'type': 'gauss'}],
                'mix': lambda: [{'type': 'exp'}, {'type': 'gauss'}]}
    for trial in range(24):
        kind = ('exp', 'gauss', 'mix')[trial % 3]
        d = 1 + trial % 2
        model = square(d, families[kind])
        times = [np.sort(rng.uniform(0, T_f, size=int(rng.integers(120, 220)))) for _ in range(d)]
        cell = {'exp': lambda: [rng.uniform(.05, .3), rng.uniform(.15, 4.)],
                'gauss': lambda: [rng.uniform(.05, .3), rng.uniform(.1, 2.5), rng.uniform(0., 6.)],
                'mix': lambda: [rng.uniform(.05, .2), rng.uniform(.3, 3.), rng.uniform(.05, .2), rng.uniform(.2, 1.5),
                                rng.uniform(.2, 4.)]}[kind]
        xs = [np.array([rng.uniform(.3, .9)] + [v for _ in range(d) for v in cell()]) for _ in range(d)]
        engine.set_events(times, T_f)
        engine.set_model(model)
        scratch = np.zeros(engine.exact_raw_len_batch(d))
        engine.exact_raw_device_batch(list(range(d)), xs, scratch.ctypes.data, True, True)
        meta, par = engine.plan_rows()
        full, cut = np.zeros((len(meta), N_COMP)), np.zeros((len(meta), N_COMP))
        for r in range(len(meta)):
            with np.errstate(over='ignore', invalid='ignore'):
                full[r] = row_components(meta[r], par[r], times, T_f, True)
                cut[r] = row_components(meta[r], par[r], times, T_f, True, truncated=True)
        if not np.all(np.isfinite(full)):
            continue   # exp x gauss beyond the reference's own overflow (inf x 0): not a truncation matter
        ks = list(range(d))
        a = engine.exact_assemble_batch(ks, xs, full.ravel(), True, True)
        b = engine.exact_assemble_batch(ks, xs, cut.ravel(), True, True)
        for k in ks:
            # per packed entry (term value or derivative): relative to itself, with a floor relative to
            # the largest entry of the sub-problem - a term that lives entirely in a far Gaussian tail
            # (its basis pair peaks at a negative lag) is cut relative to the peak, not to itself
            floor = 1e-16 * np.abs(a[k]).max()
            err = np.abs(b[k] - a[k]) / np.maximum(np.abs(a[k]), 1e-300)
            bad = (err > 1e-12) & (np.abs(b[k] - a[k]) > floor)
            assert not bad.any(), (kind, d, trial, k, np.flatnonzero(bad), a[k][bad], b[k][bad])
            worst = max(worst, float(err[np.abs(a[k]) > 1e-6 * np.abs(a[k]).max()].max()))
    print('truncation: 24 parameter draws, worst relative change of a packed sum %.1e' % worst)


def check_shards(name, engine):
    """Time-block shards (aslsd_b200/distributed.py): the plans of the blocks restrict every row to
    the block's targets of the row's own target type (the type of the sub-problem for phi rows);
    the packed sums of the blocks add up to those of the whole path."""
    from aslsd_b200.distributed import time_block_ranges
    model, xs = (CASES[name][0], case_x(name)) if name in CASES else larger_model(name)
    d = model['d']
    rng = np.random.default_rng(31)
    T_f = 30.
    times = [np.sort(rng.uniform(0, T_f, size=45 + 9 * i)) for i in range(d)]
    engine.set_events(times, T_f)
    engine.set_model(model)
    ks = list(range(d))
    scratch = np.zeros(engine.exact_raw_len_batch(d))

    def packed(ranges):
        engine.exact_raw_device_batch(ks, xs, scratch.ctypes.data, True, True, ranges=ranges)
        meta, par = engine.plan_rows()
        raw = np.array([row_components(meta[r], par[r], times, T_f, True) for r in range(len(meta))])
        return engine.exact_assemble_batch(ks, xs, raw.ravel(), True, True)
    whole = packed(None)
    blocks = time_block_ranges(times, 3)
    assert all(sum(b[i][1] - b[i][0] for b in blocks) == len(times[i]) for i in range(d))
    acc = [np.zeros_like(w) for w in whole]
    for block in blocks:
        for a, part in zip(acc, packed(block)):
            a += part
    for k in ks:
        np.testing.assert_allclose(acc[k], whole[k], rtol=1e-12, atol=1e-13 * np.abs(whole[k]).max())
    print('%-8s d=%d: three time blocks add up to the whole path' % (name, d))


class InjectingEngine:
    """The real engine (on the do-nothing runtime) whose weighted exact sums get their raw rows from
    numpy: the fused entry point is called once to plan, the rows of the plan are evaluated here
    with the weights that were set, and the call is repeated with the read-back of the raw array
    substituted (aslsd_stub_next_d2h) - planning, role / mask handling and assembly are the
    library's, as on the device."""

    def __init__(self, stub_runtime):
        import ctypes
        self._real = aslsd_b200.Engine(0)
        self._stub = ctypes.CDLL(stub_runtime)
        self._stub.aslsd_stub_next_d2h.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        self._stub.aslsd_stub_next_d2h.restype = None
        self._weights = []
        self._times = None

    def __getattr__(self, name):
        return getattr(self._real, name)

    def bind(self, process_path, model):
        self._times = [np.asarray(t, dtype=float) for t in process_path.list_times]
        self._T_f = float(process_path.T_f)
        return self._real.bind(process_path, model)

    def set_weights(self, vectors):
        self._weights = [(int(i), np.asarray(v, dtype=float)) for i, v in vectors]
        return self._real.set_weights(vectors)

    def set_horizon(self, T_f):
        self._T_f = float(T_f)
        return self._real.set_horizon(T_f)

    def exact_sums_weighted(self, k, x_k, roles, term_mask=None, compute_f_sum=True, compute_diff_f_sum=True,
                            ranges=None):
        args = (k, x_k, roles, term_mask, compute_f_sum, compute_diff_f_sum, ranges)
        self._real.exact_sums_weighted(*args)            # plans; the kernels are no-ops
        meta, par = self._real.plan_rows()
        raw = np.zeros((max(len(meta), 1), N_COMP))
        for r in range(len(meta)):
            raw[r] = row_components(meta[r], par[r], self._times, self._T_f, True, weights=self._weights)
        self._stub.aslsd_stub_next_d2h(raw.ctypes.data, raw.nbytes)
        return self._real.exact_sums_weighted(*args)     # same plan again, read-back substituted


def check_marked(stub_runtime):
    """Marked models end to end on the CPU: the B200 estimator (aslsd_b200/mtlh.py) over the library's
    planning and assembly against the golden vectors of the reference's MTLHExactEstim, one to three
    event types."""
    import test_gpu_mtlh as TM
    from cases import MTLH_CASES
    for name in sorted(MTLH_CASES):
        TM.test_mtlh_exact_matches_reference_golden(InjectingEngine(stub_runtime), name)
        print('%-16s marked model: estimator state equals the golden vectors of the reference' % name)


def main():
    engine = aslsd_b200.Engine(0)
    check_marked(os.path.join(os.path.dirname(sys.argv[1]), 'libcudart_stub.so'))
    for name in ('exp2d', 'sbf3d2', 'mix3d', 'exp5d'):
        check_shards(name, engine)
    check_truncation(engine)
    for name in ('exp1d', 'exp2d', 'exp1d2b', 'exp3d', 'exp5d'):
        check(name, engine, with_derivatives=True)
    # sums of Gaussians with (beta, delta) fixed - the model family of configuration C5: every free
    # parameter is an omega
    for name in ('sbf1d3', 'sbf2d3', 'sbf1d10', 'sbf3d2', 'sbf5d2', 'C5'):
        check(name, engine, with_derivatives=True)
    # free Gaussians and mixed fixed patterns: the beta / delta derivative components too
    for name in ('gauss1d', 'gauss2d', 'gauss1d2b', 'gauss3d'):
        check(name, engine, with_derivatives=True)
    # an exponential and a Gaussian basis in one kernel: the mixed pairs
    for name in ('mix1d', 'mix2d', 'mix3d', 'mix4d'):
        check(name, engine, with_derivatives=True)
    print('plan and assembly ok')


if __name__ == '__main__':
    main()
