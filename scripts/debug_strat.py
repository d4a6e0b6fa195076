import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import numpy as np
import aslsd_b200
from cases import CASES
from helpers import load_golden
from oracle_engine import OracleEngine

name = sys.argv[1] if len(sys.argv) > 1 else 'strat_exp2d'
g = load_golden(name)
case = str(g['case']); model = CASES[case][0]; d = int(g['d'])
times = [np.asarray(g['times/%d' % i]) for i in range(d)]; T_f = float(g['T_f'])
kw = {str(k): int(v) for k, v in zip(g['kw_names'], g['kw_values'])}
eng = aslsd_b200.Engine(0)
mirror = aslsd_b200.MHP(model)
path = aslsd_b200.ProcessPath(times, T_f, engine=eng)
ests = []
for E in (eng, OracleEngine()):
    rng = np.random.default_rng(int(g['seed']))
    est = aslsd_b200.MHPStratEstimB200(is_log_lse=True, engine=E)
    est.initialize(0, 5, mirror, path, **kw)
    orig = E.sampled_sums
    rec = {}
    def wrap(*a, _o=orig, _r=rec, **k):
        out = _o(*a, **k); _r['args'] = (a, k); _r['out'] = out; return out
    E.sampled_sums = wrap
    est.lse_k_grad_estimate(g['k0/x'][0], rng=rng, grad_alloc=bool(g['grad_alloc']))
    ests.append((est, rec))
(a, ra), (b, rb) = ests
for nm in sorted(a.stratif):
    sa, sb = a.stratif[nm], b.stratif[nm]
    if sa.n_adaptive_strata:
        err = np.max(np.abs(sa.adaptive_rel_alloc - sb.adaptive_rel_alloc) / sb.adaptive_rel_alloc)
        print(nm, 'alloc rel err %.3e' % err)
sums_a, off_a, mean_a, std_a, rel_a = ra['out']
sums_b, off_b, mean_b, std_b, rel_b = rb['out']
bad = np.flatnonzero(np.abs(std_a - std_b) > 1e-9 * np.abs(std_b))
print('strata with std mismatch', bad, 'of', len(std_a))
for s in bad[:12]:
    print(s, 'std gpu %.17g oracle %.17g mean gpu %.17g oracle %.17g' % (std_a[s], std_b[s], mean_a[s], mean_b[s]))
