"""cProfile of ASLSD fit epochs (host-side cost of the stratified estimator), GPU box."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import aslsd_b200
from aslsd_b200 import simulate
from aslsd_epoch import CONFIGS

name = sys.argv[1] if len(sys.argv) > 1 else 'C2'
model, mu, kp, T_f = CONFIGS[name]
d = model['d']
engine = aslsd_b200.Engine(0)
times = simulate.simulate_mhp(model, mu, kp, T_f, seed=1234)
path = aslsd_b200.ProcessPath(times, T_f, engine=engine)
mirror = aslsd_b200.MHP(model)
x0 = [np.concatenate([[mu[k]]] + [np.asarray(kp[k][i], float) for i in range(d)]) for k in range(d)]
mirror.fit(path, [x.copy() for x in x0], n_iter=3, engine=engine, seed=1)
pr = cProfile.Profile()
pr.enable()
mirror.fit(path, [x.copy() for x in x0], n_iter=40, engine=engine, seed=1)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('cumulative').print_stats(28)
