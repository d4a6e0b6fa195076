"""Which device context a path gets (aslsd_b200.events.engine_for): the default one while it is
free, one context per further live path (fit_episodes keeps all episodes resident).  CPU test:
the engine class is replaced by a stand-in, no device is touched."""
import types

import numpy as np

from aslsd_b200 import events


class FakeEngine:
    made = 0

    def __init__(self, device=0):
        FakeEngine.made += 1
        self.device = device
        self._events_key = None

    def hold(self, path):
        self._events_key = events._path_key(path)


def _path(n):
    return types.SimpleNamespace(list_times=[np.arange(float(n))], T_f=float(n))


def test_paths_get_their_own_context(monkeypatch):
    FakeEngine.made = 0
    monkeypatch.setattr(events, 'Engine', FakeEngine)
    monkeypatch.setattr(events, '_default_engine', None)
    a, b, c = _path(5), _path(6), _path(7)
    e_a = events.engine_for(a)
    assert e_a is events.default_engine() and FakeEngine.made == 1
    e_a.hold(a)
    assert events.engine_for(a) is e_a                    # same path: same context
    e_b = events.engine_for(b)
    assert e_b is not e_a and b.engine is e_b             # second live path: its own, kept on the path
    assert events.engine_for(b) is e_b
    e_b.hold(b)
    e_c = events.engine_for(c)
    assert e_c not in (e_a, e_b) and FakeEngine.made == 3
    # an explicit engine on the path always wins
    own = FakeEngine()
    d = _path(8)
    d.engine = own
    assert events.engine_for(d) is own
    # the default context is taken again once set_events has been called on it for other data
    e_a._events_key = None
    assert events.engine_for(_path(9)) is e_a


def test_episodes_container_statistics():
    paths = [types.SimpleNamespace(d=2, n_events=np.array([10, 20]), eta=np.array([1., 2.])),
             types.SimpleNamespace(d=2, n_events=np.array([30, 40]), eta=np.array([3., 4.]))]
    ep = events.Episodes(paths, T_f=10.)
    assert ep.n_episodes == 2 and ep.d == 2
    np.testing.assert_array_equal(ep.total_n_events, [40, 60])
    np.testing.assert_allclose(ep.average_eta, [2., 3.])
    empty = events.Episodes()
    assert empty.n_episodes == 0 and empty.d == 0
