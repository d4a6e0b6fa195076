"""List operations, statistics and serialization of the ProcessPath / Episodes mirrors
(aslsd_b200/events.py) against the live reference (stats/events/process_path.py:110-336,
episodes.py).  CPU: the index maps come from the oracle test double; the device builds the same
maps in tests/test_gpu_index_maps.py."""
import numpy as np
import pytest

import ref_shim
from oracle_engine import OracleEngine

pytestmark = pytest.mark.reference


@pytest.fixture(scope='module')
def ref():
    return ref_shim.import_reference()


def _paths(with_marks=False):
    import aslsd_b200
    from aslsd.stats.events.process_path import ProcessPath as RefPath
    rng = np.random.default_rng(8)
    T_f = 40.
    times = [np.sort(rng.uniform(0, T_f, size=n)) for n in (55, 35, 48)]
    marks = [rng.uniform(0, 1, size=(len(t), 2)) for t in times] if with_marks else None
    return (aslsd_b200.ProcessPath(times, T_f, list_marks=marks, engine=OracleEngine()),
            RefPath(times, T_f, list_marks=marks), times, T_f)


@pytest.mark.parametrize('T_trunc', [5.1, 13.3, 39.9])
@pytest.mark.parametrize('with_marks', [False, True])
def test_truncate(ref, T_trunc, with_marks):
    new, old, times, T_f = _paths(with_marks)
    a, b = old.truncate(T_trunc), new.truncate(T_trunc)
    assert b.engine is not new.engine and isinstance(b.engine, OracleEngine)
    assert b.T_f == a.T_f == T_trunc and b.d == a.d
    np.testing.assert_array_equal(b.n_events, a.n_events)
    np.testing.assert_array_equal(b.eta, a.eta)
    for i in range(3):
        np.testing.assert_array_equal(b.list_times[i], a.list_times[i])
        np.testing.assert_array_equal(b.list_times2end[i], a.list_times2end[i])
        if with_marks:
            np.testing.assert_array_equal(b.list_marks[i], a.list_marks[i])
        for j in range(3):
            # the reference's sliced maps: kappa without its terminal entry, varpi possibly with
            # entries past the last lag of the shorter path; equal wherever both are defined
            n = min(len(a.kappa[j][i]), len(b.kappa[j][i]))
            assert n == a.n_events[i]
            np.testing.assert_array_equal(b.kappa[j][i][:n], a.kappa[j][i][:n])
            n_lags = len(b.varpi[i][j])
            m = min(n_lags, len(a.varpi[i][j]))
            np.testing.assert_array_equal(np.asarray(b.varpi[i][j])[:m], np.asarray(a.varpi[i][j])[:m])
            np.testing.assert_array_equal(np.asarray(b.lag_sizes[i][j])[:m],
                                          np.asarray(a.lag_sizes[i][j], dtype=np.int64)[:m])


def test_truncate_beyond_the_end_and_clone(ref):
    new, old, times, T_f = _paths()
    c = new.truncate(T_f + 1.)          # upstream: clone() -> TypeError (unknown keyword d, :122)
    with pytest.raises(TypeError):
        old.truncate(T_f + 1.)
    assert c is not new and c.T_f == T_f and c.kappa is new.kappa and c.engine is new.engine
    for i in range(3):
        np.testing.assert_array_equal(c.list_times[i], new.list_times[i])


def test_censor(ref):
    new, old, times, T_f = _paths(with_marks=True)
    drop = [np.array([0, 5, 6]), None, np.array([], dtype=int)]
    with pytest.raises(TypeError):      # upstream: same unknown keyword (:196)
        old.censor(drop)
    c = new.censor(drop)
    assert c.book_keeping is False and c.kappa is None
    np.testing.assert_array_equal(c.list_times[0], np.delete(times[0], drop[0]))
    np.testing.assert_array_equal(c.list_marks[0], np.delete(new.list_marks[0], drop[0], axis=0))
    np.testing.assert_array_equal(c.list_times[1], times[1])
    np.testing.assert_array_equal(c.n_events, [52, 35, 48])


def test_small_helpers(ref):
    new, old, times, T_f = _paths(with_marks=True)
    for a, b in zip(old.get_inter_arrival_times(), new.get_inter_arrival_times()):
        np.testing.assert_array_equal(b, a)
    grid = np.linspace(0., T_f, 400)
    for a, b in zip(old.get_counting_process(), new.get_counting_process()):
        np.testing.assert_array_equal(b(grid), a(grid))
    t_ref = np.array([0., 3.3, 17.1, 30.])
    for a, b in zip(old.get_next_event(t_ref), new.get_next_event(t_ref)):
        assert (b.time, b.dim) == (a.time, a.dim)
        np.testing.assert_array_equal(b.mark, a.mark)
    one = new.get_next_event(3.3)
    assert one.time == old.get_next_event(3.3).time
    np.testing.assert_array_equal(new.get_sample_grid(0.7), old.get_sample_grid(0.7))
    np.testing.assert_array_equal(new.get_sample_hist(0.7), old.get_sample_hist(0.7))
    for a, b in zip(old.get_empirical_intensity(0.7), new.get_empirical_intensity(0.7)):
        np.testing.assert_array_equal(b, a)
    np.testing.assert_allclose(new.get_empirical_covariance(0.5, max_tau=4.),
                               old.get_empirical_covariance(0.5, max_tau=4.), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(new.get_empirical_covariance(2.), old.get_empirical_covariance(2.),
                               rtol=1e-12, atol=1e-14)
    bounds = [np.linspace(0, T_f, 9)] * 3
    rate = new.get_events_rate(bounds)          # upstream cannot run (:246 subtracts lists, :247 reads self.times)
    with pytest.raises((TypeError, AttributeError)):
        old.get_events_rate(bounds)
    for i in range(3):
        np.testing.assert_allclose(rate[i] * 5., np.histogram(times[i], bounds[i])[0])


def test_save_load_both_ways(ref, tmp_path):
    """One pickled dict with the reference's keys: either implementation reads the other's file."""
    import aslsd_b200
    from aslsd.stats.events.process_path import ProcessPath as RefPath
    new, old, times, T_f = _paths(with_marks=True)
    new.save(str(tmp_path / 'new'))
    old.save(str(tmp_path / 'old'))
    for source, cls in (('new', RefPath), ('old', aslsd_b200.ProcessPath), ('new', aslsd_b200.ProcessPath)):
        back = cls() if cls is RefPath else cls(book_keeping=False)
        back.load(str(tmp_path / source))
        assert back.T_f == T_f and back.d == 3 and back.book_keeping
        for i in range(3):
            np.testing.assert_array_equal(back.list_times[i], times[i])
            np.testing.assert_array_equal(back.list_marks[i], new.list_marks[i])
            for j in range(3):
                np.testing.assert_array_equal(back.kappa[j][i], old.kappa[j][i])
                np.testing.assert_array_equal(back.varpi[i][j], old.varpi[i][j])
                np.testing.assert_array_equal(np.asarray(back.lag_sizes[i][j], dtype=np.int64),
                                              np.asarray(old.lag_sizes[i][j], dtype=np.int64))


def test_episodes(ref, tmp_path):
    import aslsd_b200
    from aslsd.stats.events.episodes import Episodes as RefEpisodes
    rng = np.random.default_rng(2)
    paths = [aslsd_b200.ProcessPath([np.sort(rng.uniform(0, 20., size=n)) for n in (25, 18)], 20.,
                                    engine=OracleEngine()) for _ in range(3)]
    ep = aslsd_b200.Episodes(paths, T_f=20.)
    cut = ep.truncate(11.)              # upstream: AttributeError ('list' has no truncate, :93)
    assert cut.n_episodes == 3 and cut.T_f == 11.
    for q in range(3):
        for i in range(2):
            t = paths[q].list_times[i]
            np.testing.assert_array_equal(cut.list_paths[q].list_times[i], t[t < 11.])
    np.testing.assert_array_equal(cut.total_n_events,
                                  [sum(int((p.list_times[i] < 11.).sum()) for p in paths) for i in range(2)])
    bounds = [np.linspace(0, 20., 5)] * 2
    rate = ep.get_events_rate(bounds)
    np.testing.assert_allclose(rate[0] * 5., sum(np.histogram(p.list_times[0], bounds[0])[0] for p in paths))
    ep.save(str(tmp_path / 'ep'))
    back = RefEpisodes()
    back.load(str(tmp_path / 'ep'))     # the reference reads the mirror's files
    mine = aslsd_b200.Episodes()
    mine.load(str(tmp_path / 'ep'))
    for loaded in (back, mine):
        assert loaded.n_episodes == 3 and loaded.d == 2
        np.testing.assert_array_equal(loaded.list_n_events, ep.list_n_events)
        for q in range(3):
            for i in range(2):
                np.testing.assert_array_equal(loaded.list_paths[q].list_times[i], paths[q].list_times[i])
                np.testing.assert_array_equal(loaded.list_paths[q].kappa[1 - i][i], paths[q].kappa[1 - i][i])
