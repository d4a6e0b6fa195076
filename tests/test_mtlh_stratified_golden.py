"""MTLHStratEstimB200 (host logic: sampling order, weight roles, product rule over the impact
parameters, allocations) on the oracle-backed engine double, against traces of the reference's
MTLHStratEstim.  The same replay runs on the CUDA engine in tests/test_gpu_mtlh.py."""
import pytest

from cases import MTLH_STRAT
from mtlh_replay import replay
from oracle_engine import OracleEngine


@pytest.mark.parametrize('name', sorted(MTLH_STRAT))
def test_replay_reference_mtlh_trace(name):
    replay(name, OracleEngine())
