import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    import ref_shim
    ref_shim.make_importable()   # before aslsd_b200 is imported (see ref_shim)
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs the read-only reference tree '
                            '(/root/reference); skipped where it is absent')


def pytest_collection_modifyitems(config, items):
    import ref_shim
    if not ref_shim.reference_available():
        skip = pytest.mark.skip(reason='reference tree not present')
        for item in items:
            if 'reference' in item.keywords:
                item.add_marker(skip)
