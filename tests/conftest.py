import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)


def pytest_configure(config):
	config.addinivalue_line('markers', 'gpu: needs a B200 (run on the GPU box with `-m gpu`)')


@pytest.fixture(scope='session')
def oracle_built():
	"""the plain-C oracle (and, where /root/reference exists, the compiled reference) built by oracle/Makefile"""
	from oracle import pyoracle
	if not pyoracle.have_port() or (os.path.isdir('/root/reference') and not pyoracle.have_ref()):
		pyoracle.build()
	return pyoracle
