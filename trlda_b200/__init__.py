"""
trlda_b200 — B200-native (sm_100a CUDA) implementation of trlda's variational E-step and trust-region M-step,
behind the reference's Python API:

    from trlda_b200.models import OnlineLDA, BatchLDA, CumulativeLDA      # or: from trlda.models import ...
    from trlda_b200.utils import load_documents

Layers (reference counterparts in parentheses, paths under /root/reference/code/trlda):
    trlda_b200.models / trlda_b200.utils    Python package glue (python/__init__.py, python/models, python/utils)
    trlda_b200._trlda                       CPython extension, csrc/pymodule.cpp (python/src/*.cpp)
    libtrlda_b200.so                        C ABI (include/trlda_b200.h) + CUDA kernels (src/*.cpp, the hot path)
    trlda_b200.capi                         ctypes view of the same C ABI, used by the parity tests and bench.py

There is no CPU fallback: importing the models raises if the native pieces have not been built
(`python -m trlda_b200.build`), and constructing a model raises without a B200.
"""
__version__ = '0.1.0'
__license__ = 'MIT License <http://www.opensource.org/licenses/mit-license.php>'
__docformat__ = 'epytext'


def seed(value):
	"""Seeds the generators behind the random initial gamma / lambda (reference: trlda.seed, module.cpp:332-342)."""
	from . import _trlda
	_trlda.seed(int(value))
