"""trlda.models of the reference (code/trlda/python/models/__init__.py:1-5), served by the CUDA-backed extension."""
try:
	from .._trlda import Distribution, LDA, OnlineLDA, BatchLDA, CumulativeLDA
except ImportError as error:   # pragma: no cover - build problem, never a silent fallback
	raise ImportError(
		'trlda_b200: the native extension is not built (%s); run `python -m trlda_b200.build` — there is no '
		'CPU fallback' % error)

__all__ = ['Distribution', 'LDA', 'OnlineLDA', 'BatchLDA', 'CumulativeLDA']
