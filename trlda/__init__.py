"""Drop-in alias: `import trlda` resolves to the B200-native implementation in trlda_b200 (same names as the
reference package, code/trlda/python/__init__.py)."""
from trlda_b200 import __version__, __license__, __docformat__, seed
from trlda_b200 import models, utils

__author__ = 'trlda_b200 (API of Lucas Theis <lucas@theis.io>)'
