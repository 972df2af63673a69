"""Older name of the plugin entry module; the reference-facing entry point is elliot_b200/external/__init__.py (it must
be a package `__init__.py` for the reference's logger lookup to work, see there).  Kept so that stand-alone configs
that point `external_models_path` at this file keep working with elliot_b200.run."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from elliot_b200.recommender import *  # noqa: F401,F403,E402
from elliot_b200.recommender import BPRMF, BPRMF_batch, MF2020, MultiVAE, NeuMF  # noqa: F401,E402
