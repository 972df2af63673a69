"""Plugin entry module for the reference's `external_models_path` mechanism
(elliot/run.py:67-73, docs/source/guide/new_alg.rst): point the YAML at this file and name the
models `external.BPRMF`, ...  In a real Elliot install the classes derive from Elliot's own
BaseRecommenderModel (see recommender/_bases.py)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from elliot_b200.recommender import *  # noqa: F401,F403,E402
from elliot_b200.recommender import BPRMF, BPRMF_batch, MF2020, MultiVAE, NeuMF  # noqa: F401,E402
