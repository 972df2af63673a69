"""Model registry: the class names the reference resolves from the YAML model key
(elliot/run.py:75, elliot/recommender/__init__.py)."""
from .bprmf import BPRMF, MFModel  # noqa: F401
from .bprmf_batch import BPRMF_batch, BPRMFBatchModel  # noqa: F401
from .multi_vae import MultiVAE, VariationalAutoEncoder  # noqa: F401
from .neumf import NeuMF, NeuralMatrixFactorizationModel  # noqa: F401
from .mf2020 import MF2020, MF2020Model  # noqa: F401
from .multi_dae import MultiDAE, DenoisingAutoEncoder  # noqa: F401
from .gmf import GMF, GeneralizedMatrixFactorizationModel  # noqa: F401
