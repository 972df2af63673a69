"""Pick the base classes the plugin models derive from.

Inside a real Elliot install (tensorflow + hyperopt present) the reference's own
BaseRecommenderModel / RecMixin / init_charger are used, so `external.<Model>` classes loaded
through `external_models_path` (elliot/run.py:67-73) are genuine subclasses of the host
framework's ABC.  Otherwise (this container, the GPU box) the stand-alone mirror is used."""
try:  # pragma: no cover - exercised only inside a full reference install
    from elliot.recommender.base_recommender_model import BaseRecommenderModel, init_charger
    from elliot.recommender.recommender_utils_mixin import RecMixin
    HOST = "elliot"
except Exception:
    from .base_recommender_model import BaseRecommenderModel, init_charger
    from .recommender_utils_mixin import RecMixin
    HOST = "standalone"
