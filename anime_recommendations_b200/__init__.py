"""anime_recommendations_b200 -- B200 (sm_100a) hot path of Dyrutter/anime_recommendations.

Host side in Python (the reference's language); all arithmetic in libanimerec.so through the C-ABI of
include/animerec.h.  See DESIGN.md and INTEGRATION.md.
"""
from ._capi import AnimerecError, lib  # noqa: F401
from .callbacks import EarlyStopping, LearningRateScheduler, ModelCheckpoint  # noqa: F401
from .model import EmbeddingDotModel, History, load_model  # noqa: F401
from .schedule import lrfn  # noqa: F401
from . import similarity  # noqa: F401

__all__ = ["EmbeddingDotModel", "History", "load_model", "lrfn", "similarity", "LearningRateScheduler",
           "ModelCheckpoint", "EarlyStopping", "AnimerecError", "lib"]
