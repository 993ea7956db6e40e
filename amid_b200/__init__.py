"""amid_b200: B200-native (sm_100a) engine for AMID's SASRec cross-domain recommender hot path."""
from ._abi import AmidError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
