"""`model_seq` as the reference drivers import it (train_sr.py:18 `from model_seq import *`).

Put this directory first on sys.path (PYTHONPATH=/path/to/repo/amid_b200/dropin:/path/to/repo)
and train_sr.py / train_sr_dr.py construct the B200-native SASRec unchanged.
"""
from amid_b200.model_seq import *  # noqa: F401,F403
from amid_b200.model_seq import SASRec  # noqa: F401
