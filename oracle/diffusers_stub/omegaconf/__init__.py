"""TEST INFRASTRUCTURE ONLY: the one symbol of `omegaconf` the reference's utils/p2p_utils/ptp_utils.py touches
(an isinstance check against omegaconf.dictconfig.DictConfig, :119); the real package is not installed offline."""
from . import dictconfig  # noqa: F401
