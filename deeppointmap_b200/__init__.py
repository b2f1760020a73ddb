"""deeppointmap_b200 -- B200-native (sm_100a) implementation of DeepPointMap's per-frame hot
path: the point-cloud encoder and the registration decoder, behind the reference's own
module / op APIs.  See DESIGN.md and include/dpm_b200.h."""
from .encoder import Encoder  # noqa: F401
from .decoder import Decoder  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["Encoder", "Decoder", "ops"]
