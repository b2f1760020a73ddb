"""network.encoder.encoder.Encoder -> the B200 encoder (same ctor / forward / state_dict)."""
from deeppointmap_b200.encoder import Encoder  # noqa: F401
