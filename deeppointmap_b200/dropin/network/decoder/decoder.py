"""network.decoder.decoder.Decoder -> the B200 decoder (same ctor / methods / state_dict)."""
from deeppointmap_b200.decoder import Decoder  # noqa: F401
