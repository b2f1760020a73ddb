"""Drop-in `network` package (seam #1): put `deeppointmap_b200/dropin` in front of the reference
tree on sys.path and `from network.encoder.encoder import Encoder` / `from network.decoder.decoder
import Decoder` (pipeline/infer.py:31-32) resolve to the B200 modules; every other `network.*`
submodule (loss, pointnext, ...) still resolves to the reference's own files."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
