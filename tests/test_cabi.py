"""The C-ABI library builds, loads and exports every symbol include/dpm_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "dpm_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dpm_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    from deeppointmap_b200 import build
    so = build.build()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    syms = _header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_table_matches_header():
    from deeppointmap_b200 import _C
    assert sorted(_C.EXPORTS) == _header_symbols()
    lib = _C.lib()
    assert lib.dpm_version() >= 100
    assert lib.dpm_last_error() is not None


def test_struct_layout_matches_header():
    """ctypes mirrors of dpm_encoder_desc / dpm_decoder_desc: sizes as the C compiler lays them out."""
    import subprocess
    import tempfile
    from deeppointmap_b200 import _C
    src = '#include <stdio.h>\n#include "dpm_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(dpm_encoder_desc), sizeof(dpm_decoder_desc));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        a, b = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert ctypes.sizeof(_C.EncoderDesc) == a
    assert ctypes.sizeof(_C.DecoderDesc) == b


def test_sizing_entry_points_run_without_gpu(cfg):
    from deeppointmap_b200 import Encoder, Decoder, _C
    e, d = Encoder(cfg), Decoder(cfg)
    lib = _C.lib()
    assert lib.dpm_encoder_num_weights(ctypes.byref(e._desc)) == 110
    assert lib.dpm_decoder_num_weights(ctypes.byref(d._desc)) == 83
    assert lib.dpm_encoder_out_points(ctypes.byref(e._desc)) == 256
    nb = lib.dpm_encoder_workspace_bytes(ctypes.byref(e._desc), 1, 65536)
    assert 8 << 20 < nb < 256 << 20
    assert lib.dpm_registration_workspace_bytes(ctypes.byref(d._desc), 1, 256, 256, 128) > 1 << 20
    assert lib.dpm_fps_workspace_bytes(2, 1000, 3, 10) > 0
