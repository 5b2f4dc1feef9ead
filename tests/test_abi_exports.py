"""The C-ABI library loads on a CPU-only box and exports every symbol include/colddiff.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, 'include', 'colddiff.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\bint\s+(cd_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from cold_diffusion_models_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.cd_version() == 1


def test_error_text_roundtrip_without_gpu():
    from cold_diffusion_models_b200 import _lib
    d = _lib.ConvDesc()
    d.nsrc = 3                                   # invalid -> rejected before any CUDA call
    rc = _lib.lib.cd_conv_fwd(ctypes.byref(d), 0, None)
    assert rc != 0
    buf = ctypes.create_string_buffer(256)
    assert _lib.lib.cd_last_error(buf, 256) > 0 and b'nsrc' in buf.value


def test_struct_layout_matches_header():
    """ctypes mirrors of CdConvSrc / CdConvDesc have the C sizes (guards against silent ABI drift)."""
    import subprocess, tempfile, sys
    from cold_diffusion_models_b200 import _lib
    prog = '#include "colddiff.h"\n#include <stdio.h>\nint main(){printf("%zu %zu\\n", sizeof(CdConvSrc), sizeof(CdConvDesc));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, 'a.c')
        open(c, 'w').write(prog)
        exe = os.path.join(td, 'a.out')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), c, '-o', exe])
        a, b = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(_lib.ConvSrc) == a and ctypes.sizeof(_lib.ConvDesc) == b
