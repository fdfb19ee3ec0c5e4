"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol that ``include/pmwd_b200.h`` declares; bad arguments are rejected with an error code and
message; nothing under ``pmwd_b200/`` imports the oracle.  (No compute calls: no GPU here.)"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from pmwd_b200.build import build_lib
    build_lib()
    from pmwd_b200 import _lib
    return _lib.lib()


def _declared():
    text = open(os.path.join(ROOT, 'include', 'pmwd_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(pmwd_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/pmwd_b200.h but not exported'
    from pmwd_b200 import _lib
    assert set(_lib.EXPORTS) == set(names)


def test_abi_version_and_errors(lib):
    from pmwd_b200 import _lib
    assert lib.pmwd_abi_version() == 2
    # bad arguments -> negative status + message, no crash, no GPU needed
    rc = lib.pmwd_laplace(None, 5, _lib.shape_arr((4, 4, 4)), 1.0, None, None)
    assert rc == -1 and 'rank' in _lib.last_error()
    d = _lib.CicDesc()
    d.dim = 7
    rc = lib.pmwd_scatter(None, ctypes.byref(d), ctypes.c_void_p(8), ctypes.c_void_p(8), None, 1.0,
                          ctypes.c_void_p(8), 0, None, 0)
    assert rc == -1 and 'dim' in _lib.last_error()
    rc = lib.pmwd_kick_drift(None, -1, None, ctypes.c_void_p(16), None, 0.0, 0.0, 1, 1)
    assert rc == -1
    # power-spectrum entries: null buffers, too many bin edges
    shp = _lib.shape_arr((8, 8, 8))
    rc = lib.pmwd_powspec_bin(None, shp, None, None, 0, 0.0, ctypes.c_void_p(8), 4, 0, ctypes.c_void_p(8))
    assert rc == -1 and 'null' in _lib.last_error()
    rc = lib.pmwd_powspec_bin(None, shp, ctypes.c_void_p(8), None, 0, 0.0, ctypes.c_void_p(8), 100000, 0,
                              ctypes.c_void_p(8))
    assert rc == -1 and 'nedges' in _lib.last_error()
    rc = lib.pmwd_powspec_weight(None, shp, ctypes.c_void_p(8), 0, 0.0, ctypes.c_void_p(8), 4, 0, None,
                                 ctypes.c_void_p(8))
    assert rc == -1
    # fused x-pass: unsupported length, bad slab
    bad = _lib.shape_arr((96, 8, 8))
    arr = (ctypes.c_void_p * 3)(8, 8, 8)
    rc = lib.pmwd_xpass_force(None, bad, 0, 8, 1.0, 1.0, ctypes.c_void_p(8), arr)
    assert rc == -1 and 'nx' in _lib.last_error()
    ok = _lib.shape_arr((256, 8, 8))
    rc = lib.pmwd_xpass_force(None, ok, 4, 8, 1.0, 1.0, ctypes.c_void_p(8), arr)
    assert rc == -1 and 'slab' in _lib.last_error()
    assert lib.pmwd_xpass_supported(1024) == 1 and lib.pmwd_xpass_supported(96) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'pmwd_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert 'oracle/' not in src, f


def test_sass_has_vector_red():
    """The scatter kernel's float32 reductions are the native RED.E.ADD.F32 (+ .F32x2 pairs)."""
    import subprocess
    from pmwd_b200.build import LIB
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    assert 'REDG.E.ADD.F32x2' in out and 'REDG.E.ADD.F32.' in out


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/pmwd_b200.h must compile as C99 (-pedantic) and a plain
    C program must link against the library and call it (no compute: ABI version and error string)."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    src = tmp_path / 'use_abi.c'
    src.write_text('#include <string.h>\n#include "pmwd_b200.h"\n'
                   'int main(void) {\n'
                   '  char msg[64];\n'
                   '  if (pmwd_abi_version() != 2) return 1;\n'
                   '  /* a bad call must fail with a negative status and leave a message */\n'
                   '  if (pmwd_laplace(0, 3, 0, 1.0, 0, 0) >= 0) return 2;\n'
                   '  pmwd_last_error(msg, sizeof msg);\n'
                   '  return strlen(msg) > 0 ? 0 : 3;\n'
                   '}\n')
    exe = tmp_path / 'use_abi'
    libdir = os.path.join(ROOT, 'pmwd_b200')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I' + os.path.join(ROOT, 'include'),
                    str(src), '-L' + libdir, '-lpmwd_b200', '-Wl,-rpath,' + libdir, '-o', str(exe)], check=True)
    r = subprocess.run([str(exe)])
    assert r.returncode == 0
