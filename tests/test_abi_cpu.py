"""CPU-side checks of the C ABI: the library loads, exports every symbol include/bbb200.h declares,
and the ctypes structures have the C layout.  No device compute."""
import ctypes as C
import os
import re
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(REPO, "include", "bbb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bbb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bayesbay_b200 import _lib

    lib = _lib.load()
    declared = _declared()
    assert len(declared) >= 20
    assert set(declared) == set(_lib.SYMBOLS), "ctypes prototypes out of sync with include/bbb200.h"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bbb_abi_version() == _lib.ABI_VERSION
    assert lib.bbb_ray_tiles(10000) == (10000 + 63) // 64


def test_struct_layout_matches_c(tmp_path):
    from bayesbay_b200 import _lib

    prog = tmp_path / "sz.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "bbb200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(bbb_space_spec), sizeof(bbb_target_spec),"
        " sizeof(bbb_problem_spec), sizeof(bbb_buffers), offsetof(bbb_target_spec, G), offsetof(bbb_buffers, tape_cursor),"
        " offsetof(bbb_problem_spec, seed));return 0;}\n"
    )
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(_lib.SpaceSpec), C.sizeof(_lib.TargetSpec), C.sizeof(_lib.ProblemSpec), C.sizeof(_lib.Buffers),
            _lib.TargetSpec.G.offset, _lib.Buffers.tape_cursor.offset, _lib.ProblemSpec.seed.offset]
    assert got == want


def test_philox_known_answers_host():
    """The device generator's round function, run on the host through the ABI (Random123 vectors)."""
    from bayesbay_b200 import _lib

    lib = _lib.load()
    U4, U2 = C.c_uint32 * 4, C.c_uint32 * 2
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in cases:
        out = U4()
        assert lib.bbb_philox_kat(U4(*ctr), U2(*key), out) == 0
        assert tuple(out) == want


def test_invalid_arguments_are_reported_without_a_gpu():
    from bayesbay_b200 import _lib

    lib = _lib.load()
    assert lib.bbb_engine_create(None, 0, None) == -1
    assert b"NULL" in lib.bbb_last_error()
    spec = _lib.ProblemSpec()
    h = C.c_void_p()
    assert lib.bbb_engine_create(C.byref(spec), 0, C.byref(h)) == -1  # n_spaces == 0
    with pytest.raises(ValueError):
        _lib.check(-1)
