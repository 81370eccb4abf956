"""The C-ABI library loads without a GPU and exports every symbol include/minimp3_b200.h declares."""
import ctypes
import os
import re

import conftest  # noqa: F401
import minimp3_b200 as m
from minimp3_b200.build import build_lib


def declared_functions():
    text = open(os.path.join(conftest.ROOT, "include", "minimp3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(mp3dec_[a-z0-9_]+)\s*\(", text))
    names.discard("mp3dec_load_buf(")  # never matches; keeps the regex honest
    return sorted(names)


def test_library_builds_and_exports_header_symbols():
    build_lib()
    lib = m.load()
    names = declared_functions()
    assert "mp3dec_decode_batch_cuda" in names and "mp3dec_decode_frame" in names and "mp3dec_load_buf" in names
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.mp3dec_b200_version()


def test_struct_layouts_match_reference():
    # minimp3.h:13-23, minimp3_ex.h:38-43
    assert ctypes.sizeof(m.Mp3Dec) == 6668
    assert ctypes.sizeof(m.FrameInfo) == 24
    assert ctypes.sizeof(m.FileInfo) == 32
    assert m.Mp3Dec.qmf_state.offset == 2*288*4 and m.Mp3Dec.reserv.offset == (576 + 960)*4


def test_python_exports_cover_header():
    assert set(declared_functions()) <= set(m.EXPORTS)


def test_no_cpu_fallback_in_product():
    """The product must not import, include, link or execute anything under oracle/."""
    pat = re.compile(r"(from\s+oracle|import\s+oracle|oracle/|oracle\\|refshim|libmp3_oracle|libminimp3_ref)")
    for root, _, files in os.walk(os.path.join(conftest.ROOT, "minimp3_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert not pat.search(src), f


def build_c_example(out_dir, name="decode_file"):
    """examples/<name>.c as a maintainer would build it: plain C99, the public header, -lminimp3_b200."""
    import subprocess
    build_lib()
    exe = os.path.join(out_dir, name)
    libdir = os.path.dirname(m.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(conftest.ROOT, "include"),
                           os.path.join(conftest.ROOT, "examples", name + ".c"), "-L" + libdir, "-lminimp3_b200",
                           "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_c_callers_compile_against_the_header(tmp_path):
    assert os.path.exists(build_c_example(str(tmp_path)))
    assert os.path.exists(build_c_example(str(tmp_path), "decode_batch"))
