"""Sanitizer run of the host parser (tests/fuzz_host_parse.cpp): the reference fuzzes its decoder with libFuzzer / AFL
(fuzzing/fuzz.c, scripts/fuzz.sh); here the product's host parser -- the code that sees untrusted bytes -- is built with
AddressSanitizer + UndefinedBehaviorSanitizer and driven by the harness's own deterministic mutation loop, seeded
with the reference's fuzz seed (vectors/fuzz/l3-compl-cut.mp3) and a spread of the bundled vectors."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "minimp3_b200", "csrc")
BIN = os.path.join(ROOT, "tests", "_build", "fuzz_host_parse")
SEEDS = ["fuzz-l3-compl-cut.mp3", "l3-compl.bit", "l3-he_free.bit", "l3-he_mode.bit", "l3-si_block.bit", "l2-fl10.bit",
         "l1-fl4.bit", "M2L3_bitrate_24_all.bit", "M2L3_compl24.bit", "l3-nonstandard-sin1k0db_lame_vbrtag.bit",
         "l3-nonstandard-id3v2.bit", "l3-nonstandard-id3v1-apetag.bit", "l3-nonstandard-vbrtag-oob-read.bit",
         "l3-nonstandard-sideinfo-size.bit", "l2-nonstandard-fl1_fl2_ff.bit", "ILL2_overalloc2.bit", "ILL4_wrong_length1.bit"]


def build():
    cxx = shutil.which("g++")
    if not cxx:
        pytest.skip("g++ not available")
    srcs = [os.path.join(ROOT, "tests", "fuzz_host_parse.cpp"), os.path.join(CSRC, "host_parse.cpp")]
    deps = srcs + [os.path.join(CSRC, h) for h in ("host_parse.h", "l3_sideinfo.cuh", "l12_decode.cuh", "l3_types.h")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call([cxx, "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                           "-fno-omit-frame-pointer", "-ffp-contract=off", "-DMP3B_FUZZ_STANDALONE", "-I", CSRC, "-x", "c++"]
                          + srcs + ["-o", BIN])


def test_host_parser_under_asan_ubsan():
    build()
    seeds = [os.path.join(ROOT, "tests", "golden", "vectors", s) for s in SEEDS]
    seconds = os.environ.get("MP3B_FUZZ_SECONDS", "30")
    p = subprocess.run([BIN, "--seconds", seconds] + seeds, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "no sanitizer report" in p.stdout
