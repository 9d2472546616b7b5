"""The drop-in boundary from C: a host program written against the SDK's API names (tests/c_host/app.c) compiles
against include/lzb200.h and LINKS with -llzb200 in place of the SDK's sources (the link-line swap).
CPU: compile + link must resolve every symbol.  GPU: run it and compare with the oracle's answers."""
import os
import subprocess
import tempfile
import zlib

import numpy as np
import pytest

import cases
import lzb200

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build_app(td):
    so = lzb200.buildlib.build()
    exe = os.path.join(td, "app")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(HERE, "c_host", "app.c"), "-L", os.path.dirname(so), "-llzb200",
                           "-Wl,-rpath," + os.path.dirname(so), "-o", exe])
    return exe


def test_c_host_compiles_and_links_against_the_library():
    with tempfile.TemporaryDirectory() as td:
        exe = build_app(td)
        # every SDK symbol the program uses is provided by liblzb200.so (no SDK object files on the link line)
        undefined = subprocess.check_output(["nm", "-u", exe], text=True)
        for sym in ("LzmaUncompress", "LzmaDec_Allocate", "LzmaDec_DecodeToBuf", "LzmaDec_Free", "LzmaDecode", "CrcCalc"):
            assert sym in undefined
        assert "liblzb200.so" in subprocess.check_output(["ldd", exe], text=True)


@pytest.mark.gpu
def test_c_host_runs_and_matches_the_oracle(ora):
    blobs, cs = cases.golden()
    c = max((x for x in cs if x["kind"] == "lzma" and x["mut"] is None and x["finish"] == 0 and x["res"] == 0 and x["dest_len"] == x["cap"]),
            key=lambda x: x["cap"])
    packed, props = blobs[c["blob"]], bytes.fromhex(c["props"])
    want = ora.lzma_decode(packed, props, c["cap"], 1)
    crc = "%08x" % zlib.crc32(want["out"][: want["dest_len"]].tobytes())
    with tempfile.TemporaryDirectory() as td:
        exe = build_app(td)
        path = os.path.join(td, "in.bin")
        np.asarray(packed, dtype=np.uint8).tofile(path)
        out = subprocess.check_output([exe, path, props.hex(), str(c["cap"])], text=True).splitlines()
    assert len(out) == 3 and all(crc in line for line in out), out
    assert out[0].startswith("LzmaUncompress res=0 destLen=%d" % c["cap"])
    assert out[1].startswith("LzmaDec_DecodeToBuf res=0") and "out=%d" % c["cap"] in out[1]
    assert out[2].startswith("LzmaDecode res=%d status=%d destLen=%d" % (want["res"], want["status"], want["dest_len"]))
