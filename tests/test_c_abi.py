"""The drop-in boundary is a C ABI: the header is valid C99, every function it declares is exported by
libtrgt_b200.so (and listed in trgt_b200.EXPORTS), a plain C caller links against it, and without a CUDA
device the product fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "trgt_engine.h")
DRIVER = os.path.join(ROOT, "tests", "c_abi", "driver.c")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(trgt_[a-z0-9_]+)\s*\(", text)))


def build_driver(tmp_path):
    import trgt_b200
    lib = trgt_b200._build.build()
    exe = str(tmp_path / "driver")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-o", exe, DRIVER, lib,
                    "-Wl,-rpath," + os.path.dirname(lib)], check=True)
    return exe


def test_header_is_c99():
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", "c", HEADER], check=True)


def test_library_exports_every_declared_symbol():
    import trgt_b200
    lib = trgt_b200.load_library()
    names = declared_functions()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(trgt_b200.EXPORTS) == names, set(names) ^ set(trgt_b200.EXPORTS)


def test_plain_c_caller_links_and_fails_loudly_without_gpu(tmp_path):
    exe = build_driver(tmp_path)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a CUDA device is present: the no-device path cannot be shown here")
    out = subprocess.run([exe, "nogpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "create rc=-4" in out.stdout and "no CUDA device" in out.stdout


@pytest.mark.gpu
def test_plain_c_caller_on_the_gpu(tmp_path):
    exe = build_driver(tmp_path)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    # read 0: both flanks exact; read 1: left flank through the WFA fallback; read 2: no right flank
    assert lines[0] == "span 0 1 350 380 via 1 1"
    assert lines[1] == "span 1 1 350 383 via 2 1"
    assert lines[2].startswith("span 2 0 ")
    assert int(lines[3].split()[1]) > 0
