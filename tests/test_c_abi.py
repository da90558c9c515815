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


def test_rust_binding_declares_every_symbol():
    """integration/gpu_engine.rs (the binding a maintainer adds; not compiled here: no rustc) must declare every
    function of the header in its extern "C" block."""
    import re
    rs = open(os.path.join(os.path.dirname(HEADER), "..", "integration", "gpu_engine.rs")).read()
    declared = set(re.findall(r"pub fn (trgt_[a-z0-9_]+)\s*\(", rs))
    missing = [n for n in declared_functions() if n not in declared]
    assert not missing, missing


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
    got = {l.split()[0] + (" " + l.split()[1] if l.split()[0] in ("tr", "bamlet", "cigar", "allele") else ""): l for l in lines[4:]}
    assert got["tr 0"] == "tr 0 " + ("CAG" * 10) and got["tr 1"] == "tr 1 " + ("CAG" * 11) and got["tr 2"].strip() == "tr 2"
    # write_bam.rs:80-92: 50 bases of flank either side of the span; read 2 has no span -> no record
    assert got["bamlet 0"] == "bamlet 0 status 1 bases 300 430 ref 1300 ops 1 first 130="
    assert got["bamlet 1"] == "bamlet 1 status 1 bases 300 433 ref 2300 ops 1 first 133="
    assert got["bamlet 2"].startswith("bamlet 2 status 0 ")
    assert got["cigar 0"] == "cigar 0 score 0: 30="
    assert got["cigar 1"].startswith("cigar 1 score -8:") and got["cigar 2"] == "cigar 2 score -2: 11= 1X 18="
    assert got["consensus"] == "consensus " + "CAG" * 10
    assert got["dist"] == "dist 1.732051 1.000000 1.732051"
    assert got["allele 0"] == "allele 0 MC 11 MS 0(0-33) AP 1.000000" and got["allele 1"] == got["allele 0"].replace("allele 0", "allele 1")
    assert got["vcf"] == "vcf 33,33 11,11 0(0-33),0(0-33) 1.000000,1.000000"      # docs/tutorial.md:44
    assert got["clip"] == "clip status 1 ref 12 query 2 5 ops 3"                    # clip_region.rs:257-269
