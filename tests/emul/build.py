"""Builds tests/emul/libtrgt_emul.so: the kernel cores of trgt_b200/csrc instantiated with the one-lane
SerialGroup (TEST INFRASTRUCTURE ONLY -- lets `-m "not gpu"` tests check the cores' index arithmetic
and tie-breaking against the oracle on a machine without a GPU; never part of the product)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libtrgt_emul.so")


def build() -> str:
    srcs = [os.path.join(_HERE, f) for f in ("emul_hmm.cpp", "emul_wfa.cpp", "emul_consensus.cpp", "emul_clip.cpp", "emul_cluster.cpp")]
    csrc = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "trgt_b200", "csrc")
    deps = srcs + [os.path.join(_HERE, "lanes.h")] + [os.path.join(csrc, f) for f in
                                                     ("coop.h", "hmm_core.h", "hmm_host.h", "wfa_core.h", "consensus_core.h", "clip_core.h", "vcf_core.h", "cluster_core.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-pthread", "-o", LIB] + srcs,
                   check=True)
    return LIB
