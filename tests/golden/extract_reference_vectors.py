"""Extracts the long literal test inputs of the reference's own unit tests into a JSON fixture.

Run in the build container (where /root/reference exists):
    python tests/golden/extract_reference_vectors.py
Writes tests/golden/wfa_long_vectors.json and tests/golden/pathogenic_motifs.json (the ID and
motif list of every locus of repeats/pathogenic_repeats.hg38.bed; workload shape for BASELINE
config 2).  The expected values come from the asserts of
/root/reference/src/wfaligner.rs (test_aligner_span_2 :1246-1261, test_invalid_sequence :1438-1454).
The GPU box has no /root/reference; tests only read the committed JSON.
"""
import json
import os
import re

SRC = "/root/reference/src/wfaligner.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wfa_long_vectors.json")


def grab(body: str, name: str) -> str:
    m = re.search(r"let %s = b\"([ACGTN]+)\";" % name, body)
    assert m, name
    return m.group(1)


def fn_body(src: str, fn: str) -> str:
    i = src.index("fn %s()" % fn)
    j = src.index("#[test]", i) if "#[test]" in src[i:] else len(src)
    return src[i:j]


def main():
    src = open(SRC).read()
    span2 = fn_body(src, "test_aligner_span_2")
    inv = fn_body(src, "test_invalid_sequence")
    out = {
        "span_2": {
            "source": "src/wfaligner.rs:1246-1261",
            "pattern": grab(span2, "pattern"),
            "text": grab(span2, "text"),
            "metric": "affine2p", "penalties": [8, 4, 2, 24, 1],
            "ends_free": [0, 0, 0, "len(text)"],
            "expect": {"xstart": 78, "xend": 250, "ystart": 0, "yend": 172},
        },
        "invalid_sequence": {
            "source": "src/wfaligner.rs:1438-1454",
            "pattern": grab(inv, "read"),
            "text": grab(inv, "allele"),
            "metric": "affine2p", "penalties": [8, 4, 2, 24, 1],
            "expect": {"score_heuristic_none": -881},
        },
    }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT, {k: (len(v["pattern"]), len(v["text"])) for k, v in out.items()})
    loci = []
    for line in open("/root/reference/repeats/pathogenic_repeats.hg38.bed"):
        f = line.rstrip("\n").split("\t")
        if len(f) < 4:
            continue
        info = dict(kv.split("=", 1) for kv in f[3].split(";"))
        loci.append({"id": info["ID"], "ref_len": int(f[2]) - int(f[1]), "motifs": info["MOTIFS"].split(",")})
    out2 = os.path.join(os.path.dirname(OUT), "pathogenic_motifs.json")
    with open(out2, "w") as fh:
        json.dump({"source": "repeats/pathogenic_repeats.hg38.bed", "loci": loci}, fh, indent=0)
    print("wrote", out2, len(loci), "loci")


if __name__ == "__main__":
    main()
