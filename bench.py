#!/usr/bin/env python
"""bench.py -- loci/sec of the hot path of `trgt genotype` (flank location + consensus alignment +
motif HMM) on synthetic 30x HiFi reads, BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W            this repo's CUDA engine (C ABI)
    python bench.py --impl reference --steps K --warmup W    the reference-equivalent CPU path
                                                             (oracle port: no rustc/cargo in the image)

One "step" = one pass of phases A+B+C over this rank's shard of the catalog.  Workload at N GPUs: the
Adotto-scale genome-wide synthetic catalog of BASELINE config 4, 125 000 loci per GPU (1 M loci at 8
GPUs), 30 reads per locus; loci are independent, so ranks shard by locus with no data-path collective
(weak scaling) and one gather of per-locus records at the end of each end-to-end step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "loci/sec (genome-wide synthetic catalog, 30x HiFi; phases A+B+C of trgt genotype)"
UNIT = "loci/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = pathogenic catalog (56 loci, 30x), 3 = 100k-locus catalog, uniform 2-6 bp "
                         "motifs, 20x (strong scaling: the catalog is split over the GPUs), 4 = Adotto-scale catalog, 30x "
                         "(125000 loci per GPU; the metric's config, default), 5 = long expansions: alleles of 5-50 kb, 40x, cluster "
                         "genotyper pass (1024 loci per GPU by default; --loci 10000 for the full catalog)")
    ap.add_argument("--loci", type=int, default=0, help="override the config's locus count (per GPU for config 4, total otherwise)")
    ap.add_argument("--depth", type=int, default=0, help="override the config's depth")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--resident-only", action="store_true",
                    help="profiling aid: skip the end-to-end driver so that every launch is a whole-shard launch")
    ap.add_argument("--chunk-loci", type=int, default=0,
                    help="loci per chunk of the end-to-end driver (0: two chunks per host thread)")
    ap.add_argument("--host-threads", type=int, default=0,
                    help="host threads (one engine each) of the e2e driver (0: host cores / ranks, between 2 and 8)")
    ap.add_argument("--upload-slots", type=int, default=2,
                    help="chunks that may be inside their phase-A call (the PCIe-heavy one) at a time; 0 = no limit")
    ap.add_argument("--align-by-index", type=int, default=0,
                    help="1: end to end, phase B names backbones and members by read index (trgt_align_trs) instead of uploading their bases")
    ap.add_argument("--guided", type=int, default=0, help="1: chunks shrink towards the end of the shard and go to whichever host thread is free")
    ap.add_argument("--min-chunk-loci", type=int, default=1500)
    ap.add_argument("--uploaders", type=int, default=0,
                    help="host threads that only upload phase A's inputs (resident-batch API), the others process; 0 = every "
                         "thread runs whole chunks through the one-shot calls")
    ap.add_argument("--max-inflight", type=int, default=4, help="chunks the uploaders may be ahead of the workers")
    ap.add_argument("--e2e-input", default="seq4", choices=["seq4", "ascii"],
                    help="how the e2e driver hands reads to phase A: BAM 4-bit bases (trgt_flank_spans_seq4) or ASCII")
    return ap.parse_args()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class Spec:
    """What one rank processes under --config, and how the catalog grows with the GPU count."""

    def __init__(self, args, world: int):
        self.config = args.config
        c = args.config
        self.depth = args.depth or {2: 30, 3: 20, 4: 30, 5: 40}[c]
        self.gen = {}
        self.motif_sets = None
        if c == 4:     # weak scaling: 125 000 loci per GPU, 1 M at 8 GPUs
            self.loci_per_rank = args.loci or 125000
            self.total = self.loci_per_rank * world
            self.scaling = "weak"
            self.name = "Adotto-scale genome-wide synthetic catalog shard (BASELINE config 4)"
            self.detail = "2-6 bp motifs (57/8/25/7/2 %), TR length median 24 bp"
        elif c == 5:   # weak scaling like config 4; the cluster-genotyper pass (harness/cluster_pass.py)
            self.loci_per_rank = args.loci or 1024
            self.total = self.loci_per_rank * world
            self.scaling = "weak"
            self.gen = {"tr_len_dist": "loguniform", "tr_len_min": 5000, "tr_len_max": 50000, "het_independent": True}
            self.name = "long-expansion stress catalog (BASELINE config 5), --genotyper cluster pass"
            self.detail = "2-6 bp motifs, allele lengths log-uniform on 5-50 kb (heterozygous loci draw two lengths)"
        elif c == 3:   # strong scaling: one 100k-locus catalog split over the GPUs
            self.total = args.loci or 100000
            self.loci_per_rank = -(-self.total // world)
            self.scaling = "strong"
            self.gen = {"motif_mix": "uniform"}
            self.name = "100k-locus synthetic catalog (BASELINE config 3)"
            self.detail = "2-6 bp motifs (uniform), TR length median 24 bp"
        else:          # the 56 loci of repeats/pathogenic_repeats.hg38.bed, split over the GPUs
            from harness import workload
            self.motif_sets = workload.pathogenic_motif_sets()
            self.total = args.loci or len(self.motif_sets)
            self.loci_per_rank = -(-self.total // world)
            self.scaling = "strong"
            self.gen = {"tr_len_median": 60.0}
            self.name = "pathogenic catalog (BASELINE config 2: motif sets of repeats/pathogenic_repeats.hg38.bed)"
            self.detail = "1-10 motifs per locus, up to 170 HMM states, TR length median 60 bp"

    def rank_range(self, rank: int):
        if self.scaling == "weak":
            return rank * self.loci_per_rank, self.loci_per_rank
        lo = min(self.total, rank * self.loci_per_rank)
        return lo, min(self.total, lo + self.loci_per_rank) - lo

    def generate(self, rank: int, n_loci=None, **kw):
        from harness import workload
        lo, cnt = self.rank_range(rank)
        return workload.generate(cnt if n_loci is None else min(n_loci, cnt), self.depth, locus_begin=lo,
                                 motif_sets=self.motif_sets, name=self.name, **self.gen, **kw)

    def describe(self, n_gpus: int) -> dict:
        per = f"{self.loci_per_rank} loci/GPU x {n_gpus} GPU" if self.scaling == "weak" else \
            f"{self.total} loci split over {n_gpus} GPU"
        return {
            "workload": f"{self.name}: {per}, {self.depth}x HiFi, clipped reads = 500 bp + allele + 500 bp, 250-bp "
                        f"flank pieces, {self.detail}; sub/ins/del 2e-4/4e-4/4e-4 per base",
            "baseline_config": self.config, "loci_per_gpu": self.loci_per_rank, "loci_total": self.total,
            "depth": self.depth, "scoring": [2, 5, 1], "min_flank_id_frac": 0.7,
            "parallelism": f"locus shards x{n_gpus}, no data-path collective; one gather of per-locus records per e2e step",
            "l2": "inputs (~1 KB per read; 3.9 GB of reads per GPU at config 4) are larger than the 126 MB L2 for configs 3 "
                  "and 4; no explicit flush" if self.config != 2 else
                  "inputs (1.9 MB of reads) fit the L2: a 256 MB buffer is overwritten between timed steps",
        }


# ------------------------------------------------------------------ clocks -----------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ reference arm ------------

def cpu_pass_rate(w, n_threads: int, seconds: float, max_loci: int):
    """Time the oracle pass on a bounded head of the workload sized for ~`seconds`.  -> (loci/s, n, dt)"""
    from oracle import oracle as orc
    from harness.pipeline import oracle_pass
    probe = min(max_loci, max(64, 16 * n_threads))
    t0 = time.perf_counter()
    oracle_pass(orc, w.head(probe), n_threads)
    dt = time.perf_counter() - t0
    rate = probe / dt
    n = int(min(max_loci, max(probe, rate * seconds)))
    t0 = time.perf_counter()
    ref = oracle_pass(orc, w.head(n), n_threads)
    dt = time.perf_counter() - t0
    return n / dt, n, dt, ref


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from harness.pipeline import oracle_pass
    cores = host_cores()
    spec = Spec(args, max(1, args.gpus))
    total = max(1, args.steps + args.warmup)
    per_step = max(1.0, min(8.0, 150.0 / total))
    # size the per-step sample with a probe
    w0 = spec.generate(0, max(64, 16 * cores))
    t0 = time.perf_counter()
    oracle_pass(orc, w0, cores)
    rate = w0.n_loci / (time.perf_counter() - t0)
    n = int(min(spec.loci_per_rank, max(w0.n_loci, rate * per_step)))
    w = spec.generate(0, n)
    n = w.n_loci
    for _ in range(args.warmup):
        oracle_pass(orc, w, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_pass(orc, w, cores)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = n / dt
    cfg = spec.describe(args.gpus)
    sample = f"first {n} loci of the shard per step ({w.n_reads} reads), all host cores"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": spec.scaling,
        "vs_baseline": None, "dtype": "int32 wavefront offsets + f64 Viterbi", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference-equivalent C restatement (oracle/): the Rust reference cannot be built here (no cargo/rustc, "
                "WFA2-lib and htslib un-vendored); one task per locus over a pthread pool as the reference's rayon pool",
    }))


# ------------------------------------------------------------------ roofline -----------------

def kernel_bytes(name: str, w, hp, res) -> float | None:
    """Algorithmic bytes one launch of `name` must move for this workload."""
    P = float(np.diff(w.left.offsets.astype(np.int64)).mean()) if len(w.left) else 250.0
    n_reads = w.n_reads
    read_len = np.diff(w.reads.offsets.astype(np.int64)).astype(np.float64)
    read_bytes = float(read_len.sum())
    g = res.glue
    if name == "k_unpack_seq4":   # half a byte in and one byte out per base, starts / lengths / offsets per read
        return 1.5 * read_bytes + 28.0 * n_reads
    if name in ("k_flank_exact", "k_flank_exact_t"):   # every read base once, both pieces once per locus, one hit per (read, flank)
        return read_bytes + 2.0 * P * w.n_loci + 2 * 20.0 * n_reads + 8.0 * n_reads
    if name in ("k_flank_band", "k_flank_seed", "k_flank_band1", "k_flank_band2", "k_flank_band_wide") and res.hits is not None:
        via = res.hits["via"].reshape(-1, 2)
        pend = (via >= 2)
        pend_reads = pend.any(axis=1)
        if name == "k_flank_seed":   # hit records scanned, both 8-mer tables per locus with a pending pair, the probes of
            # every pending pair (one 32-byte sector each, a probe every P/4 - 7 bases), one list entry per pair
            loci_pend = np.add.reduceat(pend_reads.astype(np.int64), w.locus_read_off[:-1].astype(np.int64)) > 0
            step = max(1.0, P / 4.0 - 7.0)
            n_probe_sectors = float((np.repeat(read_len / step, 2).reshape(-1, 2) * pend).sum())
            return float(40.0 * n_reads + 2.0 * 1344.0 * loci_pend.sum() + 32.0 * n_probe_sectors + 8.0 * pend.sum())
        if name == "k_flank_band1":  # per listed pair: list entry, the text window its band can touch, its piece, the hit
            return float(pend.sum() * (8.0 + (P + 16.0 + 12.0) + P + 20.0))
        if name == "k_flank_band":  # hit records scanned, pending reads re-read once, pieces once per locus, hits rewritten
            return float(40.0 * n_reads + read_len[pend_reads].sum() + 2.0 * P * w.n_loci + 20.0 * pend.sum())
        n_pairs = {"k_flank_band2": hp.fallback_counts()[0], "k_flank_band_wide": hp.fallback_counts()[1]}[name]
        # per pair handed on: its read and piece once, the piece's 8-mer table (1 KB), the work-list entry, the hit
        return float(n_pairs * (read_len.mean() + P + 1024.0 + 4.0 + 20.0))
    if name in ("k_hmm_viterbi_thread", "k_hmm_viterbi"):
        L = np.diff(g.backbones.offsets.astype(np.int64)).astype(np.float64)
        nm = np.diff(w.locus_motif_off.astype(np.int64))[g.group_locus]
        mlen = np.diff(w.motifs.offsets.astype(np.int64))
        first = w.locus_motif_off[:-1][g.group_locus]
        S = 7.0 + 3.0 * mlen[first] + 1.0          # one motif per locus in this catalog
        return float(((L + 2) * (1.0 + S) + mlen[first] + 8.0 * nm).sum())
    if name in ("k_hmm_lane_viterbi", "k_hmm_lane_walk", "k_hmm_lane_spans"):
        L = np.diff(g.backbones.offsets.astype(np.int64)).astype(np.float64)
        a = res.annotations
        if name == "k_hmm_lane_viterbi":   # every base once, one packed back-pointer word per column, motif + slot + status
            return float((L * (1.0 + 4.0)).sum() + 24.0 * L.size)
        if name == "k_hmm_lane_walk":      # the words once, MC / purity / span count / scratch spans out
            return float((4.0 * L).sum() + (4.0 + 8.0 + 4.0 + 8.0) * L.size + 12.0 * a.spans.shape[0])
        return float(24.0 * a.spans.shape[0] + 16.0 * L.size)
    if name in ("k_e2e_lane", "k_wfa_score_warp", "k_cigar_gather") and res.cigars is not None:
        c = res.cigars
        slen = np.diff(g.seqs.offsets.astype(np.int64)).astype(np.float64)
        blen = np.diff(g.backbones.offsets.astype(np.int64)).astype(np.float64)[
            np.repeat(np.arange(len(g.backbones)), np.diff(g.group_seq_off.astype(np.int64)))]
        nw = np.diff(c.offsets.astype(np.int64)).astype(np.float64)
        if name == "k_cigar_gather":       # every CIGAR word read from the pool and written in CSR order, end / offsets / score / status
            return float(8.0 * nw.sum() + (16.0 + 8.0 + 8.0 + 8.0) * nw.size)
        sel = (c.scores != 0) if name == "k_e2e_lane" else (c.scores < -8)   # members that differ / cost above the lane cap
        return float((slen[sel] + blen[sel] + 4.0 * nw[sel] + 16.0 + 4.0).sum())
    if name == "k_e2e_identity":  # every member and its backbone once, one end record and its CIGAR words per member
        per_seq_bb = np.diff(g.backbones.offsets.astype(np.int64))[
            np.repeat(np.arange(len(g.backbones)), np.diff(g.group_seq_off.astype(np.int64)))]
        return float(g.seqs.data.nbytes + per_seq_bb.sum() + 16.0 * len(g.seqs) + 4.0 * res.cigars.words.size)
    return None


# ------------------------------------------------------------------ main arm -----------------

def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    import trgt_b200
    from harness import workload
    from harness.pipeline import ChunkedHotPath, HotPath, compare_with_oracle, concat_results

    eng = trgt_b200.Engine(device=local_rank)  # fails loudly without the CUDA library / a GPU
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    spec = Spec(args, world)
    args.loci = spec.rank_range(rank)[1]   # loci of this rank
    t_gen = time.perf_counter()
    w = spec.generate(rank, alloc_reads=eng.pinned_array)
    t_gen = time.perf_counter() - t_gen
    use_seq4 = args.e2e_input == "seq4"
    if use_seq4:  # untimed set-up: the reads as BAM records hold them, in pinned memory
        w.pack_seq4(alloc=eng.pinned_array)
    hp = HotPath(eng, w, want_hits=True, pinned_outputs=True)  # hits: to count the fallback pairs for the roofline

    # end-to-end driver: chunks of loci through `host_threads` engines on this GPU
    if args.host_threads <= 0:
        args.host_threads = max(2, min(8, host_cores() // max(1, world)))
    if args.chunk_loci <= 0:
        args.chunk_loci = -(-args.loci // (2 * args.host_threads))
    engines = [eng] + [trgt_b200.Engine(device=local_rank) for _ in range(max(1, args.host_threads) - 1)]
    # host cores are shared by all ranks of the box and all host threads of a rank
    glue_threads = max(1, host_cores() // max(1, world * len(engines)))
    chp = ChunkedHotPath(engines, w, chunk_loci=args.chunk_loci, glue_threads=glue_threads, use_seq4=use_seq4,
                         upload_slots=args.upload_slots, uploaders=args.uploaders, max_inflight=args.max_inflight,
                         guided=bool(args.guided), min_chunk_loci=args.min_chunk_loci,
                         align_by_index=bool(args.align_by_index))

    # ---- warm-up: end-to-end passes (also builds the resident batches) ----
    res = None
    if not args.resident_only:
        for _ in range(max(1, args.warmup)):
            res = chp.run_e2e()
    hp.prepare_resident()
    for _ in range(max(1, args.warmup)):
        hp.run_resident()

    # ---- `value`: resident inputs, device-timed, per-kernel events on ----
    eng.reset_stats()
    eng.set_profiling(True)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = eng.launches()
    small = w.reads.data.nbytes < (256 << 20)   # inputs that fit the 126 MB L2: flush it between timed steps
    if small:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        pairs = []
        with torch.cuda.stream(stream):
            for _ in range(args.steps):
                flush_buf.zero_()               # untimed: overwrites the L2 with 256 MB
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                hp.run_resident(sync=False)
                b_.record(stream)
                pairs.append((a, b_))
        barrier()
        dev_ms = sum(a.elapsed_time(b_) for a, b_ in pairs) / max(1, args.steps)
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            hp.run_resident(sync=False)
        ev1.record(stream)
        barrier()
        dev_ms = ev0.elapsed_time(ev1) / max(1, args.steps)
    launches_total = eng.launches() - launches0                 # kernels of this engine inside the timed region
    launches = launches_total // max(1, args.steps)
    stats = eng.kernel_stats()
    unpack_stat = None
    if use_seq4:  # outside the timed region: one whole-shard launch of the e2e path's decode kernel, for its roofline entry
        fb4 = eng.flank_upload_seq4(w.left, w.right, w.reads4, w.locus_read_off, w.scoring, w.min_flank_id_frac)
        eng.flank_free(fb4)
        unpack_stat = eng.kernel_stats().get("k_unpack_seq4")
    eng.set_profiling(False)
    res_resident = hp.download_resident()
    dev_ms = max_over_ranks(dev_ms)
    total_loci = spec.total   # loci all ranks process per step
    value = total_loci / (dev_ms * 1e-3)

    # ---- `e2e`: host buffers through the C ABI, host<->device copies and host glue inside ----
    if args.resident_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "ms_per_step": dev_ms, "note": "resident-only profiling run", "e2e": None,
                              "kernels": {k: {"launches": n, "ms": ms} for k, (n, ms) in stats.items()}}))
        return
    from trgt_b200.shard import RecordGather, record_parts
    gatherer = RecordGather(dev)

    def gather_records(rs):
        if world == 1:
            return 0
        got = gatherer(record_parts(rs))
        return sum(g.size for g in got) if got is not None else 0

    for _ in range(2):   # warm-up: NCCL sets the collectives up lazily
        gather_records(res)

    barrier()
    tm0 = chp.timing()
    t_gather = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = chp.run_e2e()
        tg = time.perf_counter()
        gather_records(res)
        t_gather += time.perf_counter() - tg
    barrier()
    e2e_s = (time.perf_counter() - t0) / max(1, args.steps)
    tm1 = chp.timing()
    e2e_phases = {k: (tm1[k] - tm0[k]) / max(1, args.steps) * 1e3 for k in tm1}  # ms per step, summed over host threads
    e2e_phases["gather"] = t_gather / max(1, args.steps) * 1e3
    clk = clocks.stop()
    e2e_s = max_over_ranks(e2e_s)
    e2e_value = total_loci / e2e_s
    h2d = chp.h2d_bytes(res)
    d2h = chp.d2h_bytes(res)

    # the chunked end-to-end pass and the resident pass must agree on the full shard
    c_spans, c_words, c_scores, c_mc, c_hspans, c_pur = concat_results(res)
    assert np.array_equal(c_spans, res_resident.spans), "e2e and resident spans differ"
    assert np.array_equal(c_words, res_resident.cigars.words), "e2e and resident CIGARs differ"
    assert np.array_equal(c_scores, res_resident.cigars.scores)
    assert np.array_equal(c_hspans, res_resident.annotations.spans)
    assert np.array_equal(c_mc, res_resident.annotations.motif_counts)
    assert np.array_equal(c_pur, res_resident.annotations.purity, equal_nan=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kernels = {}
    for name, (n, ms) in stats.items():
        per = ms / max(1, n)
        b = kernel_bytes(name, w, hp, res_resident)
        kernels[name] = {"launches_per_step": n / max(1, args.steps), "ms_per_launch": per,
                         "share": None, "algorithmic_bytes": b,
                         "achieved_gbs": (b / (per * 1e-3) / 1e9) if (b and per > 0) else None}
    if unpack_stat and unpack_stat[0]:  # not part of the resident step (launches_per_step 0): runs per chunk in the e2e pass
        per = unpack_stat[1] / unpack_stat[0]
        b = kernel_bytes("k_unpack_seq4", w, hp, res_resident)
        kernels["k_unpack_seq4"] = {"launches_per_step": 0.0, "ms_per_launch": per, "share": None, "algorithmic_bytes": b,
                                    "achieved_gbs": b / (per * 1e-3) / 1e9 if per > 0 else None,
                                    "note": "e2e path only: one whole-shard launch timed outside the resident step"}
    tot_ms = sum(v["ms_per_launch"] * v["launches_per_step"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["ms_per_launch"] * v["launches_per_step"] / tot_ms if tot_ms else None
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"])
    d = kernels[dom]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except OSError:
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": (d["achieved_gbs"] / peak) if d["achieved_gbs"] else None, "traffic": traffic,
                "peak_source": peak_kind, "ms_per_launch": d["ms_per_launch"], "share_of_step": d["share"],
                "algorithmic_bytes_per_launch": d["algorithmic_bytes"]}

    # ---- CPU baseline + parity on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rate, n, dt, ref = cpu_pass_rate(w, cores, args.cpu_seconds, args.loci)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n} loci of the shard ({n * spec.depth} reads), one pass in {dt:.1f} s, all host cores"}
        small = HotPath(eng, w.head(n), want_hits=False, pinned_outputs=False, use_seq4=use_seq4)
        compare_with_oracle(small.run_e2e(), ref)
        parity = {"checked_loci": n, "result": "bit-exact vs oracle (spans, CIGARs, scores, MC, MS, AP)"}

    # ---- next row: consensus repair fused behind the alignments, timed on its own ----
    consensus = None
    if world == 1:
        try:
            g0 = hp.glue
            eng.consensus_packed(g0.backbones, g0.seqs, g0.group_seq_off)  # warm-up
            t0 = time.perf_counter()
            cons, cstat = eng.consensus_packed(g0.backbones, g0.seqs, g0.group_seq_off)
            dt = time.perf_counter() - t0
            consensus = {"call": "trgt_consensus (utils::align + repair_consensus, host buffers in and out)",
                         "groups": len(g0.backbones), "members": len(g0.seqs), "ms": dt * 1e3,
                         "groups_per_s": len(g0.backbones) / dt, "failed_groups": int((cstat != 0).sum())}
            if not args.no_cpu_baseline:
                from oracle import oracle as orc
                ng = min(len(g0.backbones), 2000)
                t0 = time.perf_counter()
                same = 0
                for gi in range(ng):
                    members = [g0.seqs.get(i) for i in range(int(g0.group_seq_off[gi]), int(g0.group_seq_off[gi + 1]))]
                    same += orc.repair_consensus(g0.backbones.get(gi), members) == cons.get(gi)
                consensus["parity"] = f"{same}/{ng} groups identical to the oracle"
                consensus["cpu_port_groups_per_s_1_core_python_driver"] = ng / (time.perf_counter() - t0)
        except Exception as exc:  # never let the extra row break the headline line
            consensus = {"error": repr(exc)}

    # ---- a14 (filter_impure_trs, tr.rs:400-452): the HMM on every spanning READ's repeat sequence, timed on its own ----
    a14 = None
    if world == 1:
        try:
            trs = eng.flank_trs(hp._fb, copy=True)
            read_locus = np.repeat(np.arange(w.n_loci, dtype=np.uint32), np.diff(w.locus_read_off.astype(np.int64)))
            eng.hmm_label_packed(w.motifs, w.locus_motif_off, trs, read_locus, copy=False)  # warm-up
            eng.reset_stats()
            eng.set_profiling(True)
            t0 = time.perf_counter()
            ann = eng.hmm_label_packed(w.motifs, w.locus_motif_off, trs, read_locus, copy=False)
            dt = time.perf_counter() - t0
            st = eng.kernel_stats()
            eng.set_profiling(False)
            a14 = {"call": "trgt_hmm_label on the repeat sequence of every read (D x per-read HMM of the targeted preset), "
                           "host buffers in and out", "sequences": len(trs), "ms": dt * 1e3, "sequences_per_s": len(trs) / dt,
                   "kernel_ms": {k: v[1] for k, v in st.items() if k.startswith("k_hmm")}}
            if not args.no_cpu_baseline:
                from oracle import oracle as orc
                nchk, same = 3000, 0
                pur = np.array(ann.purity[:nchk])
                for r in range(nchk):
                    h = orc.Hmm([orc.replace_invalid_bases(m, b"ATCGN") for m in w.locus_motifs(int(read_locus[r]))])
                    e_p = h.annotate(trs.get(r))[2]
                    same += (pur[r] == e_p) or (np.isnan(pur[r]) and np.isnan(e_p))
                a14["parity"] = f"{int(same)}/{nchk} purities identical to the oracle"
        except Exception as exc:  # never let the extra row break the headline line
            a14 = {"error": repr(exc)}

    # ---- next row: VCF sample fields encoded behind phase C, timed on its own ----
    # ---- next row (BAMlet): clip_bases for every read of the resident phase-A batch, timed on its own ----
    bamlet = None
    if world == 1 and getattr(hp, "_fb", None):
        try:
            rl = np.diff(w.reads.offsets.astype(np.int64)).astype(np.uint32)
            # CIGAR of every clipped read: 10 soft-clipped bases, the rest one '=' run (query length = read length)
            ops = np.empty(2 * w.n_reads, dtype=np.uint32)
            ops[0::2] = (10 << 4) | 4
            ops[1::2] = ((rl - 10) << 4) | 7
            ooff = np.arange(0, 2 * w.n_reads + 1, 2, dtype=np.uint64)
            refs = np.arange(w.n_reads, dtype=np.int64) * 1000
            eng.bamlet_clip(hp._fb, ops, ooff, refs, 250)  # warm-up
            t0 = time.perf_counter()
            bc = eng.bamlet_clip(hp._fb, ops, ooff, refs, 250)
            dt = time.perf_counter() - t0
            bamlet = {"call": "trgt_bamlet_clip (clip_bases of clip_bases.rs:9-119 as write_bam.rs:72-92 asks for it; reads and "
                              "spans resident, CIGARs in and clips out over PCIe)", "reads": int(w.n_reads), "ms": dt * 1e3,
                      "reads_per_s": w.n_reads / dt, "records": int((bc["status"] == 1).sum())}
            if not args.no_cpu_baseline:
                from oracle import oracle as orc
                spans_r, _ = eng.flank_download(hp._fb, w.n_reads, want_hits=False)
                nchk, same = min(w.n_reads, 2000), 0
                for r in range(nchk):
                    sp = (int(spans_r[r]["start"]), int(spans_r[r]["end"])) if spans_r[r]["found"] else None
                    exp = orc.bamlet_clip(w.reads.get(r), [int(ops[2 * r]), int(ops[2 * r + 1])], int(refs[r]), sp, 250) if sp else None
                    c = bc[r]
                    if exp is None:
                        same += int(c["status"]) == 0
                    else:
                        b0, b1, m0, m1, (rp, words) = exp
                        got_words = [int(c["first_word"])] if c["n_ops"] == 1 else [int(c["first_word"]), int(c["last_word"])]
                        same += (int(c["status"]) == 1 and (int(c["base_start"]), int(c["base_end"]), int(c["meth_start"]),
                                 int(c["meth_end"]), int(c["ref_pos"])) == (b0, b1, m0, m1, rp) and got_words == list(words))
                bamlet["parity"] = f"{same}/{nchk} reads identical to the oracle"
        except Exception as exc:  # never let the extra row break the headline line
            bamlet = {"error": repr(exc)}

    vcf = None
    if world == 1:
        try:
            eng.vcf_fields(hp._hb, raw=True)  # warm-up
            t0 = time.perf_counter()
            eng.vcf_fields(hp._hb, raw=True)
            dt = time.perf_counter() - t0
            fields = eng.vcf_fields(hp._hb)
            vcf = {"call": "trgt_vcf_fields (AL / MC / MS / AP of write_vcf.rs:267-343, device buffers in, host strings out)",
                   "records": len(fields), "ms": dt * 1e3, "records_per_s": len(fields) / dt,
                   "bytes": int(sum(len(x) for f in fields for x in f))}
            if not args.no_cpu_baseline:
                from oracle import oracle as orc
                a = res_resident.annotations
                nchk = min(len(fields), 2000)
                same = 0
                gl = hp.glue.group_locus  # alleles = backbones, grouped by locus
                lo = np.searchsorted(gl, np.arange(nchk), side="left")
                hi = np.searchsorted(gl, np.arange(nchk), side="right")
                for li in range(nchk):
                    als = []
                    for ai in range(int(lo[li]), int(hi[li])):
                        x = a.annotation(ai)
                        als.append((len(hp.glue.backbones.get(ai)), x.motif_counts, x.labels, x.purity))
                    same += fields[li] == orc.vcf_fields(als)
                vcf["parity"] = f"{same}/{nchk} records identical to the oracle"
        except Exception as exc:  # never let the extra row break the headline line
            vcf = {"error": repr(exc)}

    # ---- next row (first half): read clipping on synthetic BAM CIGARs, timed on its own ----
    clip = None
    if world == 1:
        try:
            rng = np.random.default_rng(7)
            n_cl, ops_per = 1_000_000, 40   # a 15 kb HiFi alignment has a few dozen CIGAR ops
            codes = rng.choice(np.array([7, 7, 7, 8, 1, 2], dtype=np.uint32), size=n_cl * ops_per)
            lens = np.where(codes == 7, rng.integers(50, 700, size=codes.size), rng.integers(1, 4, size=codes.size)).astype(np.uint32)
            ops = (lens << 4) | codes
            offs = (np.arange(n_cl + 1, dtype=np.uint64) * ops_per)
            n_loc = n_cl // 30
            lro = np.minimum(np.arange(n_loc + 1, dtype=np.uint64) * 30, n_cl).astype(np.uint32)
            lro[-1] = n_cl
            centre = rng.integers(4000, 9000, size=n_loc).astype(np.int64)
            regions = np.stack([centre - 500, centre + 530], axis=1)
            refs = rng.integers(0, 3000, size=n_cl).astype(np.int64)
            eng.clip_reads(ops, offs, refs, regions, lro)  # warm-up
            t0 = time.perf_counter()
            clips = eng.clip_reads(ops, offs, refs, regions, lro)
            dt = time.perf_counter() - t0
            clip = {"call": "trgt_clip_reads (clip_cigar of clip_region.rs:105-186, host buffers in and out)", "reads": n_cl,
                    "cigar_ops_per_read": ops_per, "ms": dt * 1e3, "reads_per_s": n_cl / dt,
                    "overlapping": int((clips["status"] == 1).sum())}
            if not args.no_cpu_baseline:
                from oracle import oracle as orc
                same = 0
                nchk = 2000
                read_locus = np.repeat(np.arange(n_loc), np.diff(lro.astype(np.int64)))
                for r in range(nchk):
                    o = ops[r * ops_per:(r + 1) * ops_per].tolist()
                    exp = orc.clip_cigar(o, int(refs[r]), (int(regions[read_locus[r], 0]), int(regions[read_locus[r], 1])))
                    c = clips[r]
                    if exp is None:
                        same += c["status"] == 0
                    else:
                        same += (c["status"] == 1 and (int(c["ref_start"]), int(c["query_start"]), int(c["query_end"])) == exp[:3]
                                 and int(c["n_ops"]) == len(exp[3]) and int(c["first_word"]) == exp[3][0]
                                 and int(c["last_word"]) == exp[3][-1])
                clip["parity"] = f"{int(same)}/{nchk} reads identical to the oracle"
        except Exception as exc:  # never let the extra row break the headline line
            clip = {"error": repr(exc)}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": spec.scaling, "vs_baseline": None,
        "dtype": "int32 wavefront offsets + f64 Viterbi", "data": "synthetic",
        "config": spec.describe(world), "clocks": clk, "gpu_launches": int(launches_total),
        "gpu_launches_per_step": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "chunk_loci": args.chunk_loci, "host_threads": len(engines),
                "upload_slots": args.upload_slots, "uploaders": args.uploaders, "guided": args.guided, "chunks": len(chp.paths), "align_by_index": args.align_by_index,
                "reads_in": "BAM 4-bit bases (trgt_flank_spans_seq4), decoded on the device" if use_seq4 else "ASCII (trgt_flank_spans)",
                "glue_threads": glue_threads, "phase_ms_summed_over_host_threads": e2e_phases},
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "consensus_row": consensus, "clip_row": clip, "vcf_row": vcf,
        "a14_row": a14, "bamlet_row": bamlet, "kernels": kernels,
        "wfa_fallback_pairs": hp.n_wfa(), "flank_fallback_counts": dict(zip(("second_tier", "wide_band", "full_width"), hp.fallback_counts())),
        "workload_gen_s": t_gen,
    }
    print(json.dumps(out))
    hp.free_resident()
    for e2 in engines:
        e2.close()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ config 5: the cluster-genotyper pass ----

def run_cluster(args):
    """BASELINE config 5.  One step = flank spans on 6-51 kb reads, spanning order, distance matrices + Ward clusters
    + central reads, consensus of both groups, HMM on the 5-50 kb alleles (harness/cluster_pass.py).  `value`: reads
    resident in HBM, CUDA events around the whole pass (the numpy glue between the phases included); `e2e`: the same
    pass from host buffers (BAM 4-bit bases uploaded inside the timed region).  The reference arm runs the pass on
    the oracle."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    spec = Spec(args, world if args.impl != "reference" else max(1, args.gpus))
    cores = host_cores()
    from harness.cluster_pass import compare_cluster_pass, engine_cluster_pass, oracle_cluster_pass
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import oracle as orc
        n = max(4, min(spec.loci_per_rank, 2 * cores))
        w = spec.generate(0, n)
        for _ in range(args.warmup):
            oracle_cluster_pass(orc, w, cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            oracle_cluster_pass(orc, w, cores)
        dt = (time.perf_counter() - t0) / max(1, args.steps)
        value = w.n_loci / dt
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": spec.scaling,
            "vs_baseline": None, "dtype": "int32 wavefront offsets + f64 Viterbi", "data": "synthetic",
            "config": spec.describe(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"first {w.n_loci} loci of the shard per step ({w.n_reads} reads), all host cores"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import trgt_b200
    from oracle import oracle as orc
    eng = trgt_b200.Engine(device=local_rank)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = spec.generate(rank, alloc_reads=eng.pinned_array)
    w.pack_seq4(alloc=eng.pinned_array)
    fb = eng.flank_upload(w.left, w.right, w.reads, w.locus_read_off, w.scoring, w.min_flank_id_frac)

    def resident_pass():
        return engine_cluster_pass(eng, w, orc_for_redo=orc, resident_batch=fb)

    for _ in range(max(1, args.warmup)):
        res = resident_pass()
    eng.reset_stats()
    eng.set_profiling(True)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = eng.launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        res = resident_pass()
    ev1.record(stream)
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1) / max(1, args.steps))
    launches_total = eng.launches() - launches0
    stats = eng.kernel_stats()
    eng.set_profiling(False)
    eng.flank_free(fb)
    # e2e: reads go up as BAM 4-bit bases inside the timed region
    engine_cluster_pass(eng, w, orc_for_redo=orc, use_seq4=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res_e2e = engine_cluster_pass(eng, w, orc_for_redo=orc, use_seq4=True)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / max(1, args.steps))
    clk = clocks.stop()
    assert np.array_equal(res_e2e.alleles.data, res.alleles.data), "e2e and resident alleles differ"
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = spec.total / (dev_ms * 1e-3)
    h2d = int(w.reads4.data.nbytes + w.reads4.starts.nbytes + w.reads4.lengths.nbytes + w.left.data.nbytes +
              w.right.data.nbytes + 4 * res.sel_reads.size * 2 + res.alleles.data.nbytes)
    d2h = int(res.spans.nbytes + 4 * res.sel_reads.size + res.alleles.data.nbytes + res.annotations.spans.nbytes)
    # roofline of the dominant kernel: bytes it must move (every pair's two sequences once for the aligner, a read
    # once for the flank search, a base and its back-pointer word once for the lane Viterbi)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    tr_len = (res.spans["end"].astype(np.int64) - res.spans["start"].astype(np.int64))
    sel_len = tr_len[res.sel_reads]
    allele_len = np.diff(res.alleles.offsets.astype(np.int64))
    bytes_of = {
        "k_wfa_score_block": float(2 * sel_len.sum() + 16 * sel_len.size),
        "k_wfa_trace": float(2 * sel_len.sum()),
        "k_flank_exact_t": float(w.reads.data.nbytes + 80.0 * w.n_reads),
        "k_hmm_lane_viterbi": float(5 * allele_len.sum()),
        "k_hmm_lane_walk": float(4 * allele_len.sum()),
        "k_trs_gather": float(2 * sel_len.sum()),
    }
    kernels = {}
    tot = sum(ms for _, ms in stats.values())
    for name, (n, ms) in stats.items():
        per = ms / max(1, n)
        b = bytes_of.get(name)
        kernels[name] = {"launches_per_step": n / max(1, args.steps), "ms_per_launch": per, "share": ms / tot if tot else None,
                         "algorithmic_bytes": b, "achieved_gbs": (b / (per * 1e-3) / 1e9) if (b and per > 0) else None}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_launch"] * kernels[k]["launches_per_step"])
    d = kernels[dom]
    roofline = {"kernel": dom, "bound": "hbm", "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": (d["achieved_gbs"] / peak) if d["achieved_gbs"] else None, "traffic": None,
                "ms_per_launch": d["ms_per_launch"], "share_of_step": d["share"],
                "algorithmic_bytes_per_launch": d["algorithmic_bytes"],
                "note": "wavefront alignment of ~20 kb pairs is compute / latency bound (O(L s) cells on a few hundred "
                        "diagonals); the byte figure is what the kernel must read"}
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        n = max(4, min(w.n_loci, 2 * cores))
        ws = w.head(n)
        t0 = time.perf_counter()
        ref = oracle_cluster_pass(orc, ws, cores)
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n} loci of the shard ({ws.n_reads} reads), one pass in {dt:.1f} s, all host cores"}
        compare_cluster_pass(engine_cluster_pass(eng, ws, orc_for_redo=orc, use_seq4=True), ref)
        parity = {"checked_loci": n, "result": "bit-exact vs oracle (spans, clusters, central reads, alleles, MC, MS, AP)"}
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": spec.scaling, "vs_baseline": None,
        "dtype": "int32 wavefront offsets + f64 Viterbi", "data": "synthetic", "config": spec.describe(world),
        "clocks": clk, "gpu_launches": int(launches_total), "gpu_launches_per_step": int(launches_total // max(1, args.steps)),
        "e2e": {"value": spec.total / e2e_s, "unit": UNIT, "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "reads_in": "BAM 4-bit bases (trgt_flank_upload_seq4), decoded on the device"},
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "kernels": kernels,
        "outlier_redo_loci": int(res.redone.size)}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.config == 5:
        run_cluster(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
