//! gpu_engine.rs -- the reference-side binding of libtrgt_b200.so (include/trgt_engine.h).
//!
//! Drop this file into the reference crate as `src/gpu_engine.rs` (`mod gpu_engine;` in `src/main.rs`) and link
//! with `cargo:rustc-link-lib=dylib=trgt_b200`.  It was written against PacificBiosciences/trgt v3.0.0 and has
//! NOT been compiled in the environment this repository was built in (no cargo / rustc there); the Python
//! binding `trgt_b200/engine.py` drives the identical symbols through ctypes and is what the tests exercise.
//!
//! What it replaces, phase by phase, for a CHUNK of loci instead of one locus at a time:
//!   phase A   find_tr_spans            src/trgt/genotype/span_locater.rs:32-68   -> Engine::find_tr_spans / _seq4
//!   phase B   utils::align + repair    src/utils/align.rs:14-28, consensus.rs:5   -> Engine::consensus
//!             get_dist_matrix/cluster  src/trgt/genotype/genotype_cluster.rs      -> Engine::cluster
//!   phase C   label_with_hmm           src/trgt/workflows/tr.rs:454-492           -> Engine::label_with_hmm
//! The thread-local aligner trio of src/commands/genotype.rs:94-103 becomes one `Engine` per worker thread.
#![allow(dead_code)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_void};
use std::ptr;

// ---------------------------------------------------------------- raw declarations (include/trgt_engine.h) ----

#[repr(C)]
pub struct TrgtSeqs {
    pub data: *const u8,
    pub offsets: *const u64,
    pub n: u64,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct TrgtScoring {
    pub mismatch: i32,
    pub gap_open: i32,
    pub gap_extend: i32,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct TrgtSpan {
    pub found: i32,
    pub start: u32,
    pub end: u32,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct TrgtFlankHit {
    pub via: i32,
    pub matches: i32,
    pub score: i32,
    pub start: u32,
    pub end: u32,
}
#[repr(C)]
pub struct TrgtCigars {
    pub n: u64,
    pub offsets: *const u64,
    pub words: *const u32,
    pub scores: *const i32,
    pub status: *const i32,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct TrgtMotifSpan {
    pub motif_index: u32,
    pub start: u32,
    pub end: u32,
}
#[repr(C)]
pub struct TrgtAnnotations {
    pub n: u64,
    pub motif_count_offsets: *const u64,
    pub motif_counts: *const u32,
    pub span_offsets: *const u64,
    pub spans: *const TrgtMotifSpan,
    pub purity: *const f64,
    pub status: *const i32,
    pub path_offsets: *const u64,
    pub paths: *const u32,
}
#[repr(C)]
pub struct TrgtSeqsOut {
    pub n: u64,
    pub offsets: *const u64,
    pub data: *const u8,
    pub status: *const i32,
}
/// reads as the BAM record stores them: rust-htslib `record.seq().encoded`
#[repr(C)]
pub struct TrgtSeq4 {
    pub data: *const u8,
    pub data_bytes: u64,
    pub starts: *const u64,
    pub lengths: *const u32,
    pub n: u64,
}
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct TrgtClip {
    pub ref_start: i64,
    pub query_start: u64,
    pub query_end: u64,
    pub first_op: u32,
    pub n_ops: u32,
    pub first_word: u32,
    pub last_word: u32,
    pub status: i32,
}
pub enum TrgtEngine {}
/// `trgt_bamlet_clip_t`: what `HiFiRead::clip_bases` (clip_bases.rs:9-119) keeps of one read for the BAMlet
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct TrgtBamletClip {
    pub ref_pos: i64,
    pub base_start: u32,
    pub base_end: u32,
    pub meth_start: u32,
    pub meth_end: u32,
    pub first_op: u32,
    pub n_ops: u32,
    pub first_word: u32,
    pub last_word: u32,
    pub status: i32,
    pub pad: u32,
}
pub enum TrgtFlankBatch {}
pub enum TrgtAlignBatch {}
pub enum TrgtHmmBatch {}

#[link(name = "trgt_b200")]
extern "C" {
    pub fn trgt_engine_create(device: i32, out: *mut *mut TrgtEngine) -> i32;
    pub fn trgt_engine_destroy(eng: *mut TrgtEngine);
    pub fn trgt_engine_last_error(eng: *const TrgtEngine) -> *const c_char;
    pub fn trgt_last_create_error() -> *const c_char;
    pub fn trgt_engine_stream(eng: *mut TrgtEngine) -> *mut c_void;
    pub fn trgt_engine_sm_count(eng: *const TrgtEngine) -> i32;
    pub fn trgt_engine_sync(eng: *mut TrgtEngine) -> i32;
    pub fn trgt_engine_set_workspace_budget(eng: *mut TrgtEngine, bytes: usize);
    pub fn trgt_engine_set_flank_band_budget(eng: *mut TrgtEngine, max_cost: i32);
    // instrumentation (per-kernel device times through CUDA events on the engine stream; launch counter)
    pub fn trgt_engine_set_profiling(eng: *mut TrgtEngine, on: i32);
    pub fn trgt_engine_reset_stats(eng: *mut TrgtEngine);
    pub fn trgt_engine_kernel_count(eng: *mut TrgtEngine) -> i32;
    pub fn trgt_engine_kernel_stat(eng: *mut TrgtEngine, i: i32, name: *mut *const c_char, launches: *mut u64,
        total_ms: *mut f64) -> i32;
    pub fn trgt_engine_launches(eng: *const TrgtEngine) -> u64;
    pub fn trgt_engine_set_hmm_lane_path(eng: *mut TrgtEngine, on: i32);
    pub fn trgt_host_alloc(bytes: usize) -> *mut c_void;
    pub fn trgt_host_free(p: *mut c_void);

    pub fn trgt_clip_reads(eng: *mut TrgtEngine, cigar_ops: *const u32, cigar_offsets: *const u64, ref_starts: *const i64,
        n_reads: u64, regions: *const i64, locus_read_offsets: *const u32, n_loci: u32, clips_out: *mut TrgtClip) -> i32;
    pub fn trgt_flank_spans(eng: *mut TrgtEngine, left: *const TrgtSeqs, right: *const TrgtSeqs, reads: *const TrgtSeqs,
        locus_read_offsets: *const u32, n_loci: u32, scoring: TrgtScoring, min_flank_id_frac: f64,
        spans_out: *mut TrgtSpan, hits_out: *mut TrgtFlankHit) -> i32;
    pub fn trgt_flank_spans_seq4(eng: *mut TrgtEngine, left: *const TrgtSeqs, right: *const TrgtSeqs, reads: *const TrgtSeq4,
        locus_read_offsets: *const u32, n_loci: u32, scoring: TrgtScoring, min_flank_id_frac: f64,
        spans_out: *mut TrgtSpan, hits_out: *mut TrgtFlankHit) -> i32;
    pub fn trgt_bamlet_clip(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch, cigar_ops: *const u32,
                            cigar_offsets: *const u64, ref_starts: *const i64, flank_len: u32,
                            clips_out: *mut TrgtBamletClip) -> i32;
    pub fn trgt_seq4_decode(eng: *mut TrgtEngine, reads: *const TrgtSeq4, ascii_out: *mut u8, offsets_out: *mut u64) -> i32;
    pub fn trgt_flank_trs(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch, out: *mut TrgtSeqsOut) -> i32;

    pub fn trgt_align_e2e(eng: *mut TrgtEngine, backbones: *const TrgtSeqs, seqs: *const TrgtSeqs,
        group_seq_offsets: *const u32, n_groups: u32, out: *mut TrgtCigars) -> i32;
    pub fn trgt_consensus(eng: *mut TrgtEngine, backbones: *const TrgtSeqs, seqs: *const TrgtSeqs,
        group_seq_offsets: *const u32, n_groups: u32, out: *mut TrgtSeqsOut) -> i32;
    pub fn trgt_edit_dist(eng: *mut TrgtEngine, seqs: *const TrgtSeqs, locus_seq_offsets: *const u32, n_loci: u32,
        dists_out: *mut f64) -> i32;
    pub fn trgt_cluster(eng: *mut TrgtEngine, seqs: *const TrgtSeqs, locus_seq_offsets: *const u32, n_loci: u32,
        group_out: *mut i32, central_out: *mut u32, n_groups_out: *mut u32) -> i32;
    pub fn trgt_cluster_trs(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch, reads: *const u32, locus_offsets: *const u32,
        n_loci: u32, group_out: *mut i32, central_out: *mut u32, n_groups_out: *mut u32) -> i32;
    pub fn trgt_align_trs(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch, backbone_reads: *const u32,
        member_reads: *const u32, group_offsets: *const u32, n_groups: u32, out: *mut TrgtCigars) -> i32;
    pub fn trgt_consensus_trs(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch, backbone_reads: *const u32,
        member_reads: *const u32, group_offsets: *const u32, n_groups: u32, out: *mut TrgtSeqsOut) -> i32;

    pub fn trgt_hmm_label(eng: *mut TrgtEngine, motifs: *const TrgtSeqs, locus_motif_offsets: *const u32, n_loci: u32,
        alleles: *const TrgtSeqs, allele_locus: *const u32, want_paths: i32, out: *mut TrgtAnnotations) -> i32;
    pub fn trgt_vcf_fields(eng: *mut TrgtEngine, batch: *mut TrgtHmmBatch, out: *mut TrgtSeqsOut) -> i32;

    // resident batches: upload once, run many times (kernel time separated from PCIe time)
    pub fn trgt_flank_upload(eng: *mut TrgtEngine, left: *const TrgtSeqs, right: *const TrgtSeqs, reads: *const TrgtSeqs,
        locus_read_offsets: *const u32, n_loci: u32, scoring: TrgtScoring, min_flank_id_frac: f64,
        out: *mut *mut TrgtFlankBatch) -> i32;
    pub fn trgt_flank_upload_seq4(eng: *mut TrgtEngine, left: *const TrgtSeqs, right: *const TrgtSeqs, reads: *const TrgtSeq4,
        locus_read_offsets: *const u32, n_loci: u32, scoring: TrgtScoring, min_flank_id_frac: f64,
        out: *mut *mut TrgtFlankBatch) -> i32;
    pub fn trgt_flank_run(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch) -> i32;
    pub fn trgt_flank_download(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch, spans_out: *mut TrgtSpan,
        hits_out: *mut TrgtFlankHit) -> i32;
    pub fn trgt_flank_free(eng: *mut TrgtEngine, batch: *mut TrgtFlankBatch);
    // device pointers of a resident flank batch, for device-side consumers; how the last run settled the misses
    pub fn trgt_flank_device_views(batch: *mut TrgtFlankBatch, d_reads: *mut *const c_void, d_read_off: *mut *const c_void,
        d_spans: *mut *const c_void, d_hits: *mut *const c_void, n_reads: *mut u32, n_wfa: *mut u32) -> i32;
    pub fn trgt_flank_fallback_counts(batch: *mut TrgtFlankBatch, out: *mut u32) -> i32; // out[3]
    pub fn trgt_align_upload(eng: *mut TrgtEngine, backbones: *const TrgtSeqs, seqs: *const TrgtSeqs,
        group_seq_offsets: *const u32, n_groups: u32, out: *mut *mut TrgtAlignBatch) -> i32;
    pub fn trgt_align_run(eng: *mut TrgtEngine, batch: *mut TrgtAlignBatch) -> i32;
    pub fn trgt_align_download(eng: *mut TrgtEngine, batch: *mut TrgtAlignBatch, out: *mut TrgtCigars) -> i32;
    pub fn trgt_align_free(eng: *mut TrgtEngine, batch: *mut TrgtAlignBatch);
    pub fn trgt_hmm_upload(eng: *mut TrgtEngine, motifs: *const TrgtSeqs, locus_motif_offsets: *const u32, n_loci: u32,
        alleles: *const TrgtSeqs, allele_locus: *const u32, want_paths: i32, out: *mut *mut TrgtHmmBatch) -> i32;
    pub fn trgt_hmm_run(eng: *mut TrgtEngine, batch: *mut TrgtHmmBatch) -> i32;
    pub fn trgt_hmm_download(eng: *mut TrgtEngine, batch: *mut TrgtHmmBatch, out: *mut TrgtAnnotations) -> i32;
    pub fn trgt_hmm_free(eng: *mut TrgtEngine, batch: *mut TrgtHmmBatch);
}

// ---------------------------------------------------------------- packing --------------------------------------

/// A set of byte sequences as the engine takes them: concatenated bytes + CSR offsets.
#[derive(Default)]
pub struct Packed {
    pub data: Vec<u8>,
    pub offsets: Vec<u64>,
}

impl Packed {
    pub fn new() -> Self {
        Packed { data: Vec::new(), offsets: vec![0] }
    }
    pub fn push(&mut self, seq: &[u8]) {
        self.data.extend_from_slice(seq);
        self.offsets.push(self.data.len() as u64);
    }
    pub fn len(&self) -> usize {
        self.offsets.len() - 1
    }
    pub fn get(&self, i: usize) -> &[u8] {
        &self.data[self.offsets[i] as usize..self.offsets[i + 1] as usize]
    }
    fn raw(&mut self) -> TrgtSeqs {
        // the engine reads whole 16-byte words around a sequence: keep the tail padded
        let len = self.data.len();
        self.data.resize(len + 16, 0);
        self.data.truncate(len);
        TrgtSeqs { data: self.data.as_ptr(), offsets: self.offsets.as_ptr(), n: self.len() as u64 }
    }
}

pub type Result<T> = std::result::Result<T, String>; // src/utils/util.rs:3

/// One engine per worker thread (replaces THREAD_WFA_FLANK / _CONSENSUS / _ED, commands/genotype.rs:94-103).
pub struct Engine {
    raw: *mut TrgtEngine,
}

// calls on one engine serialise inside the library; the handle may move between threads
unsafe impl Send for Engine {}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { trgt_engine_destroy(self.raw) }
    }
}

/// `Annotation` of src/hmm/spans.rs:21-25, as `label_with_hmm` returns it
pub struct Annotation {
    pub labels: Option<Vec<(usize, usize, usize)>>, // (motif_index, start, end)
    pub motif_counts: Vec<usize>,
    pub purity: f64,
}

impl Engine {
    pub fn new(device: i32) -> Result<Self> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { trgt_engine_create(device, &mut raw) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(trgt_last_create_error()) }.to_string_lossy().into_owned();
            return Err(format!("trgt_engine_create failed ({}): {}", rc, msg)); // no CPU fallback
        }
        Ok(Engine { raw })
    }

    fn check(&self, rc: i32, what: &str) -> Result<()> {
        if rc == 0 {
            return Ok(());
        }
        let msg = unsafe { CStr::from_ptr(trgt_engine_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(format!("{} failed ({}): {}", what, rc, msg))
    }

    /// find_tr_spans (span_locater.rs:32-68) for a chunk of loci.  `lf_pieces[l]` = `lf[lf.len() - P..]`,
    /// `rf_pieces[l]` = `rf[..P]` (span_locater.rs:38-39); `reads` are the clipped reads of all loci, locus after
    /// locus; `locus_read_offsets[l]..[l+1]` delimits locus l.  One `Option<(usize, usize)>` per read.
    pub fn find_tr_spans(&self, lf_pieces: &mut Packed, rf_pieces: &mut Packed, reads: &mut Packed,
                         locus_read_offsets: &[u32], scoring: TrgtScoring, min_flank_id_frac: f64)
                         -> Result<Vec<Option<(usize, usize)>>> {
        let n_loci = lf_pieces.len() as u32;
        let mut spans = vec![TrgtSpan::default(); reads.len()];
        let (l, r, t) = (lf_pieces.raw(), rf_pieces.raw(), reads.raw());
        let rc = unsafe {
            trgt_flank_spans(self.raw, &l, &r, &t, locus_read_offsets.as_ptr(), n_loci, scoring, min_flank_id_frac,
                             spans.as_mut_ptr(), ptr::null_mut())
        };
        self.check(rc, "trgt_flank_spans")?;
        Ok(spans.iter().map(|s| if s.found != 0 { Some((s.start as usize, s.end as usize)) } else { None }).collect())
    }

    /// The same on reads that were never decoded: `seq4` holds the bytes of `record.seq().encoded` that cover
    /// `bases[query_start..query_end)` of each record (trgt_clip_reads gives the range), `starts[i] = 2 * byte_offset +
    /// (query_start & 1)`.  Afterwards `flank_trs` returns `trs` of tr.rs:58-62.
    pub fn find_tr_spans_seq4(&self, lf_pieces: &mut Packed, rf_pieces: &mut Packed, seq4: &[u8], starts: &[u64],
                              lengths: &[u32], locus_read_offsets: &[u32], scoring: TrgtScoring,
                              min_flank_id_frac: f64) -> Result<Vec<Option<(usize, usize)>>> {
        let n_loci = lf_pieces.len() as u32;
        let mut spans = vec![TrgtSpan::default(); starts.len()];
        let (l, r) = (lf_pieces.raw(), rf_pieces.raw());
        let t = TrgtSeq4 { data: seq4.as_ptr(), data_bytes: seq4.len() as u64, starts: starts.as_ptr(),
                           lengths: lengths.as_ptr(), n: starts.len() as u64 };
        let rc = unsafe {
            trgt_flank_spans_seq4(self.raw, &l, &r, &t, locus_read_offsets.as_ptr(), n_loci, scoring, min_flank_id_frac,
                                  spans.as_mut_ptr(), ptr::null_mut())
        };
        self.check(rc, "trgt_flank_spans_seq4")?;
        Ok(spans.iter().map(|s| if s.found != 0 { Some((s.start as usize, s.end as usize)) } else { None }).collect())
    }

    /// `trs` of tr.rs:58-62 for the reads of the last find_tr_spans* call: one repeat sequence per read (empty
    /// when the read does not span).
    pub fn flank_trs(&self) -> Result<Packed> {
        let mut out = TrgtSeqsOut { n: 0, offsets: ptr::null(), data: ptr::null(), status: ptr::null() };
        let rc = unsafe { trgt_flank_trs(self.raw, ptr::null_mut(), &mut out) };
        self.check(rc, "trgt_flank_trs")?;
        Ok(unsafe { copy_seqs_out(&out) })
    }

    /// `align` + `repair_consensus` (utils/align.rs:14-28, consensus.rs:5-72) for many (backbone, seqs) groups --
    /// what genotype_size.rs:35-36, genotype_cluster.rs:52-53 and genotype_flank.rs:19-20 chain.
    pub fn consensus(&self, backbones: &mut Packed, seqs: &mut Packed, group_seq_offsets: &[u32]) -> Result<Vec<String>> {
        let mut out = TrgtSeqsOut { n: 0, offsets: ptr::null(), data: ptr::null(), status: ptr::null() };
        let (b, s) = (backbones.raw(), seqs.raw());
        let rc = unsafe { trgt_consensus(self.raw, &b, &s, group_seq_offsets.as_ptr(), backbones.len() as u32, &mut out) };
        self.check(rc, "trgt_consensus")?;
        let p = unsafe { copy_seqs_out(&out) };
        Ok((0..p.len()).map(|i| String::from_utf8_lossy(p.get(i)).into_owned()).collect())
    }

    /// `align` alone: run-length SAM words per sequence, decoded by WFAligner::decode_sam_cigar (wfaligner.rs:961)
    pub fn align(&self, backbones: &mut Packed, seqs: &mut Packed, group_seq_offsets: &[u32]) -> Result<Vec<Vec<u32>>> {
        let mut out = TrgtCigars { n: 0, offsets: ptr::null(), words: ptr::null(), scores: ptr::null(), status: ptr::null() };
        let (b, s) = (backbones.raw(), seqs.raw());
        let rc = unsafe { trgt_align_e2e(self.raw, &b, &s, group_seq_offsets.as_ptr(), backbones.len() as u32, &mut out) };
        self.check(rc, "trgt_align_e2e")?;
        let n = out.n as usize;
        let offs = unsafe { std::slice::from_raw_parts(out.offsets, n + 1) };
        Ok((0..n).map(|i| unsafe {
            std::slice::from_raw_parts(out.words.add(offs[i] as usize), (offs[i + 1] - offs[i]) as usize).to_vec()
        }).collect())
    }

    /// genotype_cluster::genotype up to its make_consensus calls (:57-72) for many loci: per sequence 0 = group1,
    /// 1 = group2, 2 = neither; per locus the backbones (central_read) of both groups.
    pub fn cluster(&self, trs: &mut Packed, locus_seq_offsets: &[u32]) -> Result<(Vec<i32>, Vec<[Option<usize>; 2]>)> {
        let n_loci = locus_seq_offsets.len() - 1;
        let mut group = vec![0i32; trs.len().max(1)];
        let mut central = vec![0u32; 2 * n_loci.max(1)];
        let t = trs.raw();
        let rc = unsafe {
            trgt_cluster(self.raw, &t, locus_seq_offsets.as_ptr(), n_loci as u32, group.as_mut_ptr(), central.as_mut_ptr(),
                         ptr::null_mut())
        };
        self.check(rc, "trgt_cluster")?;
        group.truncate(trs.len());
        let c = (0..n_loci).map(|l| {
            let f = |v: u32| if v == u32::MAX { None } else { Some(v as usize) };
            [f(central[2 * l]), f(central[2 * l + 1])]
        }).collect();
        Ok((group, c))
    }

    /// label_with_hmm (tr.rs:454-492) for many loci; also serves filter_impure_trs (tr.rs:400-452) when `alleles`
    /// are the reads' repeat sequences.  replace_invalid_bases is applied inside.
    pub fn label_with_hmm(&self, motifs: &mut Packed, locus_motif_offsets: &[u32], alleles: &mut Packed,
                          allele_locus: &[u32]) -> Result<Vec<Annotation>> {
        let mut out: TrgtAnnotations = unsafe { std::mem::zeroed() };
        let (m, a) = (motifs.raw(), alleles.raw());
        let rc = unsafe {
            trgt_hmm_label(self.raw, &m, locus_motif_offsets.as_ptr(), (locus_motif_offsets.len() - 1) as u32, &a,
                           allele_locus.as_ptr(), 0, &mut out)
        };
        self.check(rc, "trgt_hmm_label")?;
        let n = out.n as usize;
        let (mco, so) = unsafe {
            (std::slice::from_raw_parts(out.motif_count_offsets, n + 1), std::slice::from_raw_parts(out.span_offsets, n + 1))
        };
        Ok((0..n).map(|i| unsafe {
            let spans = std::slice::from_raw_parts(out.spans.add(so[i] as usize), (so[i + 1] - so[i]) as usize);
            let mc = std::slice::from_raw_parts(out.motif_counts.add(mco[i] as usize), (mco[i + 1] - mco[i]) as usize);
            Annotation {
                labels: if spans.is_empty() { None } else {
                    Some(spans.iter().map(|s| (s.motif_index as usize, s.start as usize, s.end as usize)).collect())
                },
                motif_counts: mc.iter().map(|&c| c as usize).collect(),
                purity: *out.purity.add(i),
            }
        }).collect())
    }

    /// AL / MC / MS / AP of every locus of the last label_with_hmm call, as write_vcf.rs:267-343 encodes them
    pub fn vcf_fields(&self) -> Result<Vec<[String; 4]>> {
        let mut out = TrgtSeqsOut { n: 0, offsets: ptr::null(), data: ptr::null(), status: ptr::null() };
        let rc = unsafe { trgt_vcf_fields(self.raw, ptr::null_mut(), &mut out) };
        self.check(rc, "trgt_vcf_fields")?;
        let p = unsafe { copy_seqs_out(&out) };
        let s = |i: usize| String::from_utf8_lossy(p.get(i)).into_owned();
        Ok((0..p.len() / 4).map(|l| [s(4 * l), s(4 * l + 1), s(4 * l + 2), s(4 * l + 3)]).collect())
    }
}

unsafe fn copy_seqs_out(out: &TrgtSeqsOut) -> Packed {
    let n = out.n as usize;
    if n == 0 {
        return Packed::new();
    }
    let offsets = std::slice::from_raw_parts(out.offsets, n + 1).to_vec();
    let total = offsets[n] as usize;
    let data = if total > 0 { std::slice::from_raw_parts(out.data, total).to_vec() } else { Vec::new() };
    Packed { data, offsets }
}
