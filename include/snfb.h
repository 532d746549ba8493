/*
 * snfb.h — C ABI of libsnfb200.so: the B200-native lead -> cluster -> consensus hot path
 * of Sniffles2, callable through ctypes from the Python host (sniffles_b200/binding.py).
 *
 * Every entry point replaces a Python call site of the reference (paths relative to
 * /root/reference/src/sniffles/):
 *
 *   snfb_load_records      <- pysam `bam.fetch()` iteration + AlignedSegment accessors
 *                             (leadprov.py:488, accessor list SURVEY.md §2)
 *   snfb_load_bam          <- the same call site fed COMPRESSED BAM bytes: htslib's BGZF inflate, record decode, region
 *                             filter and long-CIGAR (CG) escape behind `bam.fetch(contig, start, end)` run on the device
 *                             (parallel.py:95-98, leadprov.py:488)
 *   snfb_extract_leads     <- LeadProvider.build_leadtab / iter_region / read_iterindels /
 *                             Lead.for_bnd / read_itersplits (leadprov.py:445-670),
 *                             sv.classify_splits (sv.py:649-782); call site parallel.py:90-102
 *   snfb_cluster_call      <- cluster.resolve + merge_inner + resplit + resplit_bnd
 *                             (cluster.py:85-353), sv.call_from / resolve_bnd (sv.py:497-639),
 *                             postprocessing.coverage (postprocessing.py:69-130);
 *                             call site Task.call_candidates parallel.py:104-127
 *   snfb_consensus         <- postprocessing.annotate_sv INS branch (postprocessing.py:33-66)
 *                             + consensus.novel_from_reads (consensus.py:280-394);
 *                             call site Task.finalize_candidates parallel.py:145
 *   snfb_poa               <- spoa.poa as LocalAsm.assembly calls it (local_asm.py:287-291); call site parallel.py:186-196
 *   snfb_coverage_bins     <- SNFile.annotate_block_coverages' reshape-mean of lead_provider.coverage
 *                             (snf.py:248-267)
 *   snfb_allgather_candidates <- the parent collecting every worker's finished task results before VCF
 *                             emission (sniffles:544-547, parallel.py:270-271), as one NCCL all-gather
 *
 * Conventions: all functions return 0 on success, non-zero on error (message via
 * snfb_last_error).  No exceptions, no Python or torch types.  Views are library-owned
 * pinned host buffers, valid until the next call on the same ctx.  A ctx is bound to
 * one CUDA device and is not thread safe.  There is NO CPU fallback: if no CUDA device
 * is usable snfb_ctx_create fails.
 */
#ifndef SNFB_H
#define SNFB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNFB_ABI_VERSION 3

/* SV types in the reference's ALL_TYPES order (sv.py:31-33); the emission order of
 * candidates follows this order (parallel.py:106). */
enum { SNFB_INS = 0, SNFB_DEL = 1, SNFB_DUP = 2, SNFB_INV = 3, SNFB_BND = 4,
       SNFB_SINGLE_LEFT = 5, SNFB_SINGLE_RIGHT = 6, SNFB_NTYPES = 7 };
/* Lead.source (leadprov.py:46) */
enum { SNFB_SRC_INLINE = 0, SNFB_SRC_SPLIT_PRIM = 1, SNFB_SRC_SPLIT_SUP = 2, SNFB_SRC_BND_SA = 3 };

/* snfb_records.on_device */
#define SNFB_MEM_HOST 0u
#define SNFB_MEM_DEVICE 1u
#define SNFB_MEM_HOST_SEQ_ON_DEMAND 2u

/* snfb_records.cigar_fmt
 *
 * SNFB_CIGAR_BAM32: the BAM record's own words, len<<4|op (op: MIDNSHP=X = 0..8), 4 bytes per op.
 *
 * SNFB_CIGAR_16 (what the kernels read; snfb_pack_cigar16 produces it): 2-byte words, half the PCIe and HBM bytes.
 *   base word       bit 15 = 0, bit 14 = E, bits 11..13 = class, bits 0..10 = length & 0x7ff
 *                   class: 0 P (and the zero-length pad word 0x0000), 1 I, 2 D, 3 M/=/X, 4 H, 5 S, 6 N
 *                   (bit 11: the op advances the read, bit 12: it advances the reference)
 *                   E: the op is an I / D / S of at least the block's event length (snfb_records.cigar_evt_min, default
 *                   SNFB_CIGAR16_EVT_MIN = 11: every SV signature and every indel the NM correction counts).  The streaming kernel
 *                   only sums lengths and ORs this bit; a configuration that looks at shorter events re-flags the arena on the device.
 *   extension word  bit 15 = 1, bits 12..14 = level (1 or 2), bits 0..11 = payload: adds payload << (11 + 12 * (level - 1)) to the
 *                   length of the base word it follows (level 1, then level 2; lengths up to 2^28 as in BAM)
 *   A base word and its extension words never straddle a 16-byte boundary and every record starts on one; the gaps and
 *   the tail of the arena are filled with pad words, so a 16-byte load never needs masking.
 *   M, = and X are one class: the path never tells them apart (leadprov.py:137-142 OPLIST). */
#define SNFB_CIGAR_BAM32 0u
#define SNFB_CIGAR_16 1u
#define SNFB_CIGAR16_EVT_MIN 11u

/* aux_flags bits of snfb_rec */
#define SNFB_AUX_NM 1u
#define SNFB_AUX_HP 2u
#define SNFB_AUX_PS 4u
#define SNFB_AUX_SA 8u

/* One packed alignment record: the fields of a BAM record the path reads,
 * fixed 64 bytes so a warp fetches it with one coalesced request.  Base qualities are
 * never shipped.  Variable-length parts live in three arenas of the block:
 *   cigar[cigar_off .. +n_cigar)   CIGAR words in the block's cigar_fmt (SNFB_CIGAR_*)
 *   var[var_off .. +l_qname)       query name bytes (no NUL), followed by
 *   var[var_off+l_qname .. +sa_len) the SA:Z tag text (no NUL)
 *   seq[seq_off .. +(l_seq+1)/2)   BAM 4-bit bases ("=ACMGRSVTWYHKDBN"), high nibble first
 */
typedef struct snfb_rec {
    int32_t  task;        /* index into the task table (one task per contig/region)   */
    int32_t  pos;         /* 0-based reference_start                                    */
    uint16_t flag;        /* SAM flag                                                   */
    uint8_t  mapq;
    uint8_t  aux_flags;   /* SNFB_AUX_*                                                 */
    uint8_t  hp;          /* HP:i tag value (0 when absent); must be 0,1,2              */
    uint8_t  l_qname;
    uint16_t _pad0;
    int32_t  nm;          /* NM:i tag value                                             */
    int32_t  ps;          /* PS:i tag value                                             */
    uint32_t n_cigar;     /* words of this record (CIGAR16: pad words inside included)   */
    int32_t  l_seq;       /* query_length (bases stored in seq)                         */
    uint32_t sa_len;
    uint32_t _pad1;
    uint64_t cigar_off;   /* in words of the block's cigar_fmt (CIGAR16: a multiple of 8) */
    uint64_t seq_off;     /* in bytes                                                   */
    uint64_t var_off;     /* in bytes                                                   */
} snfb_rec;

/* One unit of work = what the reference hands to one CallTask (parallel.py:47-60):
 * a contig and one region on it.  Clusters never cross tasks. */
typedef struct snfb_task {
    int32_t contig;       /* index into the contig table                               */
    int32_t start;        /* region.start                                              */
    int32_t end;          /* region.end (exclusive)                                    */
    int32_t contig_len;   /* bam.get_reference_length(contig)                          */
    int32_t task_id;      /* Task.id (only used for ids on the host)                   */
    int32_t tr_off;       /* first tandem-repeat interval of this task in tr[]         */
    int32_t tr_n;         /* number of intervals (sorted by start, already padded)     */
    int32_t _pad;
} snfb_task;

/* Reference names: the SA tag names its mate contig as a string (leadprov.py:82);
 * the library resolves it through FNV-1a 64 hashes of the names and needs the
 * lexicographic rank of each name for util.most_common_top ties (util.py:101-103). */
typedef struct snfb_contig {
    uint64_t name_hash;   /* snfb_hash_name(name)                                      */
    int32_t  length;
    int32_t  lex_rank;    /* rank of the name among all contig names, byte-wise order  */
} snfb_contig;

typedef struct snfb_records {
    uint64_t n_rec;
    uint64_t n_cigar;     /* words in the cigar arena (CIGAR16: a multiple of 8, padded)  */
    uint64_t n_var;       /* bytes       */
    uint64_t n_seq;       /* bytes       */
    const snfb_rec* rec;
    const void*     cigar;   /* uint32_t[] (SNFB_CIGAR_BAM32) or uint16_t[] (SNFB_CIGAR_16) */
    const uint8_t*  var;
    const uint8_t*  seq;
    uint32_t n_task;
    uint32_t n_contig;
    uint32_t n_tr;
    uint32_t on_device;   /* SNFB_MEM_*: 0 host arenas (copied), 1 device arenas (no copy), 2 host arenas with the seq arena
                             left on the host: only the base slices the consensus stage asks for are fetched ("seq on demand") */
    const snfb_task*   task;
    const snfb_contig* contig;
    const int32_t*     tr;    /* n_tr pairs (start,end), per task sorted (util.py:121-147) */
    /* optional: reference 'N' runs per task for LeadProvider._mask_N_coverage (leadprov.py:420-443, only with --reference):
     * n_mask half-open pairs (start,end), sorted and disjoint inside a task; task t owns mask[mask_task_off[t] .. mask_task_off[t+1]) */
    uint32_t n_mask;
    uint32_t cigar_fmt;   /* SNFB_CIGAR_*; BAM32 host arenas are converted on the host inside snfb_load_records (device arenas must be CIGAR16) */
    const int32_t*  mask;
    const uint32_t* mask_task_off;   /* [n_task + 1]; may be NULL when n_mask == 0 */
    uint32_t cigar_evt_min;          /* CIGAR16: the event length the E bits were set for (0 = SNFB_CIGAR16_EVT_MIN) */
    uint32_t _pad2;
} snfb_records;

/* Flat POD of the reference's config values that the path reads (config.py:449-619). */
typedef struct snfb_config {
    int32_t mapq;                      /* config.py:533-534  */
    int32_t min_alignment_length;      /* config.py:535-536  */
    int32_t exclude_flags;             /* 0 = None           */
    int32_t minsvlen;                  /* config.py:508-514  */
    int32_t minsvlen_screen;           /* config.py:517      */
    int32_t long_ins_length;           /* 2500               */
    int32_t detect_large_ins;          /* bool               */
    int32_t dev_seq_cache_maxlen;      /* 50000              */
    int32_t max_splits_base;           /* 3                  */
    int32_t dev_keep_lowqual_splits;   /* bool               */
    int32_t qc_nm_measure;             /* config.py:596-599  */
    int32_t phase;                     /* bool               */
    int32_t cluster_binsize;           /* 100                */
    int32_t cluster_merge_pos;         /* 150                */
    int32_t cluster_merge_bnd;         /* 1000               */
    int32_t cluster_resplit_binsize;   /* 20                 */
    int32_t repeat;                    /* --repeat           */
    int32_t dev_min_leads_cluster;     /* config.py:607-611  */
    int32_t dev_no_resplit;
    int32_t dev_no_resplit_repeat;
    int32_t consensus_max_reads_bin;   /* 10                 */
    int32_t consensus_min_reads;       /* 4                  */
    int32_t consensus_kmer_len;        /* 6 (only 6 is supported on device) */
    int32_t consensus_kmer_skip_base;  /* 3                  */
    int32_t no_consensus;
    int32_t symbolic;
    int32_t precise;                   /* 25                 */
    int32_t coverage_binsize;          /* = cluster_binsize  */
    int32_t coverage_updown_bins;      /* 5                  */
    int32_t _pad;
    double  max_splits_kb;             /* 0.1                */
    double  cluster_r;                 /* 2.5                */
    double  cluster_repeat_h;          /* 1.5                */
    double  cluster_repeat_h_max;      /* 1000               */
    double  cluster_merge_len;         /* 0.22 / 0.27 mosaic */
    double  consensus_kmer_skip_seqlen_mult; /* 1/500        */
} snfb_config;

/* Lead as written by the device (64 B).  One lead = one reference `Lead`
 * (leadprov.py:34-55) that landed inside its task's region (leadprov.py:464-466). */
typedef struct snfb_lead {
    uint32_t rec;          /* index of the alignment record (stands in for read_id)   */
    int32_t  ref_start;
    int32_t  ref_end;
    int32_t  qry_start;
    int32_t  qry_end;
    int32_t  svlen;        /* undefined when SNFB_LF_SVLEN_NONE                        */
    int32_t  seq_off;      /* offset into the read's query_sequence, -1 = no sequence  */
    int32_t  seq_len;      /* length of the sequence slice                             */
    int32_t  read_len;     /* query_alignment_length for INLINE leads, else 0          */
    int32_t  mate_pos;     /* BND: bnd_info.mate_ref_start                             */
    int32_t  mate_contig;  /* BND: contig-table index of bnd_info.mate_contig          */
    int32_t  nm_sa;        /* BND: NM field of the SA entry (leadprov.py:119)          */
    uint32_t flags;        /* SNFB_LF_*                                                */
    uint16_t task;
    uint16_t k;            /* ordinal of the lead inside its read (A9 ordering)        */
    uint64_t qname_hash;   /* FNV-style 64-bit hash of query_name                      */
} snfb_lead;

#define SNFB_LF_TYPE(f)      ((f) & 7u)
#define SNFB_LF_SOURCE(f)    (((f) >> 3) & 3u)
#define SNFB_LF_REVERSE      (1u << 5)   /* strand == "-"                     */
#define SNFB_LF_IS_SA        (1u << 6)   /* read.is_supplementary             */
#define SNFB_LF_SVLEN_NONE   (1u << 7)   /* svlen is None ("long INS")        */
#define SNFB_LF_BND_FIRST    (1u << 8)
#define SNFB_LF_BND_REVERSE  (1u << 9)
#define SNFB_LF_HAS_SEQ      (1u << 10)  /* seq is not None                   */
#define SNFB_LF_NM_NONE      (1u << 11)  /* BND lead of a read without NM tag */
#define SNFB_LF_MAPQ(f)      (((f) >> 16) & 255u)
#define SNFB_LF_HAP(f)       (((f) >> 24) & 3u)

typedef struct snfb_lead_view {
    uint64_t n_leads;
    const snfb_lead* leads;        /* in bin order: (task, svtype, bin, record, k)     */
    uint64_t n_pass;               /* LeadProvider.read_count summed over tasks        */
    const uint32_t* task_read_count;  /* [n_task]                                       */
    const double*   task_mean_nm;     /* [n_task] config.average_regional_nm (leadprov.py:577) */
    const double*   rec_nm;        /* [n_rec] per-read nm (leadprov.py:524), -1 if none */
    uint64_t soft_errors;          /* malformed SA entries etc. (never fatal)          */
} snfb_lead_view;

/* One SV candidate = one reference SVCall as it leaves Task.call_candidates
 * (sv.py:561-598 + postprocessing.coverage), before QC/genotyping. */
typedef struct snfb_cand {
    int32_t  task;
    int32_t  svtype;
    int32_t  pos;
    int32_t  end;
    int32_t  svlen;
    int32_t  support;
    int32_t  qual;
    int32_t  precise;
    int32_t  fwd;
    int32_t  rev;
    int32_t  support_long;       /* INS: SUPPORT_LONG                                  */
    int32_t  support_sa;         /* DEL: SUPPORT_SA                                    */
    int32_t  cov_upstream, cov_start, cov_center, cov_end, cov_downstream;
    int32_t  hap_counts[6];      /* cluster.hap_counts (cluster.py:255-260)            */
    int32_t  sa_count;           /* cluster.sa_counts[0] (cluster.py:79-82)            */
    int32_t  sa_total;           /* denominator of sa_counts[1]                        */
    int32_t  bnd_mate_contig;
    int32_t  bnd_mate_pos;
    int32_t  bnd_is_first;
    int32_t  bnd_is_reverse;
    int32_t  n_strands;          /* len(set(lead.strand))                              */
    int32_t  support_inline;     /* distinct qnames among INLINE leads (sv.py:195)     */
    int32_t  lead_off;           /* first lead of this candidate in cand_leads[]       */
    int32_t  lead_n;
    int32_t  long_off;           /* INS: cluster.leads_long in cand_leads[]            */
    int32_t  long_n;
    int32_t  alt_off;            /* INS consensus: offset into alt[] (after snfb_consensus), -1 none */
    int32_t  alt_len;
    int32_t  hp_top, hp_support, hp_other;     /* phase_sv aggregates (postprocessing.py:626-654) */
    int32_t  ps_top, ps_top_null, ps_support, ps_other;
    int32_t  cluster_seed;       /* first bin of the merged cluster                    */
    int32_t  resplit_bin;        /* id suffix of cluster.resplit (cluster.py:151)      */
    double   stdev_pos;
    double   stdev_len;          /* NaN when absent (BND)                              */
    double   nm_mean;            /* -1 unless qc_nm_measure                            */
} snfb_cand;

typedef struct snfb_cand_view {
    uint64_t n_cand;
    const snfb_cand* cand;          /* in reference emission order: task, svtype, cluster */
    uint64_t n_cand_leads;
    const snfb_lead* cand_leads;    /* post merge_inner/resplit leads, reference order    */
    const uint64_t*  rnames;        /* distinct qname hashes per candidate, CSR below     */
    const uint32_t*  rnames_off;    /* [n_cand+1]                                          */
    const double*    task_coverage_mean; /* [n_task] coverage_average_total (postprocessing.py:130) */
    uint64_t unverified_breaks;     /* chain cuts whose independence check failed in the final run (0: a failed cut is redone with whole chains) */
} snfb_cand_view;

typedef struct snfb_seq_view {
    uint64_t n_alt_bytes;
    const uint8_t* alt;             /* ASCII, indexed by snfb_cand.alt_off/alt_len        */
} snfb_seq_view;

typedef struct snfb_ctx snfb_ctx;

int         snfb_version(void);
/* sizeof of the ABI structs, for binding self-checks: 0 rec, 1 task, 2 contig, 3 records, 4 config, 5 lead, 6 cand, 7 gather_view */
size_t      snfb_sizeof(int which);
uint64_t    snfb_hash_name(const char* s, size_t n);
int         snfb_ctx_create(int device, snfb_ctx** out);
void        snfb_ctx_destroy(snfb_ctx* ctx);
const char* snfb_last_error(snfb_ctx* ctx);
int         snfb_set_config(snfb_ctx* ctx, const snfb_config* cfg);
int         snfb_load_records(snfb_ctx* ctx, const snfb_records* block);
int         snfb_extract_leads(snfb_ctx* ctx, snfb_lead_view* out);   /* out may be NULL: stay on device */
int         snfb_cluster_call(snfb_ctx* ctx, snfb_cand_view* out);
int         snfb_consensus(snfb_ctx* ctx, snfb_seq_view* out);
/* all three stages back to back.  Sizes live in device counters and every buffer has a capacity kept in the ctx, so the
 * stream is never drained between stages: the host reads the counters once on a side stream (while the consensus kernels
 * run) to size the device -> host copies, and once at the end.  A run whose capacities were too small (the first run on
 * a ctx, or a block unlike the previous one) is repeated with capacities that fit; snfb_rerun_count counts those.
 * The views (any may be NULL) are filled after the final synchronisation. */
int         snfb_run(snfb_ctx* ctx, snfb_lead_view* leads, snfb_cand_view* cands, snfb_seq_view* seqs);
/* ---- device BAM ingest: compressed BGZF bytes in, the packed record block built in device memory ----
 * `bgzf` holds whole BGZF blocks back to back (host memory; any selection of a file's blocks, in file order).  A span is a
 * record-aligned range of the inflated stream that belongs to one task, given the way a BAI index gives it: (byte offset of a
 * BGZF block inside `bgzf`, offset inside that block's inflated data) for its begin and its end — i.e. a BAM virtual offset with
 * the file offset rebased to `bgzf`.  cend == n_bytes with uend == 0 means "to the end of the buffer".  Spans are listed task by
 * task in file order and must not overlap (merge the index chunks first, as htslib does); cutting a span at any record-aligned
 * offset (the linear index of the BAI provides one per 16 kb window) only adds parallelism.  The library inflates every block
 * (16 lanes of a warp per block), follows the block_size chain of every span, decodes the records, keeps those `bam.fetch(contig, start,
 * end)` would return for the span's task (task.contig is the BAM reference id), restores CIGARs of more than 65535 operations
 * from the CG:B,I tag, and writes snfb_rec + CIGAR16 + names/SA + 4-bit bases exactly as snfb_load_records expects them.
 * The BGZF CRC32 is not verified (a corrupt block is caught by the DEFLATE decoder or the inflated size).  Tables (tasks, contigs,
 * tandem repeats, N mask) have the meaning they have in snfb_records. */
typedef struct snfb_bam_span {
    uint64_t cbeg, cend;      /* byte offsets of BGZF block starts inside bgzf[] */
    uint32_t ubeg, uend;      /* offsets inside those blocks' inflated data */
    uint32_t task;
    uint32_t _pad;
} snfb_bam_span;
typedef struct snfb_bam_input {
    const uint8_t* bgzf; uint64_t n_bytes;
    const snfb_bam_span* span; uint64_t n_span;
    uint32_t n_task, n_contig, n_tr, n_mask;
    const snfb_task* task; const snfb_contig* contig; const int32_t* tr; const int32_t* mask; const uint32_t* mask_task_off;
} snfb_bam_input;
int         snfb_load_bam(snfb_ctx* ctx, const snfb_bam_input* in);
/* what the last snfb_load_bam built: out[0] records, out[1] CIGAR16 words, out[2] var bytes, out[3] seq bytes, out[4] raw records seen,
 * out[5] BGZF blocks, out[6] inflated bytes, out[7] compressed bytes */
int         snfb_ingest_sizes(snfb_ctx* ctx, uint64_t out[8]);
/* copies the block snfb_load_bam built back to the host (tests / inspection); any pointer may be NULL */
int         snfb_ingest_fetch(snfb_ctx* ctx, snfb_rec* rec, uint16_t* cigar16, uint8_t* var, uint8_t* seq);
/* inflate whole BGZF blocks on the device and return the inflated stream (tests / inspection).  Returns 0 and *out_len = bytes;
 * out may be NULL to get the size only. */
int         snfb_inflate_bgzf(snfb_ctx* ctx, const uint8_t* bgzf, uint64_t n_bytes, uint8_t* out, uint64_t out_cap, uint64_t* out_len);
/* device-time accounting of the last run: per-kernel milliseconds from CUDA events on
 * the ctx stream; names[i] is a static string.  Returns the number of entries. */
int         snfb_last_timings(snfb_ctx* ctx, const char** names, float* ms, uint64_t* bytes, int cap);
/* device pointer + count of the candidate buffer (for the NCCL all-gather done by the
 * Python host through torch.distributed) */
int         snfb_device_candidates(snfb_ctx* ctx, void** dptr, uint64_t* n_cand);
int         snfb_device_alt(snfb_ctx* ctx, void** dptr, uint64_t* n_bytes);
/* number of kernels launched by the library on this ctx since it was created */
uint64_t    snfb_launch_count(snfb_ctx* ctx);
/* number of times a run was repeated because a buffer capacity was too small or a chain cut had to be undone */
uint64_t    snfb_rerun_count(snfb_ctx* ctx);
/* mean coverage of consecutive `binsize`-base bins over the whole contig of one task, as the SNF writer stores it
 * (snf.py:248-267: the coverage vector zero-padded to a multiple of binsize, row means; the writer rounds them).
 * *out is a library-owned buffer of *n_bins doubles, valid until the next call on the ctx.  Needs snfb_extract_leads
 * (or snfb_run) first. */
int         snfb_coverage_bins(snfb_ctx* ctx, uint32_t task, int binsize, const double** out, uint64_t* n_bins);

/* ---- multi-GPU: one process per GPU, contigs sharded over the ranks, ONE all-gather of the per-rank candidate buffers ----
 * snfb_nccl_unique_id fills the 128 bytes of an ncclUniqueId on one rank; the host hands them to every rank (any transport),
 * then every rank calls snfb_comm_init.  snfb_allgather_candidates runs after snfb_run on every rank: it packs the rank's
 * candidate records, ALT arena, read names (and, with SNFB_GATHER_LEADS, the candidates' leads) into one buffer, all-gathers
 * the buffers over NCCL on the ctx stream, and returns the concatenation in rank order with lead_off / long_off / alt_off and
 * the rnames offsets rebased into the merged arrays (so a candidate of any rank indexes the merged arenas).  Ranks own
 * disjoint tasks; the caller orders by task id for emission (sniffles:544-547).  The view is library-owned pinned host
 * memory (dev_* are the same arrays in device memory), valid until the next call. */
#define SNFB_GATHER_LEADS 1u
#define SNFB_GATHER_DEVICE_ONLY 2u     /* leave the result in device memory (no device -> host copy; host pointers are NULL) */
typedef struct snfb_gather_view {
    uint64_t n_cand;        const snfb_cand* cand;
    uint64_t n_alt_bytes;   const uint8_t*  alt;
    uint64_t n_rnames;      const uint64_t* rnames;     const uint32_t* rnames_off;   /* [n_cand + 1] */
    uint64_t n_cand_leads;  const snfb_lead* cand_leads;                              /* 0 / NULL without SNFB_GATHER_LEADS */
    const uint64_t* rank_n_cand;                                                      /* [nranks] */
    const void* dev_buffer; uint64_t dev_bytes_per_rank;                              /* the gathered device buffer (nranks slots) */
} snfb_gather_view;
int         snfb_nccl_unique_id(void* out128);
int         snfb_comm_init(snfb_ctx* ctx, const void* unique_id128, int rank, int nranks);
int         snfb_allgather_candidates(snfb_ctx* ctx, uint32_t flags, snfb_gather_view* out);
/* ---- local assembly (LocalAsm, local_asm.py:254-304; gate parallel.py:186-196): the partial-order alignment the reference hands to pyspoa ----
 * A job is either mode 0: consensus of n_seq sequences = poa(read windows, local, min_coverage) (local_asm.py:287), or mode 1: the two-row
 * MSA of (sequence 0, sequence 1) = poa([consensus, ref], local, genmsa, m, n, g, e, q, c) (local_asm.py:289-291).  Sequences are bytes
 * compared for equality only; rows of an MSA use 255 for '-'.  out_len[k] = length / number of columns, -1 when the graph outgrew its
 * bounds, -2 when the job does not fit the scratch.  The algorithm is the one restated in oracle/poa_oracle.c (parity with pyspoa unpinned). */
typedef struct snfb_poa_job {
    uint64_t seq_off;        /* first byte of the job's sequences in seqs[]                                        */
    uint32_t offs_off;       /* index in offs[] of the job's n_seq + 1 offsets (relative to seq_off, ascending)    */
    uint32_t n_seq;
    int32_t  min_cov;        /* mode 0: round(0.5 n) (local_asm.py:285)                                            */
    int32_t  m, n, g, e, q, c;  /* match, mismatch, gap open / extend, second affine piece open / extend             */
    int32_t  band;           /* half width of the band around a node's column (>= the longest sequence: no band)   */
    uint32_t mode;
    uint32_t out_cap;        /* bytes per output row                                                               */
    uint64_t out_off;        /* mode 0: one row at out[out_off], mode 1: two rows of out_cap bytes                 */
} snfb_poa_job;
int         snfb_poa(snfb_ctx* ctx, const snfb_poa_job* jobs, uint32_t n_jobs, const uint8_t* seqs, uint64_t n_seq_bytes, const int32_t* offs, uint64_t n_offs,
                     uint8_t* out, uint64_t out_bytes, int32_t* out_len);
/* ---- multi-sample combine: the grouping of CombineTask.execute (parallel.py:443-572) -----------------------------------
 * The host reads the SNF blocks and lays the candidates out the way the reference visits them: a CHAIN is one (task, svtype) — its groups are
 * carried from chunk to chunk and from block to block (groups_keep, parallel.py:475,563) —, a CHUNK is one call of
 * cluster.resolve_block_groups (cluster.py:356-390): the candidates of the bins accumulated up to bin_max_candidates, already in the order
 * sorted(key=support, reverse=True) gives them.  The device runs every chain: nearest-group assignment with the reference's distance and
 * limits, SVGroup.from_candidate / add_candidate running means (sv.py:265-321), after each chunk the coverage of the samples a group does not
 * include (max over the chunks it lives through, parallel.py:538-552) and the keep / call split (parallel.py:554-557).
 * Outputs, per candidate: its group slot; per group slot (slot = chain.cand_off + order of creation): the chunk at whose end it was called
 * (n_chunk = kept to the end of the chain, -1 = unused slot), its position among the groups called then, the non-included coverages.
 * group.align_call (sv.py:282-292): with combine_pctseq != 0 and the ALT strings given, a candidate joins the nearest eligible group only if
 * (len_mean - editDistance(ALT of the group's first candidate, its ALT)) / len_mean > combine_pctseq — edlib.align's default global edit
 * distance, computed on the device (bit-vector blocks, one warp per pair).  combine_pctseq = 0 (or alt = NULL) is the reference's behaviour
 * without edlib / with --combine-pctseq 0: every pair passes. */
typedef struct snfb_combine_chain { uint32_t cand_off, n_cand, chunk_off, n_chunk, is_bnd, pad; } snfb_combine_chain;
typedef struct snfb_combine_chunk { int32_t cand_off, n_cand, curr_bin, size, cov_block, pad; } snfb_combine_chunk;   /* cov_block: row of cov[] of the block being read, -1 none */
typedef struct snfb_combine_in {
    uint32_t n_chain, n_chunk, n_cand, n_samples;
    const snfb_combine_chain* chains; const snfb_combine_chunk* chunks;
    const int32_t* pos; const int32_t* svlen; const uint32_t* sample;        /* per candidate; sample = index into the sample list   */
    const int32_t* mate_contig; const int32_t* mate_pos;                     /* BND chains only (any consistent contig numbering)     */
    uint32_t n_cov_block; int32_t bins_per_block, cov_binsize, pad;
    const int64_t* block_start;                                              /* [n_cov_block]                                         */
    const int32_t* cov;                                                      /* [n_cov_block][n_samples][bins_per_block]; -1 = the sample has no such block / key */
    int32_t combine_match, combine_match_max, cluster_merge_bnd, combine_separate_intra, combine_overlap_abs, pad2;
    double  combine_pctseq;
    const uint8_t* alt; const uint64_t* alt_off; const uint32_t* alt_len; uint64_t n_alt_bytes;   /* per candidate ALT bytes: alt[alt_off[i] .. + alt_len[i]) */
} snfb_combine_in;
typedef struct snfb_combine_out {
    uint32_t* cand_group;     /* [n_cand]                 */
    int32_t*  emit_chunk;     /* [n_cand] per group slot  */
    uint32_t* emit_ord;       /* [n_cand] per group slot  */
    int32_t*  cov_non;        /* [n_cand][n_samples]; -1 where the sample is included (never probed) */
} snfb_combine_out;
int         snfb_combine_groups(snfb_ctx* ctx, const snfb_combine_in* in, snfb_combine_out* out);
/* developer aid: with SNFB_DEBUG set in the environment the consensus alignment kernel records, per warp, busy cycles / elapsed cycles / items /
 * longest item (cycles, consensus length, read length); returns 1 when nothing was recorded */
int         snfb_debug_dump(snfb_ctx* ctx, uint64_t* out, uint64_t n_words);
/* self-check of the exact statistics.stdev arithmetic (host build of the routine the kernels use): the correctly rounded sqrt(P / Q) for
 * P = p_hi * 2^64 + p_lo; slow != 0 selects the limb-by-limb restatement of CPython's _float_sqrt_of_frac, 0 the verified fast path */
double      snfb_selftest_sqrt_frac(uint64_t p_hi, uint64_t p_lo, uint64_t q, int slow);
/* self-check of the device edit distance behind group.align_call: n_pairs pairs (a_off/a_len, b_off/b_len into bytes[]), distances to out[] */
int         snfb_selftest_edit_distance(snfb_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* a_off, const uint32_t* a_len, const uint64_t* b_off, const uint32_t* b_len, uint32_t n_pairs, int32_t* out);
/* BAM CIGAR words -> CIGAR16 (host code, OpenMP; no GPU needed).  rec_out receives copies of rec_in with cigar_off / n_cigar
 * rewritten for the 16-bit arena.  Call with out16 == NULL to get the number of 16-bit words the arena needs (a multiple
 * of 8); returns that number, or UINT64_MAX when a record holds an op the path does not know (B) or out_cap is too small.
 * evt_min: event length of the E bits (0 = SNFB_CIGAR16_EVT_MIN). */
uint64_t    snfb_pack_cigar16(const snfb_rec* rec_in, uint64_t n_rec, const uint32_t* cigar32, snfb_rec* rec_out, uint16_t* out16, uint64_t out_cap, uint32_t evt_min);
/* page-lock / unlock caller-owned host memory so that snfb_load_records copies at full PCIe rate */
int         snfb_pin_host(void* p, size_t bytes);
int         snfb_unpin_host(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SNFB_H */
