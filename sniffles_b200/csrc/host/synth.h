/*
 * synth.h — seeded synthetic alignment-record generator (host only, test/bench input).
 * Emits the packed record block of include/snfb.h directly, following the shapes of
 * ONT / HiFi op models, planted INS/DEL/DUP/INV/BND sites with
 * mutually consistent SA tags, soft clips, low-mapq and secondary reads, HP/PS tags,
 * tandem-repeat intervals.  Every read depends only on (seed, contig, read index), so
 * the block is identical for any thread count.
 */
#ifndef SNFB_SYNTH_H
#define SNFB_SYNTH_H
#include <stdint.h>
#include "../../../include/snfb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct snfb_synth_params {
    uint64_t seed;
    int32_t  n_contig;
    int32_t  len_model;        /* 0: normal(len_mean, len_sd); 1: lognormal(mean=len_mean, sigma=len_sd/1000) */
    const int32_t* contig_len; /* [n_contig] */
    double   coverage;         /* mean depth; reads per contig = coverage*len/len_mean */
    double   len_mean;
    double   len_sd;
    int32_t  len_min;
    int32_t  len_max;
    double   op_mean_run;      /* mean M-run between small indels: 80 ONT, 700 HiFi */
    double   nm_rate;          /* substitution rate feeding NM: 0.01 ONT, 0.001 HiFi */
    double   clip_prob;        /* 0.10 */
    double   lowmapq_prob;     /* 0.05 */
    double   secondary_prob;   /* 0.02 */
    double   sv_spacing;       /* one planted site per this many bp (120000) */
    double   phased_frac;      /* fraction of reads carrying HP/PS */
    double   tr_frac;          /* fraction of sites inside a tandem-repeat interval (0.15) */
    double   ins_noise;        /* per-base substitution noise on inserted sequence (0.03) */
    int32_t  mosaic;           /* 1: 80% of sites at VAF U(0.05,0.20) */
    int32_t  with_seq;         /* 0: seq arena left zero (bases 0 = '=') */
    int32_t  ins_only;         /* 1: every site is an in-CIGAR INS (config-5 stress shape) */
    int32_t  sv_min;           /* planted size range, log-uniform */
    int32_t  sv_max;
    int32_t  threads;          /* 0 = omp default */
    int32_t  _pad;
    const uint8_t* contig_mask; /* NULL = all; otherwise only contigs with mask[c] != 0 get reads and keep records (rank sharding) */
    /* population shapes (config 4): samples share `seed` (the planted sites) and differ in `sample` (their reads; 0 = the single-sample stream);
     * a sample carries each site with probability `site_keep` (0 = every site) */
    uint64_t sample;
    double   site_keep;
} snfb_synth_params;

typedef struct snfb_synth_site {
    int32_t contig, pos, svtype, size, mate_contig, mate_pos, in_tr, hap; /* hap 0 = hom */
    double  vaf;
} snfb_synth_site;

typedef struct snfb_synth_block snfb_synth_block;

/* returns NULL on allocation failure */
snfb_synth_block* snfb_synth_generate(const snfb_synth_params* p);
const snfb_records* snfb_synth_records(const snfb_synth_block* b);
uint64_t snfb_synth_sites(const snfb_synth_block* b, const snfb_synth_site** out);
/* sum of query_alignment_length over records (all, not only those passing filters) */
uint64_t snfb_synth_aligned_bp(const snfb_synth_block* b);
void snfb_synth_free(snfb_synth_block* b);

#ifdef __cplusplus
}
#endif
#endif
