/*
 * snf_oracle.c — CPU restatement of the Sniffles2 lead -> cluster -> consensus path.
 *
 * TEST INFRASTRUCTURE.  This is the checker the CUDA path is compared against
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).  It is
 * never linked into, imported by or called from the product library libsnfb200.so.
 *
 * Parity status: PINNED.  The oracle is checked against the reference itself
 * (oracle/pyref/harness.py imports the unmodified /root/reference/src/sniffles behind a
 * stub pysam) on every synthetic shape, against the committed fixtures in tests/golden/
 * that the reference generated, and against the reference's own BND test vectors
 * (src/tests/test_bnd_leads.py, test_bnd.py).  See tests/test_oracle_golden.py.
 *
 * Plain scalar C, one task (contig) at a time, in the reference's own iteration order.
 * Citations are relative to /root/reference/src/sniffles/.
 * Compile with -ffp-contract=off: decisions use IEEE doubles exactly like CPython.
 */
#include "snf_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ small vectors */
#define VEC(T) struct { T* p; long n, cap; }
#define vpush(v, x) do { if ((v).n == (v).cap) { (v).cap = (v).cap ? (v).cap * 2 : 16; \
    (v).p = realloc((v).p, (size_t)(v).cap * sizeof *(v).p); if (!(v).p) { fprintf(stderr, "oracle: oom\n"); abort(); } } \
    (v).p[(v).n++] = (x); } while (0)

typedef struct { int rec, off, len; } piece_t;

typedef struct {
    snfb_lead L;
    double nm;            /* Lead.nm: read nm for INLINE/SPLIT leads, int(SA nm) for BND (leadprov.py:119) */
    int has_nm;           /* 0: nm is None */
    int ps, has_ps;       /* phase_set; BND leads: None (leadprov.py:109-130 passes no hap/phase_set) */
    long pc_off; int pc_n; /* sequence pieces in T->pieces (valid when SNFB_LF_HAS_SEQ) */
} olead;

typedef VEC(olead) leadvec;
typedef VEC(int) intvec;

typedef struct {
    int svtype, start, end, seed, repeat;
    intvec leads, longs;      /* indices into T->leads */
    int hap[6];
    double mean_svlen, stdev_start;
    int has_long;             /* leads_long is not None */
    int resplit_bin;
} cluster_t;

typedef struct {
    const snfb_records* R; const snfb_config* cfg; int t;
    leadvec leads;            /* emission order, then stably sorted into bin order */
    VEC(piece_t) pieces;
    int32_t* covdiff;         /* difference array of coverage, contig_len+1 */
    uint16_t* cov;
    int32_t* hapref;          /* [3][nbins+1] */
    long nbins;
    uint32_t read_count; double nm_sum; long nm_count; int cur_k;
    /* outputs */
    VEC(snfb_lead) lead_out;  /* snapshot of the leadtab before clustering mutates leads */
    VEC(snfb_cand) cands; leadvec cand_leads; VEC(uint64_t) rnames; VEC(uint32_t) rn_off;
    VEC(uint8_t) alt;
    uint64_t soft_errors;
    double cov_mean;
    int have_prev_end, prev_end;  /* `end` of postprocessing.coverage leaks across iterations (postprocessing.py:86-92) */
} task_t;

struct so_result {
    uint64_t n_leads; snfb_lead* leads;
    uint32_t n_task; uint32_t* task_read_count; double* task_mean_nm; double* rec_nm; double* task_cov_mean;
    uint64_t n_pass, soft_errors;
    uint64_t n_cand; snfb_cand* cand; uint64_t n_cand_leads; snfb_lead* cand_leads;
    uint64_t* rnames; uint32_t* rn_off; uint64_t n_alt; uint8_t* alt;
};

/* ------------------------------------------------------------------ exact statistics.stdev
 * CPython 3.12 statistics.stdev on ints: ss = (n*Sxx - Sx^2)/n exactly, mss = ss/(n-1),
 * result = correctly rounded sqrt of the rational (statistics.py _float_sqrt_of_frac).
 * Here: P = n*Sxx - Sx^2 (shift invariant), Q = n*(n-1); return RN(sqrt(P/Q)). */
static int bitlen128(u128 x) { int b = 0; while (x) { ++b; x >>= 1; } return b; }

static double sqrt_frac_rn(u128 P, uint64_t Q) {
    if (P == 0) return 0.0;
    /* V = floor(P * 2^s / Q) with ~110..112 bits; 256-bit numerator in 32-bit limbs */
    int bl = bitlen128(P) - bitlen128((u128)Q);
    int s = 111 - bl; if (s < 0) s = 0; if (s & 1) ++s;
    uint32_t num[12] = { 0 };
    for (int i = 0; i < 4; ++i) {
        uint64_t limb = (uint64_t)((P >> (32 * i)) & 0xFFFFFFFFu);
        int bitpos = 32 * i + s; int w = bitpos / 32, o = bitpos % 32;
        uint64_t v = limb << o;
        uint64_t c = (uint64_t)num[w] + (v & 0xFFFFFFFFu); num[w] = (uint32_t)c;
        c = (uint64_t)num[w + 1] + (v >> 32) + (c >> 32); num[w + 1] = (uint32_t)c;
        /* no further carry: the limbs are disjoint bit ranges */
    }
    uint32_t quo[12]; u128 rem = 0;
    for (int i = 11; i >= 0; --i) { u128 cur = (rem << 32) | num[i]; quo[i] = (uint32_t)(cur / Q); rem = cur % Q; }
    u128 V = 0; for (int i = 3; i >= 0; --i) V = (V << 32) | quo[i];
    /* V < 2^113 by construction, higher limbs are zero */
    uint64_t a = (uint64_t)sqrtl((long double)V);
    while ((u128)a * a > V) --a;
    while ((u128)(a + 1) * (a + 1) <= V) ++a;
    int sticky = ((u128)a * a != V) || rem != 0;
    a |= (uint64_t)sticky;                      /* round to odd, then one correct rounding */
    return ldexp((double)a, -(s / 2));
}

static double stdev_ints(const long* v, long n) {  /* util.stdev (util.py:25-27) / statistics.stdev */
    if (n < 2) return 0.0;
    long base = v[0]; u128 sxx = 0; __int128 sx = 0;
    for (long i = 0; i < n; ++i) { __int128 d = (__int128)v[i] - base; sx += d; sxx += (u128)(d * d); }
    u128 P = (u128)n * sxx - (u128)(sx * sx);
    return sqrt_frac_rn(P, (uint64_t)n * (uint64_t)(n - 1));
}

static int cmp_long(const void* a, const void* b) { long x = *(const long*)a, y = *(const long*)b; return x < y ? -1 : x > y; }

/* util.center = median_modes (util.py:49-58,167) */
static long center(const long* v, long n) {
    long* s = malloc((size_t)n * sizeof *s); memcpy(s, v, (size_t)n * sizeof *s);
    qsort(s, (size_t)n, sizeof *s, cmp_long);
    long maxc = 0; for (long i = 0; i < n;) { long j = i; while (j < n && s[j] == s[i]) ++j; if (j - i > maxc) maxc = j - i; i = j; }
    long m = 0; for (long i = 0; i < n;) { long j = i; while (j < n && s[j] == s[i]) ++j; if (maxc - (j - i) < 3) s[m++] = s[i]; i = j; }
    long r = s[(long)((double)m / 2)];   /* sorted distinct kept values; median_noavg (util.py:43-46) */
    free(s); return r;
}

/* util.stdev(util.trim(v)) (util.py:82-88) */
static double stdev_trim(const long* v, long n) {
    long* s = malloc((size_t)n * sizeof *s); memcpy(s, v, (size_t)n * sizeof *s);
    qsort(s, (size_t)n, sizeof *s, cmp_long);
    long trim_n = (long)((double)n / 100.0 * 25);
    double r = trim_n > 0 ? stdev_ints(s + trim_n, n - 2 * trim_n) : stdev_ints(s, n);
    free(s); return r;
}

/* ------------------------------------------------------------------ CIGAR text (SA tag) */
/* leadprov.CIGAR_analyze (leadprov.py:144-176); returns 0 ok, -1 malformed */
static int cigar_analyze(const uint8_t* c, int n, long* clip_start, long* clip_end, long* refspan, long* readspan) {
    long rs = 0, qs = 0, clip = 0, cstart = -1; long val = 0; int have = 0;
    for (int i = 0; i < n; ++i) {
        uint8_t ch = c[i];
        if (ch >= '0' && ch <= '9') { val = val * 10 + (ch - '0'); have = 1; if (val > (1L << 40)) return -1; continue; }
        if (!have) return -1;                 /* int("") raises */
        int h = 0;
        if (ch == 'M' || ch == 'I' || ch == 'X' || ch == '=') { qs += val; h = 1; }
        if (ch == 'M' || ch == 'D' || ch == 'X' || ch == '=' || ch == 'N') { rs += val; h = 1; }
        if (!h) {
            if (ch == 'S' || ch == 'H') { if (cstart < 0 && qs + rs > 0) cstart = clip; clip += val; }
            else return -1;                   /* raise "Unknown CIGAR operation" */
        }
        val = 0; have = 0;
    }
    /* trailing digits without an op are silently ignored by the reference loop */
    if (cstart < 0) cstart = clip;
    *clip_start = cstart; *clip_end = clip - cstart; *refspan = rs; *readspan = qs; return 0;
}

typedef struct { const uint8_t* f[6]; int l[6]; } sa_entry;

/* splits one SA entry "rname,pos,strand,CIGAR,mapQ,NM" */
static int sa_fields(const uint8_t* s, int n, sa_entry* e) {
    int k = 0, st = 0;
    for (int i = 0; i <= n; ++i) if (i == n || s[i] == ',') { if (k == 6) return -1; e->f[k] = s + st; e->l[k] = i - st; ++k; st = i + 1; }
    return k == 6 ? 0 : -1;
}
static int parse_int(const uint8_t* s, int n, long* out) {
    if (n <= 0) return -1; long v = 0; int i = 0, neg = 0;
    if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; if (n == 1) return -1; }
    for (; i < n; ++i) { if (s[i] < '0' || s[i] > '9') return -1; v = v * 10 + (s[i] - '0'); if (v > (1L << 40)) return -1; }
    *out = neg ? -v : v; return 0;
}
uint64_t so_hash_name(const uint8_t* s, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull; for (size_t i = 0; i < n; ++i) { h ^= s[i]; h *= 0x100000001b3ull; } return h;
}
static int contig_lookup(const snfb_records* R, const uint8_t* s, int n) {
    uint64_t h = so_hash_name(s, (size_t)n);
    for (uint32_t i = 0; i < R->n_contig; ++i) if (R->contig[i].name_hash == h) return (int)i;
    return -1;
}
/* query-name hash; the device computes the same function (csrc/common.cuh qname_hash) */
uint64_t so_qname_hash(const uint8_t* s, size_t n) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)n;
    for (size_t i = 0; i < n; i += 8) {
        uint64_t w = 0; size_t m = n - i < 8 ? n - i : 8;
        for (size_t j = 0; j < m; ++j) w |= (uint64_t)s[i + j] << (8 * j);
        uint64_t z = w + 0x9E3779B97F4A7C15ull * (uint64_t)(i / 8 + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        h += z;   /* commutative combine so that lanes can hash words independently */
    }
    h = (h ^ (h >> 33)) * 0xff51afd7ed558ccdull; h = (h ^ (h >> 33)) * 0xc4ceb9fe1a85ec53ull; return h ^ (h >> 33);
}

/* ------------------------------------------------------------------ stage A: leads */
static void emit_lead(task_t* T, olead* ld, int contig) {
    const snfb_task* tk = &T->R->task[T->t];
    /* build_leadtab keeps a lead only on the task's contig and inside its region (leadprov.py:464-468) */
    if (contig != tk->contig || ld->L.ref_start < tk->start || ld->L.ref_start >= tk->end) return;
    ld->L.task = (uint16_t)T->t; ld->L.k = (uint16_t)T->cur_k++;   /* ordinal among the read's kept leads */
    vpush(T->leads, *ld);
}
static long add_piece(task_t* T, int rec, int off, int len) { piece_t p = { rec, off, len }; vpush(T->pieces, p); return T->pieces.n - 1; }

typedef struct { int contig; long ref_start, ref_end, qry_start, qry_end; int rev, mapq, source; int nhint; int h_type[2]; long h_start[2], h_len[2]; int h_none[2];
                 int has_seq; long seq_off, seq_len; } seg_t;

/* python slice length of s[a:b] for len L, a,b >= 0 */
static void py_slice(long L, long a, long b, long* off, long* len) { if (a > L) a = L; if (b > L) b = L; *off = a; *len = b > a ? b - a : 0; }

/* sv.classify_splits (sv.py:649-782); segs sorted in place; returns new count */
static int classify_splits(const snfb_config* cfg, seg_t* s, int n, long l_seq) {
    /* leads.sort(key=qry_start): stable insertion sort */
    for (int i = 1; i < n; ++i) { seg_t x = s[i]; int j = i - 1; while (j >= 0 && s[j].qry_start > x.qry_start) { s[j + 1] = s[j]; --j; } s[j + 1] = x; }
    for (int i = 0; i < n; ++i) { s[i].nhint = 0; }
    int hints = 0; long ms = cfg->minsvlen_screen;
    if ((double)s[0].qry_start >= (double)cfg->long_ins_length * 0.5) { s[0].h_type[0] = SNFB_INS; s[0].h_start[0] = s[0].ref_start; s[0].h_none[0] = 1; s[0].h_len[0] = 0; s[0].nhint = 1; }
    for (int i = 1; i < n; ++i) {
        seg_t* cu = &s[i]; seg_t* la = &s[i - 1];
        if (cu->contig != la->contig) continue;
        int rev = cu->rev, fwd = !rev; int ty = -1; long st = 0, ln = 0;
        if (cu->rev == la->rev) {
            long dq = cu->qry_start - la->qry_end;
            if (fwd && dq >= ms && dq - (cu->ref_start - la->ref_end) >= ms) {
                ty = SNFB_INS; st = cu->ref_start; ln = dq;
                if (ln <= cfg->dev_seq_cache_maxlen) { cu->has_seq = 1; py_slice(l_seq, la->qry_end, cu->qry_start, &cu->seq_off, &cu->seq_len); } else cu->has_seq = 0;
            } else if (rev && dq >= ms && dq - (la->ref_start - cu->ref_end) >= ms) {
                ty = SNFB_INS; st = la->ref_start; ln = dq;
                if (ln <= cfg->dev_seq_cache_maxlen) { cu->has_seq = 1; py_slice(l_seq, la->qry_end, cu->qry_start, &cu->seq_off, &cu->seq_len); } else cu->has_seq = 0;
            } else if (fwd && (cu->ref_start - la->ref_end) >= ms && (cu->ref_start - la->ref_end) - dq >= ms) {
                ty = SNFB_DEL; st = cu->ref_start; ln = -(cu->ref_start - la->ref_end);
            } else if (rev && (la->ref_start - cu->ref_end) >= ms && (la->ref_start - cu->ref_end) - dq >= ms) {
                ty = SNFB_DEL; st = la->ref_start; ln = -(la->ref_start - cu->ref_end);
            } else if (fwd && cu->ref_start <= la->ref_end) {
                st = cu->ref_start; ln = la->ref_end - cu->ref_start; if (ln >= ms) ty = SNFB_DUP;
            } else if (rev && la->ref_start <= cu->ref_end) {
                st = la->ref_start; ln = cu->ref_end - la->ref_start; if (ln >= ms) ty = SNFB_DUP;
            }
        } else {
            if (fwd && cu->ref_start <= la->ref_start) { st = cu->ref_start; ln = la->ref_start - cu->ref_start; if (ln >= ms) ty = SNFB_INV; }
            else if (fwd && cu->ref_start > la->ref_start) { st = la->ref_start; ln = cu->ref_start - la->ref_start; if (ln >= ms) ty = SNFB_INV; }
            else if (rev && cu->ref_end >= la->ref_end) { st = la->ref_end; ln = cu->ref_end - la->ref_end; if (ln >= ms) ty = SNFB_INV; }
            else if (rev && cu->ref_end < la->ref_end) { st = cu->ref_end; ln = la->ref_end - cu->ref_end; if (ln >= ms) ty = SNFB_INV; }
        }
        if (ty >= 0) { int k = cu->nhint++; cu->h_type[k] = ty; cu->h_start[k] = st; cu->h_len[k] = ln; cu->h_none[k] = 0; ++hints; }
    }
    if (!hints && n > 2) {
        int m = 0; int c0 = s[0].contig, r0 = s[0].rev;
        for (int i = 0; i < n; ++i) if (s[i].contig == c0 && s[i].rev == r0) s[m++] = s[i];
        if (m == 2) { s[0].has_seq = s[1].has_seq = 0; return classify_splits(cfg, s, 2, l_seq); }
        return m;
    }
    return n;
}

static void process_read(task_t* T, uint64_t ri, double* rec_nm) {
    const snfb_records* R = T->R; const snfb_config* cfg = T->cfg; const snfb_rec* r = &R->rec[ri];
    const snfb_task* tk = &R->task[T->t];
    const uint32_t* cg = (const uint32_t*)R->cigar + r->cigar_off;   /* the oracle reads BAM words (SNFB_CIGAR_BAM32) */ long n = r->n_cigar;
    if (rec_nm) rec_nm[ri] = -1.0;
    /* pysam query_alignment_start / _end */
    long qas = 0, qae = r->l_seq;
    for (long k = 0; k < n; ++k) { int op = cg[k] & 15; if (op == 4) qas += cg[k] >> 4; else if (op != 5) break; }
    for (long k = n - 1; k >= 1; --k) { int op = cg[k] & 15; if (op == 4) qae -= cg[k] >> 4; else if (op != 5) break; }
    long alen = qae - qas;
    /* leadprov.py:494-503 */
    if (r->mapq < cfg->mapq || (r->flag & 256) || alen < cfg->min_alignment_length) return;
    if (cfg->exclude_flags && (r->flag & cfg->exclude_flags)) return;
    if (r->pos < tk->start || r->pos >= tk->end) return;
    int hp = (r->aux_flags & SNFB_AUX_HP) ? r->hp : 0;
    if (hp > 2) { hp = 0; T->soft_errors++; }   /* the reference indexes a 3-slot array (leadprov.py:392) */
    int has_ps = (r->aux_flags & SNFB_AUX_PS) != 0;
    T->read_count++;
    long ref_end = r->pos, ins_l = 0, del_l = 0;
    for (long k = 0; k < n; ++k) {
        int op = cg[k] & 15; long len = cg[k] >> 4;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_end += len;
        if (op == 1 && len > 10) ins_l += len;           /* get_cigar_indels (leadprov.py:198-224) */
        if (op == 2 && len > 10) del_l += len;
    }
    { long a = r->pos, b = ref_end; if (b > tk->contig_len) b = tk->contig_len; if (a < b) { T->covdiff[a] += 1; T->covdiff[b] -= 1; } }  /* leadprov.py:510 */
    int is_supp = (r->flag & 2048) != 0, rev = (r->flag & 16) != 0;
    int has_sa = (r->aux_flags & SNFB_AUX_SA) != 0;
    int use_clips = cfg->detect_large_ins && !is_supp && !has_sa;
    double nm = -1.0;
    if ((cfg->qc_nm_measure || cfg->phase) && (r->aux_flags & SNFB_AUX_NM)) {   /* leadprov.py:517-526 */
        nm = (double)(r->nm - (ins_l + del_l)) / (double)(alen + 1);
        T->nm_sum += nm; T->nm_count++;
    }
    if (rec_nm) rec_nm[ri] = nm;
    uint64_t qh = so_qname_hash(R->var + r->var_off, r->l_qname);
    uint32_t base_flags = (rev ? SNFB_LF_REVERSE : 0) | ((uint32_t)r->mapq << 16);
    T->cur_k = 0;
    /* read_iterindels (leadprov.py:583-670) */
    {
        long pos_read = 0, pos_ref = r->pos;
        double longinslen = (double)cfg->long_ins_length / 2.0;
        for (long k = 0; k < n; ++k) {
            int op = cg[k] & 15; long len = cg[k] >> 4;
            int add_read = (op == 0 || op == 1 || op == 4 || op == 7 || op == 8);
            int add_ref = (op == 0 || op == 2 || op == 3 || op == 7 || op == 8);
            int event = (op == 1 || op == 2 || op == 4);
            if (event && len >= cfg->minsvlen_screen) {
                olead ld; memset(&ld, 0, sizeof ld);
                ld.L.rec = (uint32_t)ri; ld.L.qname_hash = qh; ld.nm = nm; ld.has_nm = 1; ld.ps = r->ps; ld.has_ps = has_ps;
                ld.L.read_len = (int32_t)alen; ld.L.seq_off = -1; ld.L.mate_contig = -1;
                uint32_t f = base_flags | ((uint32_t)SNFB_SRC_INLINE << 3) | ((uint32_t)hp << 24) | (is_supp ? SNFB_LF_IS_SA : 0);
                int emit = 1;
                if (op == 1) {
                    f |= SNFB_INS; ld.L.ref_start = (int32_t)pos_ref; ld.L.ref_end = (int32_t)pos_ref;
                    ld.L.qry_start = (int32_t)pos_read; ld.L.qry_end = (int32_t)(pos_read + len); ld.L.svlen = (int32_t)len;
                    if (len <= cfg->dev_seq_cache_maxlen) { f |= SNFB_LF_HAS_SEQ; ld.L.seq_off = (int32_t)pos_read; ld.L.seq_len = (int32_t)len; }
                } else if (op == 2) {
                    f |= SNFB_DEL; ld.L.ref_start = (int32_t)(pos_ref + len); ld.L.ref_end = (int32_t)pos_ref;   /* sic: leadprov.py:622-626 */
                    ld.L.qry_start = (int32_t)pos_read; ld.L.qry_end = (int32_t)pos_read; ld.L.svlen = (int32_t)-len;
                } else if (use_clips && (double)len >= longinslen) {
                    f |= SNFB_INS | SNFB_LF_SVLEN_NONE; ld.L.ref_start = ld.L.ref_end = (int32_t)pos_ref;
                    ld.L.qry_start = (int32_t)pos_read; ld.L.qry_end = (int32_t)(pos_read + len);
                } else if (op == 4) {
                    f |= (pos_ref == r->pos) ? SNFB_SINGLE_LEFT : SNFB_SINGLE_RIGHT; ld.L.ref_start = ld.L.ref_end = (int32_t)pos_ref;
                    ld.L.qry_start = (int32_t)pos_read; ld.L.qry_end = (int32_t)(pos_read + len); ld.L.svlen = 0;
                } else emit = 0;
                if (emit) {
                    ld.L.flags = f;
                    if (f & SNFB_LF_HAS_SEQ) { ld.pc_off = add_piece(T, (int)ri, ld.L.seq_off, ld.L.seq_len); ld.pc_n = 1; }
                    emit_lead(T, &ld, tk->contig);
                }
            }
            pos_read += add_read * len; pos_ref += add_ref * len;
        }
    }
    if (has_sa) {
        const uint8_t* sa = R->var + r->var_off + r->l_qname; int sl = (int)r->sa_len;
        /* entries = [part for part in SA.split(";") if len(part) > 0] */
        int est[256], eln[256], ne = 0, too_many = 0;
        for (int i = 0, st = 0; i <= sl; ++i) if (i == sl || sa[i] == ';') { if (i > st) { if (ne < 256) { est[ne] = st; eln[ne] = i - st; ++ne; } else too_many = 1; } st = i + 1; }
        int sa_ok = 1; sa_entry e0;
        if (ne > 0 && sa_fields(sa + est[0], eln[0], &e0)) { sa_ok = 0; T->soft_errors++; }
        /* Lead.for_bnd (leadprov.py:57-132): first SA entry only */
        if (ne > 0 && sa_ok) {
            long left = 0, right = 0;
            { int op = cg[0] & 15; if (op == 4 || op == 5) left = cg[0] >> 4; }
            { int op = cg[n - 1] & 15; if (op == 4 || op == 5) right = cg[n - 1] >> 4; }
            long bstart; int is_first;
            if (left > right) { bstart = r->pos + 1; is_first = 0; } else { bstart = ref_end; is_first = 1; }
            int sa_rev = (e0.l[2] == 1 && e0.f[2][0] == '-');
            int same = (e0.l[2] == 1 && ((e0.f[2][0] == '-' && rev) || (e0.f[2][0] == '+' && !rev)));
            long p1, cs, ce, rs, qs, sanm = 0;
            if (!same) {
                if (parse_int(e0.f[1], e0.l[1], &p1)) { T->soft_errors++; }
                else if (cigar_analyze(e0.f[3], e0.l[3], &cs, &ce, &rs, &qs)) { T->soft_errors++; }
                else if ((r->aux_flags & SNFB_AUX_NM) && parse_int(e0.f[5], e0.l[5], &sanm)) { T->soft_errors++; }
                else {
                    long p0 = p1 - 1; int is_reverse = ce > cs; long mate;
                    if (is_reverse) mate = p0 + rs; else mate = is_first ? p0 + 1 : p0 + 2;
                    (void)sa_rev;
                    olead ld; memset(&ld, 0, sizeof ld);
                    ld.L.rec = (uint32_t)ri; ld.L.qname_hash = qh; ld.L.seq_off = -1;
                    ld.L.ref_start = ld.L.ref_end = (int32_t)bstart; ld.L.qry_start = (int32_t)qas; ld.L.qry_end = (int32_t)qae;
                    ld.L.svlen = 0; ld.L.mate_pos = (int32_t)mate; ld.L.mate_contig = contig_lookup(R, e0.f[0], e0.l[0]);
                    if (ld.L.mate_contig < 0) T->soft_errors++;
                    ld.has_nm = (r->aux_flags & SNFB_AUX_NM) != 0; ld.nm = (double)sanm; ld.L.nm_sa = (int32_t)sanm;
                    ld.L.flags = base_flags | SNFB_BND | ((uint32_t)SNFB_SRC_BND_SA << 3) | (is_first ? SNFB_LF_BND_FIRST : 0) | (is_reverse ? SNFB_LF_BND_REVERSE : 0) | (ld.has_nm ? 0 : SNFB_LF_NM_NONE);
                    /* hap "0", phase_set None, is_sa False, read_len 0: dataclass defaults */
                    emit_lead(T, &ld, tk->contig);
                }
            }
        }
        /* read_itersplits (leadprov.py:227-355): primary alignments only (leadprov.py:553) */
        if (!is_supp && sa_ok && !too_many) {
            double lim = (double)cfg->max_splits_base + cfg->max_splits_kb * ((double)r->l_seq / 1000.0);
            if (!((double)ne > lim)) {
                seg_t* segs = calloc((size_t)ne + 1, sizeof *segs); int ok = 1;
                segs[0].contig = tk->contig; segs[0].ref_start = r->pos; segs[0].ref_end = ref_end;
                segs[0].qry_start = rev ? r->l_seq - qae : qas; segs[0].qry_end = segs[0].qry_start + alen;
                segs[0].rev = rev; segs[0].mapq = r->mapq; segs[0].source = SNFB_SRC_SPLIT_PRIM;
                for (int i = 0; i < ne && ok; ++i) {
                    sa_entry e; long p1, cs, ce, rs, qs, mq;
                    if (sa_fields(sa + est[i], eln[i], &e) || parse_int(e.f[4], e.l[4], &mq)) { ok = 0; T->soft_errors++; break; }
                    int srev = (e.l[2] == 1 && e.f[2][0] == '-');
                    if (cigar_analyze(e.f[3], e.l[3], &cs, &ce, &rs, &qs)) { ok = 0; T->soft_errors++; break; }   /* leadprov.py:268-272 */
                    if (parse_int(e.f[1], e.l[1], &p1)) { ok = 0; T->soft_errors++; break; }
                    seg_t* g = &segs[i + 1];
                    g->contig = contig_lookup(R, e.f[0], e.l[0]); if (g->contig < 0) { g->contig = -2 - i; T->soft_errors++; }
                    g->ref_start = p1 - 1; g->ref_end = p1 - 1 + rs; g->qry_start = srev ? ce : cs; g->qry_end = g->qry_start + qs;
                    g->rev = srev; g->mapq = (int)mq; g->source = SNFB_SRC_SPLIT_SUP;
                }
                if (ok) {
                    int m = classify_splits(cfg, segs, ne + 1, r->l_seq);
                    for (int i = 0; i < m; ++i) for (int h = 0; h < segs[i].nhint; ++h) {
                        int pm = segs[i > 0 ? i - 1 : 0].mapq; int mn = segs[i].mapq < pm ? segs[i].mapq : pm;
                        if (!cfg->dev_keep_lowqual_splits && mn < cfg->mapq) continue;
                        olead ld; memset(&ld, 0, sizeof ld);
                        int ty = segs[i].h_type[h];
                        ld.L.rec = (uint32_t)ri; ld.L.qname_hash = qh; ld.nm = nm; ld.has_nm = 1; ld.ps = r->ps; ld.has_ps = has_ps;
                        ld.L.seq_off = -1; ld.L.mate_contig = -1;
                        ld.L.ref_start = (int32_t)segs[i].h_start[h];
                        ld.L.ref_end = (!segs[i].h_none[h] && ty != SNFB_INS) ? (int32_t)(segs[i].h_start[h] + segs[i].h_len[h]) : ld.L.ref_start;
                        ld.L.qry_start = (int32_t)segs[i].qry_start; ld.L.qry_end = (int32_t)segs[i].qry_end;
                        ld.L.svlen = (int32_t)segs[i].h_len[h];
                        uint32_t f = (uint32_t)ty | ((uint32_t)segs[i].source << 3) | (segs[i].rev ? SNFB_LF_REVERSE : 0) | ((uint32_t)segs[i].mapq << 16) | ((uint32_t)hp << 24);
                        if (segs[i].h_none[h]) f |= SNFB_LF_SVLEN_NONE;
                        if (ty == SNFB_INS && segs[i].has_seq) { f |= SNFB_LF_HAS_SEQ; ld.L.seq_off = (int32_t)segs[i].seq_off; ld.L.seq_len = (int32_t)segs[i].seq_len;
                            ld.pc_off = add_piece(T, (int)ri, ld.L.seq_off, ld.L.seq_len); ld.pc_n = 1; }
                        ld.L.flags = f;
                        emit_lead(T, &ld, segs[i].contig);
                    }
                }
                free(segs);
            }
        } else if (!is_supp && too_many) T->soft_errors++;
    }
    /* record_hap_ref (leadprov.py:387-398, 567-571) */
    { long b0 = r->pos / cfg->cluster_binsize, b1 = ref_end / cfg->cluster_binsize;
      if (b1 > T->nbins) b1 = T->nbins;
      if (b0 < b1) { T->hapref[(long)hp * (T->nbins + 1) + b0] += 1; T->hapref[(long)hp * (T->nbins + 1) + b1] -= 1; } }
}

/* ------------------------------------------------------------------ stage B helpers */
static int lead_bin(const snfb_config* cfg, const olead* l) { return (int)((long)l->L.ref_start / cfg->cluster_binsize) * cfg->cluster_binsize; }

static void compute_metrics(task_t* T, cluster_t* c) {   /* Cluster.compute_metrics (cluster.py:48-61) */
    long len = c->leads.n; long n = len < 100 ? len : 100;
    if (n == 0) { c->mean_svlen = 0; c->stdev_start = 0; return; }
    long step = (long)((double)len / (double)n);
    if (n > 1) {
        long sum = 0, m = 0; long* v = malloc((size_t)len * sizeof *v);
        for (long i = 0; i < len; i += step) { const olead* l = &T->leads.p[c->leads.p[i]]; sum += l->L.svlen; v[m++] = l->L.ref_start; }
        c->mean_svlen = (double)sum / (double)n;        /* numerator may hold more than n terms: keep the quirk */
        c->stdev_start = stdev_ints(v, m); free(v);
    } else { c->mean_svlen = (double)T->leads.p[c->leads.p[0]].L.svlen; c->stdev_start = 0; }
}

typedef struct { uint64_t q; long ref_start; long idx; } mi_key;

/* cluster.merge_inner (cluster.py:85-122) */
static void merge_inner(task_t* T, cluster_t* c, int threshold) {
    long n = c->leads.n; if (n == 0) return;
    /* group by qname in first-seen order; stable sort each group by ref_start */
    int* idx = malloc((size_t)n * sizeof *idx); int* grp = malloc((size_t)n * sizeof *grp);
    uint64_t* gq = malloc((size_t)n * sizeof *gq); long ng = 0;
    for (long i = 0; i < n; ++i) { uint64_t q = T->leads.p[c->leads.p[i]].L.qname_hash; long g = 0; for (; g < ng; ++g) if (gq[g] == q) break; if (g == ng) gq[ng++] = q; grp[i] = (int)g; }
    intvec out = { 0 };
    for (long g = 0; g < ng; ++g) {
        long m = 0; for (long i = 0; i < n; ++i) if (grp[i] == g) idx[m++] = c->leads.p[i];
        for (long i = 1; i < m; ++i) { int x = idx[i]; long j = i - 1; while (j >= 0 && T->leads.p[idx[j]].L.ref_start > T->leads.p[x].L.ref_start) { idx[j + 1] = idx[j]; --j; } idx[j + 1] = x; }
        int cur = idx[0]; olead* tm = &T->leads.p[cur];
        long lre = tm->L.ref_end, lqe = tm->L.qry_end, lrs = tm->L.ref_start, lqs = tm->L.qry_start;
        for (long i = 1; i < m; ++i) {
            olead* to = &T->leads.p[idx[i]]; olead* cl = &T->leads.p[cur];
            int merge = (threshold == -1) ||
                (((labs((long)to->L.ref_start - lre) < threshold || labs((long)to->L.ref_start - lrs) < threshold) &&
                  (labs((long)to->L.qry_start - lqe) < threshold || labs((long)to->L.qry_start - lqs) < threshold)) &&
                 ((cl->L.flags & SNFB_LF_REVERSE) == (to->L.flags & SNFB_LF_REVERSE)));
            if (merge) {
                cl->L.svlen += to->L.svlen;
                if (!(to->L.flags & SNFB_LF_HAS_SEQ) || !(cl->L.flags & SNFB_LF_HAS_SEQ)) { cl->L.flags &= ~SNFB_LF_HAS_SEQ; cl->L.seq_len = 0; cl->L.seq_off = -1; }
                else {
                    long off = T->pieces.n;
                    for (int k = 0; k < cl->pc_n; ++k) { piece_t p = T->pieces.p[cl->pc_off + k]; vpush(T->pieces, p); }
                    for (int k = 0; k < to->pc_n; ++k) { piece_t p = T->pieces.p[to->pc_off + k]; vpush(T->pieces, p); }
                    cl = &T->leads.p[cur]; to = &T->leads.p[idx[i]];
                    cl->pc_off = off; cl->pc_n += to->pc_n; cl->L.seq_len += to->L.seq_len;
                }
            } else { vpush(out, cur); cur = idx[i]; }
            lre = to->L.ref_end; lqe = to->L.qry_end; lrs = to->L.ref_start; lqs = to->L.qry_start;
        }
        vpush(out, cur);
    }
    free(c->leads.p); c->leads.p = out.p; c->leads.n = out.n; c->leads.cap = out.cap;
    free(idx); free(grp); free(gq);
}

/* util.mean(v.nm for v in leads) (sv.py:545): builtin sum() over floats is Neumaier-compensated
 * since CPython 3.12 (Python/bltinmodule.c) */
static double lead_nm_mean(task_t* T, const intvec* v) {
    double s = 0.0, c = 0.0;
    for (long i = 0; i < v->n; ++i) { double x = T->leads.p[v->p[i]].nm; double t = s + x;
        if (fabs(s) >= fabs(x)) c += (s - t) + x; else c += (x - t) + s; s = t; }
    if (c != 0.0 && isfinite(c)) s += c;
    return s / (double)v->n;
}

static int cmp_u64(const void* a, const void* b) { uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b; return x < y ? -1 : x > y; }
static long distinct_u64(uint64_t* v, long n) { if (!n) return 0; qsort(v, (size_t)n, sizeof *v, cmp_u64); long m = 1; for (long i = 1; i < n; ++i) if (v[i] != v[m - 1]) v[m++] = v[i]; return m; }

/* decimal-string comparison used by util.most_common ties on str(ps) (util.py:91-98) */
static int cmp_decstr(long a, long b) { char x[32], y[32]; sprintf(x, "%ld", a); sprintf(y, "%ld", b); return strcmp(x, y); }

/* sv.call_from (sv.py:497-598) + get_sa_count (cluster.py:79-82); appends a candidate */
static void call_from(task_t* T, cluster_t* c) {
    const snfb_config* cfg = T->cfg; long n = c->leads.n; if (n == 0) return;
    int svtype = c->svtype;
    long* svl = malloc((size_t)n * sizeof *svl); long* rst = malloc((size_t)n * sizeof *rst);
    for (long i = 0; i < n; ++i) { const olead* l = &T->leads.p[c->leads.p[i]]; svl[i] = l->L.svlen; rst[i] = l->L.ref_start; }
    long svlen = center(svl, n);
    int single = svtype == SNFB_SINGLE_LEFT || svtype == SNFB_SINGLE_RIGHT;
    if (!single && svtype != SNFB_BND && labs(svlen) < cfg->minsvlen_screen) { free(svl); free(rst); return; }
    snfb_cand cd; memset(&cd, 0, sizeof cd);
    /* get_sa_count */
    { long all = n + (c->has_long ? c->longs.n : 0), sa = 0;
      for (long i = 0; i < n; ++i) sa += (T->leads.p[c->leads.p[i]].L.flags & SNFB_LF_IS_SA) != 0;
      if (c->has_long) for (long i = 0; i < c->longs.n; ++i) sa += (T->leads.p[c->longs.p[i]].L.flags & SNFB_LF_IS_SA) != 0;
      cd.sa_count = (int32_t)sa; cd.sa_total = (int32_t)all; }
    uint64_t* qs = malloc((size_t)(n + c->longs.n + 1) * sizeof *qs); long nq = 0;
    for (long i = 0; i < n; ++i) qs[nq++] = T->leads.p[c->leads.p[i]].L.qname_hash;
    long support, support_long = 0;
    if (svtype == SNFB_INS && svlen >= cfg->long_ins_length) {
        uint64_t* ql = malloc((size_t)(c->longs.n + 1) * sizeof *ql); long nl = 0;
        for (long i = 0; i < c->longs.n; ++i) { ql[nl++] = T->leads.p[c->longs.p[i]].L.qname_hash; qs[nq++] = ql[nl - 1]; }
        support_long = distinct_u64(ql, nl); free(ql);
    }
    nq = distinct_u64(qs, nq); support = nq;
    long ref_start = center(rst, n);
    double sd_pos = stdev_trim(rst, n), sd_len = NAN; int precise;
    if (svtype != SNFB_BND) { sd_len = stdev_trim(svl, n); precise = (sd_pos + sd_len) < (double)cfg->precise; }
    else precise = sd_pos < (double)cfg->precise;
    long svstart, svend;                              /* calculate_bounds (sv.py:484-494) */
    if (svtype == SNFB_INS) { svstart = svend = ref_start; }
    else if (svtype == SNFB_DEL) { svstart = ref_start + svlen; svend = ref_start; }
    else { svstart = ref_start; svend = svstart + labs(svlen); }
    long mq = 0, fwd = 0; for (long i = 0; i < n; ++i) { const olead* l = &T->leads.p[c->leads.p[i]]; mq += SNFB_LF_MAPQ(l->L.flags); fwd += !(l->L.flags & SNFB_LF_REVERSE); }
    cd.task = T->t; cd.svtype = svtype; cd.pos = (int32_t)svstart; cd.end = (int32_t)svend; cd.svlen = (int32_t)svlen;
    cd.qual = (int32_t)((double)mq / (double)n); cd.precise = precise; cd.fwd = (int32_t)fwd; cd.rev = (int32_t)(n - fwd);
    cd.stdev_pos = sd_pos; cd.stdev_len = sd_len;
    cd.nm_mean = cfg->qc_nm_measure ? lead_nm_mean(T, &c->leads) : -1.0;
    cd.support_long = (int32_t)support_long; cd.bnd_mate_contig = -1;
    if (svtype == SNFB_DEL) { long k = 0; for (long i = 0; i < n; ++i) k += SNFB_LF_SOURCE(T->leads.p[c->leads.p[i]].L.flags) != SNFB_SRC_INLINE; cd.support_sa = (int32_t)k; }
    if (svtype == SNFB_BND) {                          /* sv.resolve_bnd (sv.py:625-639) */
        /* most_common_top(mate_contig): highest count, ties -> smallest name */
        int best = -2; long bestc = 0; int bestrank = 0;
        for (long i = 0; i < n; ++i) { int mc = T->leads.p[c->leads.p[i]].L.mate_contig; long k = 0; for (long j = 0; j < n; ++j) k += T->leads.p[c->leads.p[j]].L.mate_contig == mc;
            int rk = mc >= 0 ? T->R->contig[mc].lex_rank : 1 << 30;
            if (k > bestc || (k == bestc && rk < bestrank)) { best = mc; bestc = k; bestrank = rk; } }
        intvec sel = { 0 }; for (long i = 0; i < n; ++i) if (T->leads.p[c->leads.p[i]].L.mate_contig == best) vpush(sel, c->leads.p[i]);
        long* mp = malloc((size_t)sel.n * sizeof *mp); long nf = 0, nr = 0; nq = 0;
        for (long i = 0; i < sel.n; ++i) { const olead* l = &T->leads.p[sel.p[i]]; mp[i] = l->L.mate_pos; nf += (l->L.flags & SNFB_LF_BND_FIRST) != 0; nr += (l->L.flags & SNFB_LF_BND_REVERSE) != 0; qs[nq++] = l->L.qname_hash; }
        cd.bnd_mate_contig = best; cd.bnd_mate_pos = (int32_t)center(mp, sel.n);
        cd.bnd_is_first = nf > sel.n - nf;             /* ties -> sorted([False, True])[0] == False */
        cd.bnd_is_reverse = nr > sel.n - nr;
        nq = distinct_u64(qs, nq); support = nq;
        free(mp); free(c->leads.p); c->leads.p = sel.p; c->leads.n = sel.n; c->leads.cap = sel.cap; n = sel.n;
    }
    cd.support = (int32_t)support;
    memcpy(cd.hap_counts, c->hap, sizeof cd.hap_counts);
    cd.cluster_seed = c->seed; cd.resplit_bin = c->resplit_bin;
    /* final leads, strands, inline support, phase aggregates (postprocessing.phase_sv 626-654) */
    cd.lead_off = (int32_t)T->cand_leads.n; cd.lead_n = (int32_t)n;
    { int f = 0, r = 0; uint64_t* qi = malloc((size_t)(n + 1) * sizeof *qi); long ni = 0;
      for (long i = 0; i < n; ++i) { olead l = T->leads.p[c->leads.p[i]]; vpush(T->cand_leads, l); if (l.L.flags & SNFB_LF_REVERSE) r = 1; else f = 1;
          if (SNFB_LF_SOURCE(l.L.flags) == SNFB_SRC_INLINE) qi[ni++] = l.L.qname_hash; }
      cd.n_strands = f + r; cd.support_inline = (int32_t)distinct_u64(qi, ni); free(qi); }
    cd.long_off = (int32_t)T->cand_leads.n; cd.long_n = 0;
    if (c->has_long && svtype != SNFB_BND) { cd.long_n = (int32_t)c->longs.n; for (long i = 0; i < c->longs.n; ++i) { olead l = T->leads.p[c->longs.p[i]]; vpush(T->cand_leads, l); } }
    {   /* reads_phases = {read_id: (hap, phase_set)}: last lead of a record wins */
        long hc[3] = { 0, 0, 0 }; long* psv = malloc((size_t)n * sizeof *psv); int* psn = malloc((size_t)n * sizeof *psn); long* psc = calloc((size_t)n, sizeof *psc); long np = 0;
        for (long i = 0; i < n; ++i) {
            const olead* l = &T->leads.p[c->leads.p[i]]; int last = 1;
            for (long j = i + 1; j < n; ++j) if (T->leads.p[c->leads.p[j]].L.rec == l->L.rec) { last = 0; break; }
            if (!last) continue;
            hc[SNFB_LF_HAP(l->L.flags)]++;
            int isnull = !l->has_ps; long v = isnull ? 0 : l->ps; long k = 0;
            for (; k < np; ++k) if (psn[k] == isnull && (isnull || psv[k] == v)) break;
            if (k == np) { psv[np] = v; psn[np] = isnull; ++np; } psc[k]++;
        }
        /* hp_list[0]: highest (count, str(hap)); ties go to the larger hap string */
        int ht = 0; for (int h = 1; h < 3; ++h) if (hc[h] > 0 && hc[h] >= hc[ht]) ht = h;
        cd.hp_top = ht; cd.hp_support = (int32_t)hc[ht]; cd.hp_other = (int32_t)(hc[0] + hc[1] + hc[2] - hc[ht]);
        long bt = 0; for (long k = 1; k < np; ++k) {
            int gt;
            if (psc[k] != psc[bt]) gt = psc[k] > psc[bt];
            else if (psn[k] != psn[bt]) gt = psn[k];                 /* "NULL" > any digit string */
            else gt = cmp_decstr(psv[k], psv[bt]) > 0;
            if (gt) bt = k; }
        cd.ps_top = (int32_t)psv[bt]; cd.ps_top_null = psn[bt]; cd.ps_support = (int32_t)psc[bt];
        long oth = 0; for (long k = 0; k < np; ++k) if (k != bt && !psn[k]) oth += psc[k];
        cd.ps_other = (int32_t)oth;
        free(psv); free(psn); free(psc);
    }
    cd.alt_off = -1; cd.alt_len = 0;
    vpush(T->rn_off, (uint32_t)T->rnames.n);
    for (long i = 0; i < nq; ++i) vpush(T->rnames, qs[i]);
    vpush(T->cands, cd);
    free(qs); free(svl); free(rst);
}

/* cluster.resplit (cluster.py:125-161) */
static void resplit_and_call(task_t* T, cluster_t* c) {
    const snfb_config* cfg = T->cfg; long n = c->leads.n;
    long* bins = malloc((size_t)n * sizeof *bins); long nb = 0;
    intvec* bl = calloc((size_t)n, sizeof *bl);
    for (long i = 0; i < n; ++i) {
        long sv = T->leads.p[c->leads.p[i]].L.svlen; long b = (long)((double)labs(sv) / (double)cfg->cluster_resplit_binsize) * cfg->cluster_resplit_binsize;
        long k = 0; for (; k < nb; ++k) if (bins[k] == b) break; if (k == nb) bins[nb++] = b;
        vpush(bl[k], c->leads.p[i]);
    }
    /* new_clusters = sorted(keys) holding indices into bins[]/bl[] */
    long* nc = malloc((size_t)nb * sizeof *nc); for (long k = 0; k < nb; ++k) nc[k] = k;
    for (long i = 1; i < nb; ++i) { long x = nc[i]; long j = i - 1; while (j >= 0 && bins[nc[j]] > bins[x]) { nc[j + 1] = nc[j]; --j; } nc[j + 1] = x; }
    long len = nb, i = 1;
    while (len > 1 && i < len) {
        long li = i - 1 < 0 ? len - 1 : i - 1;             /* python negative index: new_clusters[-1] */
        long last = bins[nc[li]], curr = bins[nc[i]];
        double thr = (double)(curr < last ? curr : last) * cfg->cluster_merge_len; if ((double)cfg->minsvlen > thr) thr = (double)cfg->minsvlen;
        if ((double)labs(curr - last) <= thr) {
            intvec* dst = &bl[nc[i]]; intvec* src = &bl[nc[li]];
            for (long k = 0; k < src->n; ++k) vpush(*dst, src->p[k]);
            for (long k = li; k + 1 < len; ++k) nc[k] = nc[k + 1]; --len;   /* pop(i-1) */
            i = i - 2 > 0 ? i - 2 : 0;
        } else ++i;
    }
    for (long k = 0; k < len; ++k) {
        cluster_t nw = *c; nw.leads = bl[nc[k]]; nw.resplit_bin = (int)bins[nc[k]];
        intvec copy = { 0 }; for (long q = 0; q < nw.leads.n; ++q) vpush(copy, nw.leads.p[q]); nw.leads = copy;
        call_from(T, &nw); free(nw.leads.p);
    }
    for (long k = 0; k < nb; ++k) free(bl[k].p);
    free(bl); free(bins); free(nc);
}

/* cluster.resplit_bnd (cluster.py:164-216) */
static void resplit_bnd_and_call(task_t* T, cluster_t* c, int thr) {
    long n = c->leads.n;
    if (n <= 1) { cluster_t nw = *c; intvec copy = { 0 }; for (long q = 0; q < n; ++q) vpush(copy, c->leads.p[q]); nw.leads = copy; call_from(T, &nw); free(nw.leads.p); return; }
    int* gid = malloc((size_t)n * sizeof *gid); int gmc[n > 0 ? n : 1], gfi[n > 0 ? n : 1]; long ng = 0;
    for (long i = 0; i < n; ++i) { const olead* l = &T->leads.p[c->leads.p[i]]; int mc = l->L.mate_contig, fi = (l->L.flags & SNFB_LF_BND_FIRST) != 0; long g = 0; for (; g < ng; ++g) if (gmc[g] == mc && gfi[g] == fi) break; if (g == ng) { gmc[ng] = mc; gfi[ng] = fi; ++ng; } gid[i] = (int)g; }
    for (long g = 0; g < ng; ++g) {
        /* distinct position bins of the group, ascending; leads keep their order inside a bin */
        long m = 0; long* pb = malloc((size_t)n * sizeof *pb); int* li = malloc((size_t)n * sizeof *li);
        for (long i = 0; i < n; ++i) if (gid[i] == g) { long mp = T->leads.p[c->leads.p[i]].L.mate_pos; pb[m] = thr > 0 ? (mp / thr) * thr : 0; li[m] = c->leads.p[i]; ++m; }
        /* stable sort by bin */
        for (long i = 1; i < m; ++i) { long x = pb[i]; int y = li[i]; long j = i - 1; while (j >= 0 && pb[j] > x) { pb[j + 1] = pb[j]; li[j + 1] = li[j]; --j; } pb[j + 1] = x; li[j + 1] = y; }
        intvec cur = { 0 }; long lastb = pb[0];
        for (long i = 0; i < m; ++i) {
            if (pb[i] != lastb && pb[i] - lastb > thr) {
                cluster_t nw = *c; nw.leads = cur; nw.has_long = 0; nw.longs.n = 0; call_from(T, &nw); free(nw.leads.p); cur.p = NULL; cur.n = cur.cap = 0;
            }
            vpush(cur, li[i]); lastb = pb[i];
        }
        if (cur.n) { cluster_t nw = *c; nw.leads = cur; nw.has_long = 0; nw.longs.n = 0; call_from(T, &nw); free(nw.leads.p); }
        free(pb); free(li);
    }
    free(gid);
}

/* cluster.resolve (cluster.py:219-353) for one svtype over T->leads[lo,hi) (bin-sorted) */
static void resolve(task_t* T, int svtype, long lo, long hi) {
    const snfb_config* cfg = T->cfg; const snfb_task* tk = &T->R->task[T->t];
    if (lo >= hi) return;
    const int32_t* tr = T->R->tr ? T->R->tr + 2 * (long)tk->tr_off : NULL; long ntr = tk->tr_n; long tr_index = 0; long tr_start = 0, tr_end = 0;
    int use_tr = tr != NULL && ntr > 0; if (use_tr) { tr_start = tr[0]; tr_end = tr[1]; }
    VEC(cluster_t) cl = { 0 };
    for (long i = lo; i < hi;) {
        int seed = lead_bin(cfg, &T->leads.p[i]); long j = i; while (j < hi && lead_bin(cfg, &T->leads.p[j]) == seed) ++j;
        int within = 0;
        if (use_tr && tr_index < ntr) {
            while (tr_end < seed && tr_index + 1 < ntr) { ++tr_index; tr_start = tr[2 * tr_index]; tr_end = tr[2 * tr_index + 1]; }
            if (tr_start < seed && seed < tr_end) within = 1;
        }
        cluster_t c; memset(&c, 0, sizeof c);
        c.svtype = svtype; c.start = seed; c.end = seed + cfg->cluster_binsize; c.seed = seed; c.repeat = within || cfg->repeat; c.has_long = svtype == SNFB_INS; c.resplit_bin = -1;
        long hc[3] = { 0, 0, 0 };
        for (long k = i; k < j; ++k) {
            const olead* l = &T->leads.p[k]; hc[SNFB_LF_HAP(l->L.flags)]++;
            if (svtype == SNFB_INS && (l->L.flags & SNFB_LF_SVLEN_NONE)) vpush(c.longs, (int)k); else vpush(c.leads, (int)k);
        }
        for (int h = 0; h < 3; ++h) { c.hap[h] = (int)(hc[h] > 65535 ? 65535 : hc[h]); long b = seed / cfg->cluster_binsize; long v = b <= T->nbins ? T->hapref[(long)h * (T->nbins + 1) + b] : 0; c.hap[3 + h] = (int)(v > 65535 ? 65535 : v); }
        if (c.leads.n >= cfg->dev_min_leads_cluster) { compute_metrics(T, &c); vpush(cl, c); } else { free(c.leads.p); free(c.longs.p); }
        i = j;
    }
    /* merge loop (cluster.py:278-308) */
    long i = 0;
    while (i < cl.n - 1) {
        cluster_t* cu = &cl.p[i]; cluster_t* nx = &cl.p[i + 1];
        long inner = (long)nx->start - cu->end, outer = (long)nx->end - cu->start;
        double msd = cu->stdev_start < nx->stdev_start ? cu->stdev_start : nx->stdev_start;
        int merge = (double)inner <= msd * cfg->cluster_r;
        if (!merge && (cfg->repeat || cu->repeat || nx->repeat)) {
            double lim = (fabs(cu->mean_svlen) + fabs(nx->mean_svlen)) * cfg->cluster_repeat_h; if (cfg->cluster_repeat_h_max < lim) lim = cfg->cluster_repeat_h_max;
            merge = (double)outer <= lim;
        }
        if (!merge) merge = svtype == SNFB_BND && inner <= cfg->cluster_merge_bnd;
        if (merge) {
            for (long k = 0; k < nx->leads.n; ++k) vpush(cu->leads, nx->leads.p[k]);
            if (svtype == SNFB_INS) for (long k = 0; k < nx->longs.n; ++k) vpush(cu->longs, nx->longs.p[k]);
            cu->end = nx->end; cu->repeat = cu->repeat || nx->repeat;
            free(nx->leads.p); free(nx->longs.p);
            for (long k = i + 1; k + 1 < cl.n; ++k) cl.p[k] = cl.p[k + 1]; --cl.n;
            compute_metrics(T, &cl.p[i]);
            i = i - 2 > 0 ? i - 2 : 0;
        }
        ++i;
    }
    for (long q = 0; q < cl.n; ++q) {
        cluster_t* c = &cl.p[q];
        if (c->leads.n == 0) continue;
        if (svtype == SNFB_BND) {
            if (cfg->dev_no_resplit) { c->has_long = 0; call_from(T, c); } else resplit_bnd_and_call(T, c, cfg->cluster_merge_bnd);
        } else {
            if (svtype == SNFB_INS || svtype == SNFB_DEL) merge_inner(T, c, c->repeat ? -1 : cfg->cluster_merge_pos);
            if (!cfg->dev_no_resplit_repeat && !cfg->dev_no_resplit) resplit_and_call(T, c); else call_from(T, c);
        }
    }
    for (long q = 0; q < cl.n; ++q) { free(cl.p[q].leads.p); free(cl.p[q].longs.p); }
    free(cl.p);
}

/* ------------------------------------------------------------------ coverage probes (postprocessing.py:69-130) */
static int cov_at(const task_t* T, long idx, int32_t* out) {
    long L = T->R->task[T->t].contig_len;
    if (idx < 0) idx += L;                 /* numpy negative index */
    if (idx < 0 || idx >= L) return 0;     /* IndexError: field keeps its default */
    *out = T->cov[idx]; return 1;
}
static void coverage_probes(task_t* T) {
    const snfb_config* cfg = T->cfg; long bs = cfg->coverage_binsize, ud = (long)cfg->coverage_binsize * cfg->coverage_updown_bins;
    for (long i = 0; i < T->cands.n; ++i) {
        snfb_cand* c = &T->cands.p[i]; long start = c->pos, end;
        if (c->svtype == SNFB_INS) end = start + 1;
        else if (c->svtype == SNFB_BND) { if (c->bnd_is_first) start -= 1; if (!T->have_prev_end) { T->soft_errors++; end = start; } else end = T->prev_end; }
        else end = (long)c->pos + labs((long)c->svlen);
        T->prev_end = (int)end; T->have_prev_end = 1;
        if (c->svtype == SNFB_INS || c->svtype == SNFB_BND) { cov_at(T, start - bs, &c->cov_start); cov_at(T, start, &c->cov_center); cov_at(T, end + bs, &c->cov_end); }
        else { cov_at(T, start, &c->cov_start); cov_at(T, (long)((double)(start + end) / 2), &c->cov_center); cov_at(T, end - bs, &c->cov_end); }
        cov_at(T, start - ud, &c->cov_upstream); cov_at(T, end + ud, &c->cov_downstream);
    }
}

/* ------------------------------------------------------------------ stage C: INS consensus */
static const char SEQ_CODE[17] = "=ACMGRSVTWYHKDBN";
static char* lead_seq(task_t* T, const olead* l, long* len) {
    char* s = malloc((size_t)l->L.seq_len + 1); long m = 0;
    for (int k = 0; k < l->pc_n; ++k) { piece_t p = T->pieces.p[l->pc_off + k]; const snfb_rec* r = &T->R->rec[p.rec]; const uint8_t* sq = T->R->seq + r->seq_off;
        for (int j = 0; j < p.len; ++j) { long q = (long)p.off + j; uint8_t b = sq[q >> 1]; s[m++] = SEQ_CODE[(q & 1) ? (b & 15) : (b >> 4)]; } }
    s[m] = 0; *len = m; return s;
}

typedef struct { uint64_t key; long pos; int state; } kslot;   /* state 0 empty, 1 anchor, 2 taboo */
static uint64_t kmer_key(const char* s, int k) { uint64_t v = 0; for (int i = 0; i < k; ++i) v = (v << 8) | (uint8_t)s[i]; return v; }

/* consensus.novel_from_reads (consensus.py:280-394) */
static char* novel_from_reads(const char* best, long L, char** oth, const long* olen, long no, int klen, long skip) {
    long cap = 64; while (cap < 4 * (L / (skip > 0 ? skip : 1) + 2)) cap *= 2;
    kslot* tab = calloc((size_t)cap, sizeof *tab);
#define SLOT(key) ({ uint64_t h_ = (key) * 0x9E3779B97F4A7C15ull; long p_ = (long)(h_ >> 20) & (cap - 1); while (tab[p_].state && tab[p_].key != (key)) p_ = (p_ + 1) & (cap - 1); p_; })
    for (long i = 0; i < L - klen; i += skip) { uint64_t key = kmer_key(best + i, klen); long p = SLOT(key);
        if (tab[p].state == 2) continue; if (tab[p].state == 1) { tab[p].state = 2; continue; } tab[p].key = key; tab[p].pos = i; tab[p].state = 1; }
    char** aln = malloc((size_t)(no + 1) * sizeof *aln); long na = 0;
    for (long li = 0; li < no; ++li) {
        const char* rd = oth[li]; long Lo = olen[li];
        char* cs = malloc((size_t)L + 1); long cl = 0; long last_i = -1, last_j = -1; int have = 0; long span = 0;
        for (long j = 0; j < Lo - klen; j += skip) {
            uint64_t key = kmer_key(rd + j, klen); long p = SLOT(key);
            if (tab[p].state != 1) continue;
            long i = tab[p].pos;
            if (labs(i - j) > klen) continue;
            if (have && i <= last_i) continue;
            if (!have) { if (j > 0) { for (long q = 0; q < i; ++q) cs[cl++] = '-'; } }
            else {
                long fwd_i = i - last_i, fwd_j = j - last_j;
                if (cl + fwd_j > L) fwd_j = L - cl;
                if (fwd_i == fwd_j && fwd_j > 0) {
                    long d = j - last_j; span += d; long m = 0;
                    for (long l = 1; l <= d; ++l) if (last_i + l < L && rd[last_j + l] == best[last_i + l]) ++m;
                    if ((double)m / (double)d >= 0.5) { for (long q = 0; q < fwd_j; ++q) cs[cl++] = rd[last_j + q]; }
                    else for (long q = 0; q < fwd_j; ++q) cs[cl++] = '-';
                } else for (long q = 0; q < fwd_j; ++q) cs[cl++] = '-';
            }
            last_i = i; last_j = j; have = 1;
        }
        while (cl < L) cs[cl++] = '-';
        for (long h = 0; h < L;) {
            if (cs[h] == '-') { ++h; continue; }
            long b = h, ident = 0; while (h < L && cs[h] != '-') { ident += best[h] == cs[h]; ++h; }
            if (!((double)ident / (double)(h - b) > 0.5 && ident > 5)) for (long q = b; q < h; ++q) cs[q] = '-';
        }
        if ((double)span / (double)L > 0.2) aln[na++] = cs; else free(cs);
    }
    double maxal = L > 0 ? (double)(1 + na) : 1.0;
    char* out = malloc((size_t)L + 1);
    for (long i = 0; i < L; ++i) {
        long cnt[256] = { 0 }; long nal = 0;
        for (long a = 0; a < na; ++a) if (aln[a][i] != '-') { cnt[(uint8_t)aln[a][i]]++; ++nal; }
        if (nal < 2 || (double)nal / maxal < 0.25) { out[i] = best[i]; continue; }
        cnt[(uint8_t)best[i]]++;
        long t0 = -1, t1 = -1; int c0 = 0, nd = 0;
        for (int ch = 255; ch >= 0; --ch) if (cnt[ch]) { ++nd; if (cnt[ch] > t0) { t1 = t0; t0 = cnt[ch]; c0 = ch; } else if (cnt[ch] > t1) t1 = cnt[ch]; }
        out[i] = (nd > 1 && t0 - t1 >= 3) ? (char)c0 : best[i];
    }
    out[L] = 0;
    for (long a = 0; a < na; ++a) free(aln[a]); free(aln); free(tab);
    return out;
#undef SLOT
}

/* postprocessing.annotate_sv INS branch (postprocessing.py:33-66) */
static void annotate_ins(task_t* T) {
    const snfb_config* cfg = T->cfg; if (cfg->symbolic) return;
    for (long ci = 0; ci < T->cands.n; ++ci) {
        snfb_cand* c = &T->cands.p[ci]; if (c->svtype != SNFB_INS) continue;
        long nm = 0; long* ml = malloc((size_t)(c->lead_n + 1) * sizeof *ml);
        for (long i = 0; i < c->lead_n; ++i) if (T->cand_leads.p[c->lead_off + i].L.flags & SNFB_LF_HAS_SEQ) ml[nm++] = c->lead_off + i;
        if (!nm) { free(ml); continue; }
        long bi = 0; double bd = 0;
        for (long i = 0; i < nm; ++i) { const olead* l = &T->cand_leads.p[ml[i]];
            double d = (double)labs((long)l->L.seq_len - c->svlen) + (double)labs((long)l->L.ref_start - c->pos) * 1.5;
            if (i == 0 || d < bd) { bd = d; bi = i; } }
        long bl; char* best = lead_seq(T, &T->cand_leads.p[ml[bi]], &bl); char* alt = best;
        if (nm - 1 >= cfg->consensus_min_reads && !cfg->no_consensus) {
            char** oth = malloc((size_t)nm * sizeof *oth); long* ol = malloc((size_t)nm * sizeof *ol); long no = 0;
            for (long i = 0; i < nm; ++i) if (i != bi) { oth[no] = lead_seq(T, &T->cand_leads.p[ml[i]], &ol[no]); ++no; }
            long skip = cfg->consensus_kmer_skip_base + (long)((double)bl * cfg->consensus_kmer_skip_seqlen_mult);
            alt = novel_from_reads(best, bl, oth, ol, no, cfg->consensus_kmer_len, skip);
            for (long i = 0; i < no; ++i) free(oth[i]); free(oth); free(ol); free(best);
        }
        c->alt_off = (int32_t)T->alt.n; c->alt_len = (int32_t)bl;
        for (long i = 0; i < bl; ++i) vpush(T->alt, (uint8_t)alt[i]);
        free(alt); free(ml);
    }
}

/* ------------------------------------------------------------------ driver */
static void run_task(task_t* T, double* rec_nm, int stages, uint64_t rec_lo, uint64_t rec_hi) {
    const snfb_records* R = T->R; const snfb_config* cfg = T->cfg; const snfb_task* tk = &R->task[T->t];
    long L = tk->contig_len; T->nbins = L / cfg->cluster_binsize + 1;
    T->covdiff = calloc((size_t)L + 2, sizeof *T->covdiff);
    T->hapref = calloc((size_t)3 * (size_t)(T->nbins + 1), sizeof *T->hapref);
    for (uint64_t i = rec_lo; i < rec_hi; ++i) if (R->rec[i].task == T->t) process_read(T, i, rec_nm);
    /* leadtab: stable order (svtype, bin, emission) — merge sort via index keys */
    {
        long n = T->leads.n; olead* tmp = malloc((size_t)(n + 1) * sizeof *tmp); long* key = malloc((size_t)(n + 1) * sizeof *key); long* ord = malloc((size_t)(n + 1) * sizeof *ord);
        for (long i = 0; i < n; ++i) { key[i] = ((long)SNFB_LF_TYPE(T->leads.p[i].L.flags) << 40) | (long)(T->leads.p[i].L.ref_start / cfg->cluster_binsize); ord[i] = i; }
        /* bottom-up stable merge sort of ord by key */
        long* buf = malloc((size_t)(n + 1) * sizeof *buf);
        for (long w = 1; w < n; w *= 2) { for (long lo = 0; lo < n; lo += 2 * w) { long mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n; long a = lo, b = mid, o = lo;
                while (a < mid && b < hi) buf[o++] = key[ord[b]] < key[ord[a]] ? ord[b++] : ord[a++]; while (a < mid) buf[o++] = ord[a++]; while (b < hi) buf[o++] = ord[b++]; }
            long* t = ord; ord = buf; buf = t; }
        for (long i = 0; i < n; ++i) tmp[i] = T->leads.p[ord[i]];
        memcpy(T->leads.p, tmp, (size_t)n * sizeof *tmp); free(tmp); free(key); free(ord); free(buf);
        /* record_lead: beyond consensus_max_reads_bin leads in a bin, seq = None (leadprov.py:406-408) */
        for (long i = 0; i < n;) { long j = i; int ty = SNFB_LF_TYPE(T->leads.p[i].L.flags), b = lead_bin(cfg, &T->leads.p[i]);
            while (j < n && SNFB_LF_TYPE(T->leads.p[j].L.flags) == ty && lead_bin(cfg, &T->leads.p[j]) == b) { if (j - i >= cfg->consensus_max_reads_bin) { T->leads.p[j].L.flags &= ~SNFB_LF_HAS_SEQ; T->leads.p[j].L.seq_off = -1; T->leads.p[j].L.seq_len = 0; } ++j; }
            i = j; }
    }
    for (long i = 0; i < T->leads.n; ++i) vpush(T->lead_out, T->leads.p[i].L);
    /* coverage and hap-REF prefix sums */
    T->cov = malloc((size_t)(L + 1) * sizeof *T->cov);
    { int32_t acc = 0; for (long i = 0; i < L; ++i) { acc += T->covdiff[i]; T->cov[i] = (uint16_t)acc; } }
    if (R->n_mask && R->mask && R->mask_task_off)        /* _mask_N_coverage: coverage[mask == 'N'] = 0 (leadprov.py:420-443) */
        for (uint32_t m = R->mask_task_off[T->t]; m < R->mask_task_off[T->t + 1]; ++m) { long a = R->mask[2 * m], b = R->mask[2 * m + 1]; if (a < tk->start) a = tk->start; if (b > tk->end) b = tk->end;   /* the mask is fetched per region (leadprov.py:436-438) */
            if (a < 0) a = 0; if (b > L) b = L; for (long i = a; i < b; ++i) T->cov[i] = 0; }
    { double tot = 0; for (long i = 0; i < L; ++i) tot += (double)T->cov[i]; T->cov_mean = L > 0 ? tot / (double)L : 0.0; }
    for (int h = 0; h < 3; ++h) { int32_t acc = 0; for (long b = 0; b <= T->nbins; ++b) { acc += T->hapref[(long)h * (T->nbins + 1) + b]; T->hapref[(long)h * (T->nbins + 1) + b] = acc; } }
    if (stages >= 2) {
        long n = T->leads.n, i = 0;
        for (int ty = 0; ty < SNFB_NTYPES; ++ty) { long j = i; while (j < n && (int)SNFB_LF_TYPE(T->leads.p[j].L.flags) == ty) ++j; resolve(T, ty, i, j); i = j; }
        vpush(T->rn_off, (uint32_t)T->rnames.n);
        coverage_probes(T);
        if (stages >= 3) annotate_ins(T);
    }
    free(T->covdiff); free(T->cov); free(T->hapref); T->covdiff = NULL; T->cov = NULL; T->hapref = NULL;
}

so_result* so_run(const snfb_records* R, const snfb_config* cfg, int stages, int threads) {
    so_result* out = calloc(1, sizeof *out); uint32_t nt = R->n_task;
    task_t* T = calloc(nt ? nt : 1, sizeof *T);
    out->n_task = nt; out->task_read_count = calloc(nt + 1, sizeof(uint32_t)); out->task_mean_nm = calloc(nt + 1, sizeof(double));
    out->task_cov_mean = calloc(nt + 1, sizeof(double)); out->rec_nm = malloc((R->n_rec + 1) * sizeof(double));
    /* records are sorted by task: find each task's range once */
    uint64_t* lo = calloc(nt + 1, sizeof *lo); uint64_t* hi = calloc(nt + 1, sizeof *hi);
    for (uint32_t t = 0; t < nt; ++t) { lo[t] = R->n_rec; hi[t] = 0; }
    for (uint64_t i = 0; i < R->n_rec; ++i) { int t = R->rec[i].task; if (t < 0 || (uint32_t)t >= nt) continue; if (i < lo[t]) lo[t] = i; if (i + 1 > hi[t]) hi[t] = i + 1; }
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
    for (int t = 0; t < (int)nt; ++t) {
        T[t].R = R; T[t].cfg = cfg; T[t].t = t;
        if (!(lo[t] < hi[t])) { vpush(T[t].rn_off, 0u); continue; }   /* no records: nothing to do for this task */
        run_task(&T[t], out->rec_nm, stages, lo[t] < hi[t] ? lo[t] : 0, lo[t] < hi[t] ? hi[t] : 0);
    }
    uint64_t nl = 0, ncd = 0, ncl = 0, nrn = 0, nalt = 0;
    for (uint32_t t = 0; t < nt; ++t) { nl += (uint64_t)T[t].lead_out.n; ncd += (uint64_t)T[t].cands.n; ncl += (uint64_t)T[t].cand_leads.n; nrn += (uint64_t)T[t].rnames.n; nalt += (uint64_t)T[t].alt.n; }
    out->leads = malloc((nl + 1) * sizeof(snfb_lead)); out->cand = malloc((ncd + 1) * sizeof(snfb_cand)); out->cand_leads = malloc((ncl + 1) * sizeof(snfb_lead));
    out->rnames = malloc((nrn + 1) * 8); out->rn_off = malloc((ncd + 2) * 4); out->alt = malloc(nalt + 1);
    uint64_t ol = 0, oc = 0, ocl = 0, orn = 0, oalt = 0;
    for (uint32_t t = 0; t < nt; ++t) {
        for (long i = 0; i < T[t].lead_out.n; ++i) out->leads[ol++] = T[t].lead_out.p[i];
        for (long i = 0; i < T[t].cands.n; ++i) { snfb_cand c = T[t].cands.p[i]; out->rn_off[oc] = (uint32_t)(orn + T[t].rn_off.p[i]); c.lead_off += (int32_t)ocl; c.long_off += (int32_t)ocl; if (c.alt_off >= 0) c.alt_off += (int32_t)oalt; out->cand[oc++] = c; }
        for (long i = 0; i < T[t].cand_leads.n; ++i) out->cand_leads[ocl++] = T[t].cand_leads.p[i].L;
        memcpy(out->rnames + orn, T[t].rnames.p, (size_t)T[t].rnames.n * 8); orn += (uint64_t)T[t].rnames.n;
        memcpy(out->alt + oalt, T[t].alt.p, (size_t)T[t].alt.n); oalt += (uint64_t)T[t].alt.n;
        out->task_read_count[t] = T[t].read_count; out->n_pass += T[t].read_count; out->soft_errors += T[t].soft_errors;
        out->task_mean_nm[t] = T[t].nm_sum / (double)(T[t].nm_count > 1 ? T[t].nm_count : 1);   /* leadprov.py:577 */
        out->task_cov_mean[t] = T[t].cov_mean;
        free(T[t].leads.p); free(T[t].lead_out.p); free(T[t].pieces.p); free(T[t].cands.p); free(T[t].cand_leads.p); free(T[t].rnames.p); free(T[t].rn_off.p); free(T[t].alt.p);
    }
    out->rn_off[oc] = (uint32_t)orn;
    out->n_leads = nl; out->n_cand = ncd; out->n_cand_leads = ncl; out->n_alt = nalt;
    free(T); free(lo); free(hi);
    return out;
}

void so_get(const so_result* r, so_view* v) {
    v->n_leads = r->n_leads; v->leads = r->leads; v->n_task = r->n_task; v->task_read_count = r->task_read_count; v->task_mean_nm = r->task_mean_nm;
    v->rec_nm = r->rec_nm; v->task_cov_mean = r->task_cov_mean; v->n_pass = r->n_pass; v->soft_errors = r->soft_errors;
    v->n_cand = r->n_cand; v->cand = r->cand; v->n_cand_leads = r->n_cand_leads; v->cand_leads = r->cand_leads; v->rnames = r->rnames; v->rn_off = r->rn_off;
    v->n_alt = r->n_alt; v->alt = r->alt;
}
void so_free(so_result* r) {
    if (!r) return;
    free(r->leads); free(r->task_read_count); free(r->task_mean_nm); free(r->rec_nm); free(r->task_cov_mean); free(r->cand); free(r->cand_leads);
    free(r->rnames); free(r->rn_off); free(r->alt); free(r);
}
double so_sqrt_frac(uint64_t p_hi, uint64_t p_lo, uint64_t q) { return sqrt_frac_rn(((u128)p_hi << 64) | p_lo, q); }
/* unit-level entry points for the known-answer vectors of SURVEY.md Appendix A (tests/test_known_answers.py) */
long so_center(const long* v, long n) { return center(v, n); }
double so_stdev(const long* v, long n) { return stdev_ints(v, n); }
double so_stdev_trim(const long* v, long n) { return stdev_trim(v, n); }
int so_cigar_analyze(const uint8_t* c, int n, long out[4]) { return cigar_analyze(c, n, &out[0], &out[1], &out[2], &out[3]); }
