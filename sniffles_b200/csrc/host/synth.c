/*
 * synth.c — seeded synthetic alignment-record generator (see synth.h).
 * Host-only test/bench input generator; not part of the product path.
 */
#include "synth.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------- rng ---------- */
typedef struct { uint64_t s; } rng_t;
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t rnext(rng_t* r) { r->s += 0x9E3779B97F4A7C15ull; return mix64(r->s); }
static inline double runif(rng_t* r) { return (double)(rnext(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int64_t rrange(rng_t* r, int64_t lo, int64_t hi) { /* inclusive */
    return lo + (int64_t)(rnext(r) % (uint64_t)(hi - lo + 1));
}
static inline double rnorm(rng_t* r) {
    double u1 = runif(r), u2 = runif(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

/* ---------- growable buffers ---------- */
typedef struct { uint8_t* p; size_t n, cap; } buf_t;
static int buf_reserve(buf_t* b, size_t extra) {
    if (b->n + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap * 2 : 4096;
    while (nc < b->n + extra) nc *= 2;
    uint8_t* q = (uint8_t*)realloc(b->p, nc);
    if (!q) return -1;
    b->p = q; b->cap = nc; return 0;
}
static void* buf_push(buf_t* b, const void* src, size_t nbytes) {
    if (buf_reserve(b, nbytes)) { fprintf(stderr, "synth: out of memory\n"); abort(); }
    void* dst = b->p + b->n;
    if (src) memcpy(dst, src, nbytes);
    b->n += nbytes;
    return dst;
}

/* inserted-sequence placement, resolved when the seq arena is written */
typedef struct { int32_t qpos, len, site, _p; uint64_t noise_seed; } insrec_t;

typedef struct {
    snfb_rec r;           /* offsets are local to the generating contig's arenas */
    uint32_t ins_off, ins_n;
    uint64_t order;       /* generation order, tie break of the final sort */
    uint64_t seq_seed;
    int32_t  src;         /* generating contig (owner of the local arenas) */
    int32_t  qalen;       /* query_alignment_length */
} grec_t;

typedef struct {
    buf_t recs, cigar, var, ins;
} local_t;

struct snfb_synth_block {
    snfb_records R;
    snfb_rec* rec; uint32_t* cigar; uint8_t* var; uint8_t* seq;
    snfb_task* task; snfb_contig* contig; int32_t* tr;
    snfb_synth_site* sites; uint64_t n_sites;
    uint64_t aligned_bp;
};

/* ---------- op model ---------- */
#define GEO_TAB 4096
typedef struct {
    const snfb_synth_params* p;
    int32_t geo[GEO_TAB];
    const snfb_synth_site* sites; /* all sites, sorted by (contig,pos) */
    const int64_t* site_first;    /* [n_contig+1] */
} model_t;

static inline void push_op(buf_t* cg, size_t floor, uint32_t len, uint32_t op) {
    if (len == 0) return;
    /* merge with a previous op of the same kind (never across the record's first op) */
    if (cg->n >= floor + 4) {
        uint32_t* last = (uint32_t*)(cg->p + cg->n - 4);
        if ((*last & 15u) == op && (uint64_t)(*last >> 4) + len < (1u << 28)) { *last += len << 4; return; }
    }
    uint32_t v = (len << 4) | op;
    buf_push(cg, &v, 4);
}

/* emit noisy alignment ops covering ref [from,to); returns query bases consumed */
static int64_t gen_noise(const model_t* m, rng_t* r, buf_t* cg, size_t floor, int64_t from, int64_t to, int64_t* nm_small) {
    int64_t q = 0, pos = from;
    while (pos < to) {
        int64_t run = m->geo[rnext(r) & (GEO_TAB - 1)];
        if (run > to - pos) run = to - pos;
        push_op(cg, floor, (uint32_t)run, 0); q += run; pos += run;
        if (pos >= to) break;
        uint64_t u = rnext(r);
        int len; unsigned c = (unsigned)(u & 1023);
        if (c < 717) len = 1; else if (c < 922) len = 2; else len = 3 + (int)((u >> 10) % 8);
        if ((u >> 20) & 1) { push_op(cg, floor, (uint32_t)len, 1); q += len; }
        else { if (len > to - pos) len = (int)(to - pos); push_op(cg, floor, (uint32_t)len, 2); pos += len; }
        *nm_small += len;
    }
    return q;
}

static size_t fmt_sa(char* out, int contig, int64_t pos0, int rev, int64_t clipL, int64_t q, int64_t rspan, int64_t clipR, int mapq, int nm) {
    char cg[96]; size_t k = 0;
    int64_t mm = q < rspan ? q : rspan;
    if (clipL > 0) k += (size_t)sprintf(cg + k, "%lldS", (long long)clipL);
    k += (size_t)sprintf(cg + k, "%lldM", (long long)mm);
    if (q > rspan) k += (size_t)sprintf(cg + k, "%lldI", (long long)(q - rspan));
    else if (rspan > q) k += (size_t)sprintf(cg + k, "%lldD", (long long)(rspan - q));
    if (clipR > 0) k += (size_t)sprintf(cg + k, "%lldS", (long long)clipR);
    return (size_t)sprintf(out, "ctg%d,%lld,%c,%s,%d,%d;", contig + 1, (long long)(pos0 + 1), rev ? '-' : '+', cg, mapq, nm);
}

typedef struct {
    int contig; int64_t pos; int rev; int64_t clipL, q, rspan, clipR; int mapq, nm;
    size_t cig_from, cig_to;     /* byte range in local cigar buf (alignment ops only, without clips) */
    uint32_t ins_off, ins_n;
} seg_t;

static void emit_record(local_t* L, const seg_t* s, const seg_t* other, int supplementary, int secondary,
                        int64_t Q, int qn_contig, int64_t qn_idx, int hp, int ps, int has_phase, uint64_t seq_seed,
                        int src, uint64_t order) {
    grec_t g; memset(&g, 0, sizeof g);
    g.src = src; g.order = order; g.seq_seed = seq_seed;
    g.r.task = s->contig; g.r.pos = (int32_t)s->pos; g.r.mapq = (uint8_t)s->mapq;
    g.r.flag = (uint16_t)((s->rev ? 16 : 0) | (supplementary ? 2048 : 0) | (secondary ? 256 : 0));
    g.r.aux_flags = SNFB_AUX_NM; g.r.nm = s->nm;
    if (has_phase) { g.r.aux_flags |= SNFB_AUX_HP | SNFB_AUX_PS; g.r.hp = (uint8_t)hp; g.r.ps = ps; }
    g.r.l_seq = (int32_t)Q;
    /* cigar: clipL + ops + clipR, copied to the tail of the local cigar buffer */
    size_t nops_bytes = s->cig_to - s->cig_from;
    size_t start = L->cigar.n;
    if (s->clipL > 0) { uint32_t v = ((uint32_t)s->clipL << 4) | 4u; buf_push(&L->cigar, &v, 4); }
    buf_reserve(&L->cigar, nops_bytes);
    memmove(L->cigar.p + L->cigar.n, L->cigar.p + s->cig_from, nops_bytes); L->cigar.n += nops_bytes;
    if (s->clipR > 0) { uint32_t v = ((uint32_t)s->clipR << 4) | 4u; buf_push(&L->cigar, &v, 4); }
    g.r.cigar_off = start / 4; g.r.n_cigar = (uint32_t)((L->cigar.n - start) / 4);
    g.qalen = (int32_t)s->q;
    /* var: qname + SA */
    char tmp[256];
    int ql = sprintf(tmp, "r%d_%lld", qn_contig + 1, (long long)qn_idx);
    g.r.var_off = L->var.n; g.r.l_qname = (uint8_t)ql; buf_push(&L->var, tmp, (size_t)ql);
    if (other) {
        size_t sl = fmt_sa(tmp, other->contig, other->pos, other->rev, other->clipL, other->q, other->rspan, other->clipR, other->mapq, other->nm);
        buf_push(&L->var, tmp, sl); g.r.sa_len = (uint32_t)sl; g.r.aux_flags |= SNFB_AUX_SA;
    }
    g.ins_off = s->ins_off; g.ins_n = s->ins_n;
    buf_push(&L->recs, &g, sizeof g);
}

static inline int carries(const snfb_synth_site* st, int read_hap, uint64_t h) {
    if (st->vaf < 1.0) return (double)(h >> 11) * (1.0 / 9007199254740992.0) < st->vaf;
    return st->hap == 0 || st->hap == read_hap;
}

/* the first draws of a read: its length and start (shared by gen_read and the ownership pre-check) */
static inline void read_span(const snfb_synth_params* p, rng_t* r, int c, int64_t idx, int64_t nreads, int64_t* start_out, int64_t* len_out) {
    int64_t clen = p->contig_len[c];
    double len;
    if (p->len_model == 1) { double sg = p->len_sd / 1000.0; len = exp(log(p->len_mean) - 0.5 * sg * sg + sg * rnorm(r)); }
    else len = p->len_mean + p->len_sd * rnorm(r);
    int64_t Lr = (int64_t)len;
    if (Lr < p->len_min) Lr = p->len_min; if (Lr > p->len_max) Lr = p->len_max;
    if (Lr > clen - 2) Lr = clen - 2;
    int64_t span = clen - Lr; if (span < 1) span = 1;
    int64_t start = (int64_t)(((double)idx + runif(r)) * (double)span / (double)nreads);
    if (start > clen - Lr - 1) start = clen - Lr - 1; if (start < 0) start = 0;
    *start_out = start; *len_out = Lr;
}

/* Sharded generation (contig_mask): a read that starts on a contig this shard does not own still matters when it can
 * split at a BND site whose mate lies on an owned contig (its second record lands there).  Superset test: the read's span
 * holds such a site.  Everything else of an unowned contig is never generated. */
static int read_may_reach_owned(const model_t* m, int c, int64_t idx, int64_t nreads) {
    const snfb_synth_params* p = m->p;
    rng_t r = { mix64(p->seed ^ mix64(p->sample * 0xD1B54A32D192ED03ull) * (p->sample != 0) ^ mix64(((uint64_t)(uint32_t)c << 40) ^ (uint64_t)idx)) };
    int64_t start, Lr; read_span(p, &r, c, idx, nreads, &start, &Lr);
    int64_t s0 = m->site_first[c], s1 = m->site_first[c + 1];
    int64_t lo = s0, hi = s1;
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (m->sites[mid].pos <= start + 290) lo = mid + 1; else hi = mid; }
    for (int64_t si = lo; si < s1; ++si) {
        const snfb_synth_site* st = &m->sites[si];
        if (st->pos >= start + Lr + 100000) break;     /* in-read deletions extend `end`: a generous bound keeps this a superset */
        if (st->svtype == SNFB_BND && p->contig_mask[st->mate_contig]) return 1;
    }
    return 0;
}

static void gen_read(const model_t* m, local_t* L, int c, int64_t idx, int64_t nreads, int src) {
    const snfb_synth_params* p = m->p;
    rng_t r = { mix64(p->seed ^ mix64(p->sample * 0xD1B54A32D192ED03ull) * (p->sample != 0) ^ mix64(((uint64_t)(uint32_t)c << 40) ^ (uint64_t)idx)) };
    int64_t clen = p->contig_len[c];
    int64_t start, Lr; read_span(p, &r, c, idx, nreads, &start, &Lr);
    int rev = (int)(rnext(&r) & 1);
    int mapq = runif(&r) < p->lowmapq_prob ? (int)rrange(&r, 0, 59) : 60;
    int secondary = runif(&r) < p->secondary_prob;
    int read_hap = 1 + (int)(rnext(&r) & 1);
    int has_phase = runif(&r) < p->phased_frac;
    int ps = (int)((start / 500000) * 500000 + 1);
    uint64_t seq_seed = rnext(&r);
    int64_t end = start + Lr;

    /* scratch region at the tail of the local cigar buffer for the raw alignment ops */
    size_t scratch0 = L->cigar.n;
    uint32_t ins0 = (uint32_t)(L->ins.n / sizeof(insrec_t));
    int64_t nm_small = 0, big = 0, q = 0, pos = start;
    int split = 0; seg_t A, B; memset(&A, 0, sizeof A); memset(&B, 0, sizeof B);
    int64_t gap = 0; int gap_site = -1;

    int64_t s0 = m->site_first[c], s1 = m->site_first[c + 1];
    /* first site with pos > start+300 */
    int64_t lo = s0, hi = s1;
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (m->sites[mid].pos <= start + 300) lo = mid + 1; else hi = mid; }
    for (int64_t si = lo; si < s1 && !split; ++si) {
        const snfb_synth_site* st = &m->sites[si];
        if (st->pos >= end - 300) break;
        uint64_t h = mix64(seq_seed ^ (uint64_t)si * 0x9E3779B97F4A7C15ull);
        if (!carries(st, read_hap, h)) continue;
        if (p->site_keep > 0.0 && (double)(mix64(p->seed ^ mix64(p->sample + 0x51ull) ^ (uint64_t)si * 0xA24BAED4963EE407ull) >> 11) * (1.0 / 9007199254740992.0) >= p->site_keep) continue;
        int jit[5] = { 0, 0, 0, (int)(h & 1) ? 1 : -1, (int)(h & 2) ? 2 : -2 };
        int64_t sp = st->pos + jit[(h >> 8) % 5];
        if (sp <= pos + 10 || sp >= end - 10) continue;
        if (st->in_tr && ((h >> 16) & 3) == 0) { /* extra noise event inside the repeat */
            int64_t np = sp - (int64_t)((h >> 20) % 200) - 60;
            if (np > pos + 5) {
                q += gen_noise(m, &r, &L->cigar, scratch0, pos, np, &nm_small); pos = np;
                int nl = 45 + (int)((h >> 32) % 76);
                if ((h >> 40) & 1) { push_op(&L->cigar, scratch0, (uint32_t)nl, 1);
                    insrec_t ir = { (int32_t)q, nl, -1, 0, h }; buf_push(&L->ins, &ir, sizeof ir); q += nl; }
                else { push_op(&L->cigar, scratch0, (uint32_t)nl, 2); pos += nl; }
                big += nl;
            }
        }
        int64_t size = st->size;
        int inline_ok = (st->svtype == SNFB_INS || st->svtype == SNFB_DEL) && size <= 5000 && size * 3 < Lr;
        if (inline_ok) {
            if (sp <= pos) continue;
            q += gen_noise(m, &r, &L->cigar, scratch0, pos, sp, &nm_small); pos = sp;
            double f = 0.97 + 0.06 * ((double)((h >> 12) & 0xFFFF) / 65535.0);
            int64_t sz = (int64_t)((double)size * f + 0.5); if (sz < 1) sz = 1;
            if (st->svtype == SNFB_INS) {
                push_op(&L->cigar, scratch0, (uint32_t)sz, 1);
                insrec_t ir = { (int32_t)q, (int32_t)sz, (int32_t)si, 0, h }; buf_push(&L->ins, &ir, sizeof ir);
                q += sz;
            } else {
                if (pos + sz >= clen - 50) continue;
                push_op(&L->cigar, scratch0, (uint32_t)sz, 2); pos += sz; if (end < pos + 50) end = pos + 50;
            }
            big += sz;
        } else {
            /* split alignment at sp; needs >= 1200 aligned bases on both sides */
            if (sp - start < 1200 || end - sp < 1200) continue;
            q += gen_noise(m, &r, &L->cigar, scratch0, pos, sp, &nm_small); pos = sp;
            A.contig = c; A.pos = start; A.rev = rev; A.q = q; A.rspan = sp - start; A.mapq = mapq;
            A.cig_from = scratch0; A.cig_to = L->cigar.n; A.nm = (int)(nm_small + big);
            int64_t rem = end - sp;
            B.mapq = (h >> 50) % 20 == 0 ? (int)((h >> 44) % 20) : 60;
            B.contig = c; B.rev = rev;
            int64_t bfrom;
            switch (st->svtype) {
            case SNFB_DEL: bfrom = sp + size; break;
            case SNFB_INS: bfrom = sp; gap = size; gap_site = (int)si; break;
            case SNFB_DUP: bfrom = sp - size; break;
            case SNFB_INV: bfrom = sp + size - rem; B.rev = !rev; break;
            default: /* BND */ B.contig = st->mate_contig; bfrom = st->mate_pos; B.rev = !rev; break;
            }
            int64_t blen = p->contig_len[B.contig];
            if (bfrom < 0) bfrom = 0;
            if (bfrom + rem > blen - 1) rem = blen - 1 - bfrom;
            if (rem < 1000) { /* cannot place the mate: keep as plain read */
                continue;
            }
            B.pos = bfrom; B.rspan = rem;
            B.cig_from = L->cigar.n;
            int64_t nmB = 0;
            B.q = gen_noise(m, &r, &L->cigar, B.cig_from, bfrom, bfrom + rem, &nmB);
            B.cig_to = L->cigar.n; B.nm = (int)nmB;
            split = 1;
        }
    }
    uint64_t order = ((uint64_t)(uint32_t)c << 40) | ((uint64_t)idx << 1);
    uint32_t ins1 = (uint32_t)(L->ins.n / sizeof(insrec_t));
    if (!split) {
        q += gen_noise(m, &r, &L->cigar, scratch0, pos, end, &nm_small);
        seg_t S; memset(&S, 0, sizeof S);
        S.contig = c; S.pos = start; S.rev = rev; S.q = q; S.rspan = end - start; S.mapq = mapq;
        S.cig_from = scratch0; S.cig_to = L->cigar.n;
        if (runif(&r) < p->clip_prob) { int64_t cl = rrange(&r, 50, 2000); if (rnext(&r) & 1) S.clipL = cl; else S.clipR = cl; }
        double subs = (double)Lr * p->nm_rate;
        S.nm = (int)(nm_small + big + (int64_t)(subs + sqrt(subs) * rnorm(&r) + 0.5)); if (S.nm < 0) S.nm = 0;
        S.ins_off = ins0; S.ins_n = ins1 - ins0;
        /* query positions of insertions shift by the leading clip */
        for (uint32_t k = ins0; k < ins1; ++k) ((insrec_t*)L->ins.p)[k].qpos += (int32_t)S.clipL;
        emit_record(L, &S, NULL, 0, secondary, S.clipL + q + S.clipR, c, idx, read_hap, ps, has_phase, seq_seed, src, order);
        /* drop the scratch ops that were copied: compact by moving the record's ops down */
        grec_t* g = (grec_t*)(L->recs.p + L->recs.n - sizeof(grec_t));
        size_t nbytes = (size_t)g->r.n_cigar * 4;
        memmove(L->cigar.p + scratch0, L->cigar.p + g->r.cigar_off * 4, nbytes);
        g->r.cigar_off = scratch0 / 4; L->cigar.n = scratch0 + nbytes;
        return;
    }
    int64_t Q = A.q + gap + B.q;
    /* BAM-order clips (see DESIGN.md "synthetic split reads"): same strand: A=[ops,S] B=[S,ops];
     * opposite strand mate: B=[ops,S] */
    A.clipL = 0; A.clipR = Q - A.q;
    if (B.rev == A.rev) { B.clipL = Q - B.q; B.clipR = 0; } else { B.clipL = 0; B.clipR = Q - B.q; }
    A.ins_off = ins0; A.ins_n = ins1 - ins0;
    if (gap > 0) {
        /* where the reference slices query_sequence for a split INS: [last.qry_end, curr.qry_start)
         * (sv.py:682,694) in the coordinates it derives; place the site's sequence there in the primary */
        int a_primary = A.q >= B.q;
        int32_t qp = (int32_t)(A.rev ? B.q : A.q);
        insrec_t ir = { qp, (int32_t)gap, gap_site, 0, mix64(seq_seed ^ 77) };
        buf_push(&L->ins, &ir, sizeof ir);
        uint32_t gi = (uint32_t)(L->ins.n / sizeof(insrec_t)) - 1;
        if (a_primary) { A.ins_n += 1; }
        else { B.ins_off = gi; B.ins_n = 1; }
    }
    int a_primary = A.q >= B.q;
    emit_record(L, &A, &B, !a_primary, secondary, Q, c, idx, read_hap, ps, has_phase, seq_seed, src, order);
    emit_record(L, &B, &A, a_primary, secondary, Q, c, idx, read_hap, ps, has_phase, mix64(seq_seed ^ 5), src, order | 1);
    /* compact: the two emitted records sit after the scratch ops; move them down */
    grec_t* gb = (grec_t*)(L->recs.p + L->recs.n - sizeof(grec_t));
    grec_t* ga = gb - 1;
    size_t na = (size_t)ga->r.n_cigar * 4, nb = (size_t)gb->r.n_cigar * 4;
    size_t from = ga->r.cigar_off * 4;
    memmove(L->cigar.p + scratch0, L->cigar.p + from, na + nb);
    ga->r.cigar_off = scratch0 / 4; gb->r.cigar_off = (scratch0 + na) / 4; L->cigar.n = scratch0 + na + nb;
}

/* ---------- sites ---------- */
static int cmp_site(const void* a, const void* b) {
    const snfb_synth_site* x = (const snfb_synth_site*)a; const snfb_synth_site* y = (const snfb_synth_site*)b;
    if (x->contig != y->contig) return x->contig < y->contig ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

static const grec_t** g_sort_base;
static int cmp_rec(const void* a, const void* b) {
    const grec_t* x = *(const grec_t* const*)a; const grec_t* y = *(const grec_t* const*)b;
    if (x->r.task != y->r.task) return x->r.task < y->r.task ? -1 : 1;
    if (x->r.pos != y->r.pos) return x->r.pos < y->r.pos ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}

static inline uint8_t site_base(int32_t site, int64_t i) {
    static const uint8_t code[4] = { 1, 2, 4, 8 };
    return code[mix64(((uint64_t)(uint32_t)site << 32) ^ (uint64_t)i ^ 0xABCDull) & 3];
}

uint64_t snfb_hash_name_host(const char* s, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= (uint8_t)s[i]; h *= 0x100000001b3ull; }
    return h;
}

snfb_synth_block* snfb_synth_generate(const snfb_synth_params* p) {
    snfb_synth_block* blk = (snfb_synth_block*)calloc(1, sizeof *blk);
    if (!blk) return NULL;
    int nc = p->n_contig;
    model_t m; m.p = p;
    { /* geometric quantile table */
        double pr = 1.0 / (p->op_mean_run > 1 ? p->op_mean_run : 1.0);
        for (int i = 0; i < GEO_TAB; ++i) {
            double u = ((double)i + 0.5) / GEO_TAB;
            double v = floor(log(1.0 - u) / log(1.0 - pr)) + 1.0;
            if (v < 1) v = 1; if (v > 1e6) v = 1e6;
            m.geo[i] = (int32_t)v;
        }
        /* decorrelate table order from low rng bits */
        rng_t r = { mix64(p->seed ^ 0x5151) };
        for (int i = GEO_TAB - 1; i > 0; --i) { int j = (int)(rnext(&r) % (uint64_t)(i + 1)); int32_t t = m.geo[i]; m.geo[i] = m.geo[j]; m.geo[j] = t; }
    }
    /* sites */
    uint64_t ns_cap = 16; for (int c = 0; c < nc; ++c) ns_cap += (uint64_t)(p->contig_len[c] / (p->sv_spacing > 1 ? p->sv_spacing : 1)) + 2;
    snfb_synth_site* sites = (snfb_synth_site*)calloc(ns_cap, sizeof *sites);
    buf_t trbuf = { 0 }; int64_t* tr_first = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
    uint64_t ns = 0;
    for (int c = 0; c < nc; ++c) {
        rng_t r = { mix64(p->seed ^ mix64(0xC0FFEEull + (uint64_t)c)) };
        int64_t n = (int64_t)(p->contig_len[c] / p->sv_spacing);
        tr_first[c] = (int64_t)(trbuf.n / 8);
        int64_t last_tr_end = -1;
        for (int64_t i = 0; i < n; ++i) {
            snfb_synth_site s; memset(&s, 0, sizeof s);
            s.contig = c;
            s.pos = (int32_t)(((double)i + 0.15 + 0.7 * runif(&r)) * p->sv_spacing);
            if (s.pos < 3000 || s.pos > p->contig_len[c] - 3000) continue;
            double u = runif(&r);
            if (p->ins_only) s.svtype = SNFB_INS;
            else s.svtype = u < 0.45 ? SNFB_INS : u < 0.90 ? SNFB_DEL : u < 0.94 ? SNFB_DUP : u < 0.97 ? SNFB_INV : SNFB_BND;
            double lo = log((double)p->sv_min), hi = log((double)p->sv_max);
            double sz = exp(lo + (hi - lo) * runif(&r));
            if (!p->ins_only && runif(&r) < 0.10) sz = exp(log(5000.0) + (log(50000.0) - log(5000.0)) * runif(&r));
            s.size = (int32_t)sz; if (s.size < 1) s.size = 1;
            if (p->mosaic && runif(&r) < 0.8) { s.vaf = 0.05 + 0.15 * runif(&r); s.hap = 0; }
            else { s.vaf = 1.0; s.hap = runif(&r) < (2.0 / 3.0) ? 1 + (int)(rnext(&r) & 1) : 0; }
            if (p->ins_only) { s.vaf = 1.0; s.hap = 0; }
            if (s.svtype == SNFB_BND) {
                s.mate_contig = nc > 1 ? (int)((c + 1 + (int)(rnext(&r) % (uint64_t)(nc - 1))) % nc) : c;
                int64_t ml = p->contig_len[s.mate_contig];
                s.mate_pos = (int32_t)(2000 + (int64_t)(runif(&r) * (double)(ml > 8000 ? ml - 8000 : 1)));
            }
            if (runif(&r) < p->tr_frac) {
                s.in_tr = 1;
                int32_t a = s.pos - (int32_t)rrange(&r, 100, 1000), b = s.pos + (int32_t)rrange(&r, 100, 1000);
                a = a - 500 < 0 ? 0 : a - 500; b = b + 500; /* util.load_tandem_repeats pads +-500 */
                if (a > last_tr_end) { int32_t pr2[2] = { a, b }; buf_push(&trbuf, pr2, 8); last_tr_end = b; }
                else s.in_tr = 0;
            }
            sites[ns++] = s;
        }
    }
    tr_first[nc] = (int64_t)(trbuf.n / 8);
    qsort(sites, ns, sizeof *sites, cmp_site);
    int64_t* site_first = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
    { uint64_t k = 0; for (int c = 0; c <= nc; ++c) { while (k < ns && sites[k].contig < c) ++k; site_first[c] = (int64_t)k; } site_first[nc] = (int64_t)ns; }
    m.sites = sites; m.site_first = site_first;

    /* reads: work units of at most UNIT reads, one local arena each (a read depends only on (seed, contig, index)) */
    enum { UNIT = 4096 };
    int64_t* nreads_c = (int64_t*)calloc((size_t)nc + 1, sizeof(int64_t));
    int64_t nunits = 0;
    uint8_t* gen_c = (uint8_t*)calloc((size_t)nc + 1, 1);      /* 1 = owned (every read), 2 = unowned but some BND site reaches an owned contig */
    for (int c = 0; c < nc; ++c) {
        int64_t nr = (int64_t)(p->coverage * (double)p->contig_len[c] / p->len_mean + 0.5); if (nr < 1) nr = 1;
        nreads_c[c] = nr;
        if (!p->contig_mask || p->contig_mask[c]) gen_c[c] = 1;
        else for (int64_t si = site_first[c]; si < site_first[c + 1]; ++si) if (sites[si].svtype == SNFB_BND && p->contig_mask[sites[si].mate_contig]) { gen_c[c] = 2; break; }
        if (gen_c[c]) nunits += (nr + UNIT - 1) / UNIT;
    }
    int32_t* unit_c = (int32_t*)malloc((size_t)(nunits + 1) * sizeof(int32_t)); int64_t* unit_i0 = (int64_t*)malloc((size_t)(nunits + 1) * sizeof(int64_t));
    { int64_t u = 0; for (int c = 0; c < nc; ++c) if (gen_c[c]) for (int64_t i0 = 0; i0 < nreads_c[c]; i0 += UNIT) { unit_c[u] = c; unit_i0[u] = i0; ++u; } }
    local_t* loc = (local_t*)calloc((size_t)nunits + 1, sizeof *loc);
#ifdef _OPENMP
    if (p->threads > 0) omp_set_num_threads(p->threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t u = 0; u < nunits; ++u) {
        const int c = unit_c[u]; const int64_t i0 = unit_i0[u], i1 = i0 + UNIT < nreads_c[c] ? i0 + UNIT : nreads_c[c];
        for (int64_t i = i0; i < i1; ++i) {
            if (gen_c[c] == 2 && !read_may_reach_owned(&m, c, i, nreads_c[c])) continue;
            gen_read(&m, &loc[u], c, i, nreads_c[c], (int)u);
        }
    }
    /* global order */
    uint64_t nrec = 0; for (int64_t c = 0; c < nunits; ++c) nrec += loc[c].recs.n / sizeof(grec_t);
    const grec_t** ord = (const grec_t**)malloc((nrec ? nrec : 1) * sizeof *ord);
    { uint64_t k = 0; for (int64_t c = 0; c < nunits; ++c) { grec_t* g = (grec_t*)loc[c].recs.p; uint64_t n = loc[c].recs.n / sizeof(grec_t);
        for (uint64_t i = 0; i < n; ++i) if (!p->contig_mask || p->contig_mask[g[i].r.task]) ord[k++] = &g[i]; }
      nrec = k; }
    g_sort_base = ord;
    qsort(ord, nrec, sizeof *ord, cmp_rec);
    /* offsets */
    uint64_t* coff = (uint64_t*)malloc((nrec + 1) * 8), *voff = (uint64_t*)malloc((nrec + 1) * 8), *soff = (uint64_t*)malloc((nrec + 1) * 8);
    coff[0] = voff[0] = soff[0] = 0; uint64_t abp = 0;
    for (uint64_t i = 0; i < nrec; ++i) {
        const grec_t* g = ord[i];
        coff[i + 1] = coff[i] + g->r.n_cigar;
        voff[i + 1] = voff[i] + g->r.l_qname + g->r.sa_len;
        soff[i + 1] = soff[i] + (uint64_t)((g->r.l_seq + 1) / 2);
        abp += (uint64_t)g->qalen;
    }
    blk->aligned_bp = abp;
    blk->rec = (snfb_rec*)malloc((nrec ? nrec : 1) * sizeof(snfb_rec));
    blk->cigar = (uint32_t*)malloc((coff[nrec] ? coff[nrec] : 1) * 4);
    blk->var = (uint8_t*)malloc(voff[nrec] ? voff[nrec] : 1);
    blk->seq = (uint8_t*)calloc(soff[nrec] ? soff[nrec] : 1, 1);
    if (!blk->rec || !blk->cigar || !blk->var || !blk->seq) { fprintf(stderr, "synth: out of memory\n"); abort(); }
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)nrec; ++i) {
        const grec_t* g = ord[i]; const local_t* L = &loc[g->src];
        snfb_rec r = g->r;
        memcpy(blk->cigar + coff[i], L->cigar.p + g->r.cigar_off * 4, (size_t)r.n_cigar * 4);
        memcpy(blk->var + voff[i], L->var.p + g->r.var_off, (size_t)r.l_qname + r.sa_len);
        r.cigar_off = coff[i]; r.var_off = voff[i]; r.seq_off = soff[i];
        blk->rec[i] = r;
        if (p->with_seq) {
            uint8_t* sq = blk->seq + soff[i]; int64_t nb = (r.l_seq + 1) / 2;
            rng_t rr = { g->seq_seed };
            static const uint8_t pair[16] = { 0x11, 0x12, 0x14, 0x18, 0x21, 0x22, 0x24, 0x28, 0x41, 0x42, 0x44, 0x48, 0x81, 0x82, 0x84, 0x88 };
            int64_t b = 0;
            for (; b + 16 <= nb; b += 16) { uint64_t x = rnext(&rr); for (int k = 0; k < 16; ++k) sq[b + k] = pair[(x >> (4 * k)) & 15]; }
            { uint64_t x = rnext(&rr); for (int k = 0; b < nb; ++b, ++k) sq[b] = pair[(x >> (4 * k)) & 15]; }
            if (r.l_seq & 1) sq[nb - 1] &= 0xF0;
            const insrec_t* ir = (const insrec_t*)L->ins.p + g->ins_off;
            for (uint32_t k = 0; k < g->ins_n; ++k) {
                rng_t nr = { ir[k].noise_seed };
                for (int32_t j = 0; j < ir[k].len; ++j) {
                    int64_t qp = (int64_t)ir[k].qpos + j; if (qp < 0 || qp >= r.l_seq) continue;
                    uint8_t base;
                    if (ir[k].site < 0) base = (uint8_t)(1u << (rnext(&nr) & 3));
                    else { base = site_base(ir[k].site, j); uint64_t x = rnext(&nr); if ((double)(x >> 11) * (1.0 / 9007199254740992.0) < p->ins_noise) base = (uint8_t)(1u << ((x >> 3) & 3)); }
                    uint8_t* cell = sq + (qp >> 1);
                    if (qp & 1) *cell = (uint8_t)((*cell & 0xF0) | base); else *cell = (uint8_t)((*cell & 0x0F) | (base << 4));
                }
            }
        }
    }
    /* tables */
    blk->task = (snfb_task*)calloc((size_t)nc, sizeof(snfb_task));
    blk->contig = (snfb_contig*)calloc((size_t)nc, sizeof(snfb_contig));
    char** names = (char**)malloc((size_t)nc * sizeof(char*));
    for (int c = 0; c < nc; ++c) { names[c] = (char*)malloc(32); sprintf(names[c], "ctg%d", c + 1); }
    for (int c = 0; c < nc; ++c) {
        int rank = 0; for (int d = 0; d < nc; ++d) if (strcmp(names[d], names[c]) < 0) ++rank;
        blk->contig[c].name_hash = snfb_hash_name_host(names[c], strlen(names[c]));
        blk->contig[c].length = p->contig_len[c]; blk->contig[c].lex_rank = rank;
        blk->task[c].contig = c; blk->task[c].start = 0; blk->task[c].end = p->contig_len[c] - 1; /* sniffles:313-358 */
        blk->task[c].contig_len = p->contig_len[c]; blk->task[c].task_id = c;
        blk->task[c].tr_off = (int32_t)tr_first[c]; blk->task[c].tr_n = (int32_t)(tr_first[c + 1] - tr_first[c]);
    }
    for (int c = 0; c < nc; ++c) free(names[c]); free(names);
    blk->tr = (int32_t*)malloc(trbuf.n ? trbuf.n : 8); memcpy(blk->tr, trbuf.p, trbuf.n);
    blk->sites = sites; blk->n_sites = ns;
    snfb_records* R = &blk->R;
    R->n_rec = nrec; R->n_cigar = coff[nrec]; R->n_var = voff[nrec]; R->n_seq = soff[nrec];
    R->rec = blk->rec; R->cigar = blk->cigar; R->var = blk->var; R->seq = blk->seq;
    R->n_task = (uint32_t)nc; R->n_contig = (uint32_t)nc; R->n_tr = (uint32_t)(trbuf.n / 8); R->on_device = 0;
    R->task = blk->task; R->contig = blk->contig; R->tr = blk->tr;
    for (int64_t c = 0; c < nunits; ++c) { free(loc[c].recs.p); free(loc[c].cigar.p); free(loc[c].var.p); free(loc[c].ins.p); }
    free(loc); free(gen_c); free(nreads_c); free(unit_c); free(unit_i0); free(ord); free(coff); free(voff); free(soff); free(site_first); free(tr_first); free(trbuf.p);
    return blk;
}

const snfb_records* snfb_synth_records(const snfb_synth_block* b) { return &b->R; }
uint64_t snfb_synth_sites(const snfb_synth_block* b, const snfb_synth_site** out) { if (out) *out = b->sites; return b->n_sites; }
uint64_t snfb_synth_aligned_bp(const snfb_synth_block* b) { return b->aligned_bp; }
void snfb_synth_free(snfb_synth_block* b) {
    if (!b) return;
    free(b->rec); free(b->cigar); free(b->var); free(b->seq); free(b->task); free(b->contig); free(b->tr); free(b->sites); free(b);
}
