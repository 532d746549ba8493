// extract.cuh — stage A: alignment records -> SV leads.
//
// One warp per alignment record streams the record's CIGAR16 words with 16-byte loads (256 words
// per warp step, coalesced); SV signatures become 64-byte leads in a second, event-driven pass.
// Reference behaviour reproduced (paths relative to /root/reference/src/sniffles/):
//   LeadProvider.iter_region filters / NM / coverage bookkeeping     leadprov.py:474-581
//   get_cigar_indels                                                  leadprov.py:198-224
//   read_iterindels                                                   leadprov.py:583-670
//   Lead.for_bnd + CIGAR_analyze                                      leadprov.py:57-176
//   read_itersplits + sv.classify_splits                              leadprov.py:227-355, sv.py:649-782
//   build_leadtab region filter                                       leadprov.py:464-468
#pragma once
#include "common.cuh"

namespace extract {

constexpr int MAXSEG = 40;     // primary + supplementary segments per read held in shared memory

// rec_flags bits
constexpr uint8_t RF_PASS = 1, RF_HAS_NM = 2;   // bits 2..3: hp

struct Seg {
    int contig, ref_start, ref_end, qry_start, qry_end;
    int meta;                 // bit0 rev, bits 8..15 mapq, bits 16..17 source, bit 20 has_seq
    int nhint;
    int h_type[2], h_start[2], h_len[2], h_none[2];
    int seq_off, seq_len;
};

__device__ __forceinline__ bool op_adds_read(int op) { return (0x193u >> op) & 1u; }   // M I S = X  (0,1,4,7,8)
__device__ __forceinline__ bool op_adds_ref(int op) { return (0x18Du >> op) & 1u; }    // M D N = X  (0,2,3,7,8)
__device__ __forceinline__ bool op_is_event(int op) { return (0x016u >> op) & 1u; }    // I D S      (1,2,4)

__device__ __forceinline__ void store_lead(snfb_lead* dst, const snfb_lead& l) {
    const uint4* s = reinterpret_cast<const uint4*>(&l); uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

// ---- text helpers (lane-serial; SA tags are short in the compact form aligners write) ----
__device__ inline bool parse_int_dev(const uint8_t* s, int n, long long* out) {
    if (n <= 0) return false; long long v = 0; int i = 0; bool neg = false;
    if (s[0] == '-' || s[0] == '+') { neg = s[0] == '-'; i = 1; if (n == 1) return false; }
    for (; i < n; ++i) { int c = s[i]; if (c < '0' || c > '9') return false; v = v * 10 + (c - '0'); if (v > (1LL << 40)) return false; }
    *out = neg ? -v : v; return true;
}
// leadprov.CIGAR_analyze
__device__ inline bool cigar_analyze_dev(const uint8_t* c, int n, long long* clip_start, long long* clip_end, long long* refspan, long long* readspan) {
    long long rs = 0, qs = 0, clip = 0, cstart = -1, val = 0; bool have = false;
    for (int i = 0; i < n; ++i) {
        int ch = c[i];
        if (ch >= '0' && ch <= '9') { val = val * 10 + (ch - '0'); have = true; if (val > (1LL << 40)) return false; continue; }
        if (!have) return false;
        bool h = false;
        if (ch == 'M' || ch == 'I' || ch == 'X' || ch == '=') { qs += val; h = true; }
        if (ch == 'M' || ch == 'D' || ch == 'X' || ch == '=' || ch == 'N') { rs += val; h = true; }
        if (!h) { if (ch == 'S' || ch == 'H') { if (cstart < 0 && qs + rs > 0) cstart = clip; clip += val; } else return false; }
        val = 0; have = false;
    }
    if (cstart < 0) cstart = clip;
    *clip_start = cstart; *clip_end = clip - cstart; *refspan = rs; *readspan = qs; return true;
}
struct SaEntry { int off[6]; int len[6]; };
__device__ inline bool sa_fields_dev(const uint8_t* s, int n, SaEntry* e) {
    int k = 0, st = 0;
    for (int i = 0; i <= n; ++i) if (i == n || s[i] == ',') { if (k == 6) return false; e->off[k] = st; e->len[k] = i - st; ++k; st = i + 1; }
    return k == 6;
}
__device__ inline int contig_lookup_dev(const snfb_contig* ct, uint32_t nct, const uint8_t* s, int n) {
    uint64_t h = fnv1a64(s, n);
    for (uint32_t i = 0; i < nct; ++i) if (ct[i].name_hash == h) return (int)i;
    return -1;
}
__device__ __forceinline__ void py_slice(long long L, long long a, long long b, int* off, int* len) {
    if (a > L) a = L; if (b > L) b = L; *off = (int)a; *len = b > a ? (int)(b - a) : 0;
}

// sv.classify_splits on the warp's shared segment list (lane-serial); returns the new count
__device__ inline int classify_splits_dev(const snfb_config& cfg, Seg* s, int n, int l_seq) {
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 1; i < n; ++i) { Seg x = s[i]; int j = i - 1; while (j >= 0 && s[j].qry_start > x.qry_start) { s[j + 1] = s[j]; --j; } s[j + 1] = x; }
        for (int i = 0; i < n; ++i) s[i].nhint = 0;
        int hints = 0; const int ms = cfg.minsvlen_screen;
        if ((double)s[0].qry_start >= (double)cfg.long_ins_length * 0.5) { s[0].h_type[0] = SNFB_INS; s[0].h_start[0] = s[0].ref_start; s[0].h_len[0] = 0; s[0].h_none[0] = 1; s[0].nhint = 1; }
        for (int i = 1; i < n; ++i) {
            Seg* cu = &s[i]; const Seg* la = &s[i - 1];
            if (cu->contig != la->contig) continue;
            const bool rev = cu->meta & 1, fwd = !rev; int ty = -1; long long st = 0, ln = 0;
            if ((cu->meta & 1) == (la->meta & 1)) {
                long long dq = (long long)cu->qry_start - la->qry_end;
                if (fwd && dq >= ms && dq - ((long long)cu->ref_start - la->ref_end) >= ms) {
                    ty = SNFB_INS; st = cu->ref_start; ln = dq;
                    if (ln <= cfg.dev_seq_cache_maxlen) { cu->meta |= 1 << 20; py_slice(l_seq, la->qry_end, cu->qry_start, &cu->seq_off, &cu->seq_len); } else cu->meta &= ~(1 << 20);
                } else if (rev && dq >= ms && dq - ((long long)la->ref_start - cu->ref_end) >= ms) {
                    ty = SNFB_INS; st = la->ref_start; ln = dq;
                    if (ln <= cfg.dev_seq_cache_maxlen) { cu->meta |= 1 << 20; py_slice(l_seq, la->qry_end, cu->qry_start, &cu->seq_off, &cu->seq_len); } else cu->meta &= ~(1 << 20);
                } else if (fwd && ((long long)cu->ref_start - la->ref_end) >= ms && ((long long)cu->ref_start - la->ref_end) - dq >= ms) {
                    ty = SNFB_DEL; st = cu->ref_start; ln = -((long long)cu->ref_start - la->ref_end);
                } else if (rev && ((long long)la->ref_start - cu->ref_end) >= ms && ((long long)la->ref_start - cu->ref_end) - dq >= ms) {
                    ty = SNFB_DEL; st = la->ref_start; ln = -((long long)la->ref_start - cu->ref_end);
                } else if (fwd && cu->ref_start <= la->ref_end) {
                    st = cu->ref_start; ln = (long long)la->ref_end - cu->ref_start; if (ln >= ms) ty = SNFB_DUP;
                } else if (rev && la->ref_start <= cu->ref_end) {
                    st = la->ref_start; ln = (long long)cu->ref_end - la->ref_start; if (ln >= ms) ty = SNFB_DUP;
                }
            } else {
                if (fwd && cu->ref_start <= la->ref_start) { st = cu->ref_start; ln = (long long)la->ref_start - cu->ref_start; if (ln >= ms) ty = SNFB_INV; }
                else if (fwd && cu->ref_start > la->ref_start) { st = la->ref_start; ln = (long long)cu->ref_start - la->ref_start; if (ln >= ms) ty = SNFB_INV; }
                else if (rev && cu->ref_end >= la->ref_end) { st = la->ref_end; ln = (long long)cu->ref_end - la->ref_end; if (ln >= ms) ty = SNFB_INV; }
                else if (rev && cu->ref_end < la->ref_end) { st = cu->ref_end; ln = (long long)la->ref_end - cu->ref_end; if (ln >= ms) ty = SNFB_INV; }
            }
            if (ty >= 0) { int k = cu->nhint++; cu->h_type[k] = ty; cu->h_start[k] = (int)st; cu->h_len[k] = (int)ln; cu->h_none[k] = 0; ++hints; }
        }
        if (!hints && n > 2) {
            int m = 0; const int c0 = s[0].contig, r0 = s[0].meta & 1;
            for (int i = 0; i < n; ++i) if (s[i].contig == c0 && (s[i].meta & 1) == r0) { Seg t = s[i]; s[m++] = t; }
            if (m == 2) { s[0].meta &= ~(1 << 20); s[1].meta &= ~(1 << 20); n = 2; continue; }   // one recursion level (sv.py:779-780)
            return m;
        }
        return n;
    }
    return n;
}

// ---- lead slot allocation.  A single global counter bumped once per lead serialises in L2; instead every thread reserves
//      SLOT_CHUNK slots at a time and hands them out locally.  Unused slots of a retired chunk are marked as holes
//      (rec == HOLE) and skipped later; canonical (record, k) order never depended on slot numbers.
constexpr unsigned SLOT_CHUNK = 8;          // per allocating thread: at most SLOT_CHUNK - 1 holes each
constexpr uint32_t HOLE = 0xffffffffu;
struct SlotState { unsigned long long cur, end; };
__device__ __forceinline__ unsigned long long alloc_slot_lane(SlotState& st, snfb_lead* leads, unsigned long long lead_cap, unsigned long long* n_slots) {   // one lane only
    if (st.cur + 1 > st.end) { const unsigned long long base = atomicAdd(n_slots, (unsigned long long)SLOT_CHUNK); st.cur = base; st.end = base + SLOT_CHUNK; }
    return st.cur++;
}

// ---- supplementary alignments (SA tag) of one record: Lead.for_bnd + read_itersplits, run by lane 0 ----
struct SaArgs {
    uint32_t rec; int qas, qae, alen, ref_end, hp; uint32_t base_flags; uint64_t qh; unsigned nlead; bool rev, is_supp;
    // the pieces of Params / snfb_rec / snfb_task the SA path needs, by value (keeps the caller's structs out of local memory)
    const uint8_t* sa; int sa_len; int clip_left, clip_right; int pos, l_seq, mapq, aux_flags, task;
    int tk_contig, tk_start, tk_end;
    const snfb_contig* contig; uint32_t n_contig; snfb_lead* leads; unsigned long long lead_cap; unsigned long long* n_slots; SlotState* slots;
    int mapq_min, dev_keep_lowqual_splits, max_splits_base; double max_splits_kb;
};
__device__ __noinline__ unsigned process_sa(const snfb_config* __restrict__ cfgp, Seg* sg, const SaArgs a, unsigned long long* soft_p, unsigned long long* overflow_p) {
    const snfb_config& cfg = *cfgp;
    const uint32_t rec = a.rec; const bool rev = a.rev;
    unsigned long long soft = 0, overflow = 0; unsigned added = 0;
    const uint8_t* sa = a.sa; const int sl = a.sa_len;
    struct { int pos, l_seq, mapq, aux_flags, task; } r = { a.pos, a.l_seq, a.mapq, a.aux_flags, a.task };
    struct { int contig, start, end; } tk = { a.tk_contig, a.tk_start, a.tk_end };
    struct { const snfb_contig* contig; uint32_t n_contig; snfb_lead* leads; unsigned long long lead_cap; } P = { a.contig, a.n_contig, a.leads, a.lead_cap };
    // pass 1: count the non-empty entries and locate the first one
    int ne = 0, f_off = 0, f_len = 0;
    for (int i = 0, st = 0; i <= sl; ++i) if (i == sl || sa[i] == ';') { if (i > st) { if (ne == 0) { f_off = st; f_len = i - st; } ++ne; } st = i + 1; }
    bool sa_ok = true; SaEntry e0;
    if (ne > 0 && !sa_fields_dev(sa + f_off, f_len, &e0)) { sa_ok = false; ++soft; }
    if (ne > 0 && sa_ok) {                                   // Lead.for_bnd: first entry only
        const uint8_t* e = sa + f_off;
        const int left = a.clip_left, right = a.clip_right;
        int bstart; bool is_first;
        if (left > right) { bstart = r.pos + 1; is_first = false; } else { bstart = a.ref_end; is_first = true; }
        const bool same = e0.len[2] == 1 && ((e[e0.off[2]] == '-' && rev) || (e[e0.off[2]] == '+' && !rev));
        if (!same) {
            long long p1, cs, ce, rs, qs, sanm = 0;
            if (!parse_int_dev(e + e0.off[1], e0.len[1], &p1)) ++soft;
            else if (!cigar_analyze_dev(e + e0.off[3], e0.len[3], &cs, &ce, &rs, &qs)) ++soft;
            else if ((r.aux_flags & SNFB_AUX_NM) && !parse_int_dev(e + e0.off[5], e0.len[5], &sanm)) ++soft;
            else {
                const long long p0 = p1 - 1; const bool is_reverse = ce > cs;
                const long long mate = is_reverse ? p0 + rs : (is_first ? p0 + 1 : p0 + 2);
                if (bstart >= tk.start && bstart < tk.end) {
                    snfb_lead L;
                    L.rec = rec; L.qname_hash = a.qh; L.read_len = 0; L.seq_off = -1; L.seq_len = 0;
                    L.ref_start = L.ref_end = bstart; L.qry_start = a.qas; L.qry_end = a.qae; L.svlen = 0;
                    L.mate_pos = (int)mate; L.mate_contig = contig_lookup_dev(P.contig, P.n_contig, e + e0.off[0], e0.len[0]);
                    if (L.mate_contig < 0) ++soft;
                    L.nm_sa = (int)sanm; L.task = (uint16_t)r.task; L.k = (uint16_t)(a.nlead + added);
                    L.flags = a.base_flags | SNFB_BND | ((uint32_t)SNFB_SRC_BND_SA << 3) | (is_first ? SNFB_LF_BND_FIRST : 0u) | (is_reverse ? SNFB_LF_BND_REVERSE : 0u)
                              | ((r.aux_flags & SNFB_AUX_NM) ? 0u : SNFB_LF_NM_NONE);
                    const unsigned long long slot = alloc_slot_lane(*a.slots, P.leads, P.lead_cap, a.n_slots);
                    if (slot < P.lead_cap) store_lead(P.leads + slot, L); else ++overflow;
                    ++added;
                }
            }
        }
    }
    // read_itersplits: primary alignments only
    if (!a.is_supp && sa_ok && ne > 0) {
        const double lim = __dadd_rn((double)cfg.max_splits_base, __dmul_rn(cfg.max_splits_kb, __ddiv_rn((double)r.l_seq, 1000.0)));
        if (!((double)ne > lim)) {
            if (ne + 1 > MAXSEG) ++soft;
            else {
                bool ok = true;
                sg[0].contig = tk.contig; sg[0].ref_start = r.pos; sg[0].ref_end = a.ref_end;
                sg[0].qry_start = rev ? r.l_seq - a.qae : a.qas; sg[0].qry_end = sg[0].qry_start + a.alen;
                sg[0].meta = (rev ? 1 : 0) | ((int)r.mapq << 8) | (SNFB_SRC_SPLIT_PRIM << 16); sg[0].nhint = 0;
                int ei = 0;
                for (int i = 0, st = 0; i <= sl && ok; ++i) if (i == sl || sa[i] == ';') {
                    if (i > st) {
                        SaEntry en; const uint8_t* e = sa + st; long long p1, cs, ce, rs, qs, mq;
                        if (!sa_fields_dev(e, i - st, &en) || !parse_int_dev(e + en.off[4], en.len[4], &mq)) { ok = false; ++soft; break; }
                        const bool srev = en.len[2] == 1 && e[en.off[2]] == '-';
                        if (!cigar_analyze_dev(e + en.off[3], en.len[3], &cs, &ce, &rs, &qs)) { ok = false; ++soft; break; }
                        if (!parse_int_dev(e + en.off[1], en.len[1], &p1)) { ok = false; ++soft; break; }
                        Seg* g = &sg[++ei];
                        g->contig = contig_lookup_dev(P.contig, P.n_contig, e + en.off[0], en.len[0]); if (g->contig < 0) { g->contig = -2 - ei; ++soft; }
                        g->ref_start = (int)(p1 - 1); g->ref_end = (int)(p1 - 1 + rs); g->qry_start = (int)(srev ? ce : cs); g->qry_end = g->qry_start + (int)qs;
                        g->meta = (srev ? 1 : 0) | ((int)mq << 8) | (SNFB_SRC_SPLIT_SUP << 16); g->nhint = 0;
                    }
                    st = i + 1;
                }
                if (ok) {
                    const int m = classify_splits_dev(cfg, sg, ne + 1, r.l_seq);
                    for (int i = 0; i < m; ++i) for (int h = 0; h < sg[i].nhint; ++h) {
                        const int mqc = (sg[i].meta >> 8) & 255, mqp = (sg[i > 0 ? i - 1 : 0].meta >> 8) & 255;
                        if (!cfg.dev_keep_lowqual_splits && (mqc < mqp ? mqc : mqp) < cfg.mapq) continue;
                        const int ty = sg[i].h_type[h]; const int hs = sg[i].h_start[h];
                        if (sg[i].contig != tk.contig || hs < tk.start || hs >= tk.end) continue;
                        snfb_lead L;
                        L.rec = rec; L.qname_hash = a.qh; L.read_len = 0; L.seq_off = -1; L.seq_len = 0; L.mate_contig = -1; L.mate_pos = 0; L.nm_sa = 0;
                        L.ref_start = hs; L.ref_end = (!sg[i].h_none[h] && ty != SNFB_INS) ? hs + sg[i].h_len[h] : hs;
                        L.qry_start = sg[i].qry_start; L.qry_end = sg[i].qry_end; L.svlen = sg[i].h_len[h];
                        uint32_t f = (uint32_t)ty | ((uint32_t)((sg[i].meta >> 16) & 3) << 3) | ((sg[i].meta & 1) ? SNFB_LF_REVERSE : 0u) | ((uint32_t)mqc << 16) | ((uint32_t)a.hp << 24);
                        if (sg[i].h_none[h]) f |= SNFB_LF_SVLEN_NONE;
                        if (ty == SNFB_INS && (sg[i].meta & (1 << 20))) { f |= SNFB_LF_HAS_SEQ; L.seq_off = sg[i].seq_off; L.seq_len = sg[i].seq_len; }
                        L.flags = f; L.task = (uint16_t)r.task; L.k = (uint16_t)(a.nlead + added);
                        const unsigned long long slot = alloc_slot_lane(*a.slots, P.leads, P.lead_cap, a.n_slots);
                        if (slot < P.lead_cap) store_lead(P.leads + slot, L); else ++overflow;
                        ++added;
                    }
                }
            }
        }
    }
    *soft_p += soft; *overflow_p += overflow;
    return added;
}

// ================================================================================================
// Stage A kernels:
//   k_rec_index (thread per record): task boundaries, sortedness, and everything that only needs the record core and the
//           clip ops at the two ends of its CIGAR: query_alignment_start/end, the read filters of iter_region, the
//           16-byte scan descriptor and the clip facts k_emit / k_sa use, the record's chunk count.
//   k_cdesc / k_chunk_sum (hot, HBM-bound) / k_rec_base / k_chunk_rare / k_rec_fin: the CIGAR walk, see "stage A streaming: chunks" below:
//           reference end, the NM correction, lead counts, and one 32-byte Event per SV signature.  No lead is built here.
//   k_rec_post (thread per record): nm per read (the division), per-task read count / covered bases / longest span.
//   k_emit  one thread per Event: the 64-byte lead.
//   k_sa    one thread per record with an SA tag: Lead.for_bnd + read_itersplits.
// ================================================================================================
// CIGAR16 (include/snfb.h): base word = [15] 0 | [14] E | [13:11] class | [10:0] length & 0x7ff; class bit 0 (word bit 11) = advances the
// read, class bit 1 (word bit 12) = advances the reference; E = an I / D / S of at least the block's event length (what the streaming
// kernel must look at).  Extension word = [15] 1 | [14:12] level (1, 2) | [11:0] payload, adding payload << (11 + 12 * (level - 1)).
constexpr unsigned C16_I = 1, C16_D = 2, C16_M = 3, C16_H = 4, C16_S = 5;
constexpr unsigned C16_LEN_BITS = 11, C16_LEN_MASK = 0x7ffu, C16_E = 0x4000u;
__device__ __forceinline__ bool c16_is_event(unsigned cls) { return (0x26u >> cls) & 1u; }        // I D S
__device__ __forceinline__ unsigned c16_class(unsigned w) { return (w >> C16_LEN_BITS) & 7u; }
__device__ __forceinline__ unsigned c16_ext_add(unsigned e) { return (e & 0xfffu) << (C16_LEN_BITS + 12u * (((e >> 12) & 7u) - 1u)); }

// the eight 16-bit words one lane holds -> up to eight ops at their word positions (cls 0 / len 0 where a pad, P or
// extension word sits).  Groups never straddle a 16-byte boundary, so this is lane-local.
__device__ __forceinline__ void c16_decode8(const uint32_t (&ww)[4], unsigned (&cls)[8], unsigned (&len)[8]) {
    unsigned x[8];
    #pragma unroll
    for (int h = 0; h < 8; ++h) x[h] = (ww[h >> 1] >> (16 * (h & 1))) & 0xffffu;
    #pragma unroll
    for (int h = 0; h < 8; ++h) {
        unsigned c = 0, l = 0;
        if (!(x[h] & 0x8000u)) {
            c = c16_class(x[h]); l = x[h] & C16_LEN_MASK;
            if (h + 1 < 8 && (x[h + 1 < 8 ? h + 1 : 7] & 0x8000u)) {
                l += c16_ext_add(x[h + 1 < 8 ? h + 1 : 7]);
                if (h + 2 < 8 && (x[h + 2 < 8 ? h + 2 : 7] & 0x8000u)) l += c16_ext_add(x[h + 2 < 8 ? h + 2 : 7]);
            }
        }
        cls[h] = c; len[h] = l;
    }
}

struct RecScan { uint32_t cig8; uint32_t n_words; int32_t pos; uint32_t meta; };    // what the CIGAR walk needs of a record, one 16-byte load
constexpr int CH = 16;                                                              // 16-byte groups per chunk of the streaming pass (k_chunk_sum)
__host__ __device__ __forceinline__ uint32_t chunks_of(uint32_t n_words) { return (((n_words + 7u) >> 3) + (uint32_t)CH - 1u) / (uint32_t)CH; }
struct RecClip { int32_t alen, qas, clip_left, clip_right; };                        // query_alignment_length/start, first / last op if it is a clip
constexpr uint32_t RM_PASS = 1u << 24, RM_HAS_NM = 1u << 25, RM_HAS_SA = 1u << 26;   // RecScan.meta: task (0..15) | mapq (16..23) | flags | hp (27..28)

struct IndexParams {
    const snfb_rec* rec; const uint16_t* cigar; const snfb_task* task; uint32_t n_rec; uint32_t n_task; unsigned long long n_cigar;
    int32_t* rec_pos; uint32_t* task_first; uint32_t* task_last;
    RecScan* scan; RecClip* clip; int32_t* rec_end; uint8_t* rec_flags; double* rec_nm; uint32_t* rec_nlead;
    uint32_t* pass_chunks;      // number of chunks (CH 16-byte CIGAR16 groups each) of a passing record, else 0 (scanned into the chunk table)
    DevCounters* ctr; int mapq_min, alen_min, excl, want_nm;
};
// leadprov.py:488-516 (filters), pysam query_alignment_start / query_alignment_end
__global__ void __launch_bounds__(256) k_rec_index(const IndexParams P) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x; if (i >= P.n_rec) return;
    const uint4* core = reinterpret_cast<const uint4*>(P.rec + i);
    const uint4 c0 = __ldg(core), c1 = __ldg(core + 1), c2 = __ldg(core + 2);
    const int task = (int)c0.x, pos = (int)c0.y; const unsigned flag = c0.z & 0xffffu, mapq = (c0.z >> 16) & 255u, aux = c0.z >> 24; unsigned hp = c0.w & 255u;
    const int nm = (int)c1.x; const uint32_t n = c1.z; const int l_seq = (int)c1.w;
    const unsigned long long cigar_off = (unsigned long long)c2.z | ((unsigned long long)c2.w << 32);
    if ((uint32_t)task >= P.n_task || (cigar_off & 7) || cigar_off + n > P.n_cigar) {      // malformed record (counted by k_validate: the run fails); touch nothing through its offsets
        RecScan s; s.cig8 = 0; s.n_words = 0; s.pos = pos; s.meta = 0; *reinterpret_cast<uint4*>(P.scan + i) = *reinterpret_cast<const uint4*>(&s);
        RecClip c; c.alen = 0; c.qas = 0; c.clip_left = 0; c.clip_right = 0; *reinterpret_cast<int4*>(P.clip + i) = *reinterpret_cast<const int4*>(&c);
        P.rec_pos[i] = pos; P.rec_flags[i] = 0; P.rec_nm[i] = -1.0; P.rec_end[i] = -1; P.rec_nlead[i] = 0; P.pass_chunks[i] = 0; return;
    }
    P.rec_pos[i] = pos;
    if (i == 0) P.task_first[task] = 0;
    else { const int2 pv = __ldg(reinterpret_cast<const int2*>(P.rec + i - 1)); if (pv.x != task) { P.task_first[task] = i; if ((uint32_t)pv.x < P.n_task) P.task_last[pv.x] = i; } else if (pv.y > pos) atomicAdd(&P.ctr->unsorted, 1ULL); }
    if (i + 1 == P.n_rec) P.task_last[task] = P.n_rec;
    // clips at the two ends
    const uint16_t* cg = P.cigar + cigar_off;
    int qas = 0, qae = l_seq, clip_left = 0, clip_right = 0; uint32_t fe = 0;
    { bool first = true; uint32_t k = 0;
      while (k < n) {
          const unsigned w = __ldg(cg + k); if (w == 0) { ++k; continue; }
          unsigned len = w & C16_LEN_MASK; const unsigned cls = c16_class(w); uint32_t k2 = k + 1;
          while (k2 < n) { const unsigned e = __ldg(cg + k2); if (!(e & 0x8000u)) break; len += c16_ext_add(e); ++k2; }
          if (first) { if (cls == C16_S || cls == C16_H) clip_left = (int)len; first = false; fe = k2; }
          if (cls == C16_S) qas += (int)len; else if (cls != C16_H) break;
          k = k2;
      } }
    { bool last = true; long k = (long)n - 1;
      while (k >= (long)fe) {
          long b = k; while (b > (long)fe && (__ldg(cg + b) & 0x8000u)) --b;
          const unsigned w = __ldg(cg + b); if (w == 0) { k = b - 1; continue; }
          unsigned len = w & C16_LEN_MASK; const unsigned cls = c16_class(w);
          for (long e2 = b + 1; e2 <= k; ++e2) { const unsigned e = __ldg(cg + e2); len += c16_ext_add(e); }
          if (last) { if (cls == C16_S || cls == C16_H) clip_right = (int)len; last = false; }
          if (cls == C16_S) qae -= (int)len; else if (cls != C16_H) break;
          k = b - 1;
      }
      if (last) clip_right = clip_left; }            // a single op is both the first and the last one
    const int alen = qae - qas;
    const snfb_task tk = P.task[task];
    const bool pass = !((int)mapq < P.mapq_min || (flag & 256u) || alen < P.alen_min) && !(P.excl && (flag & (unsigned)P.excl)) && pos >= tk.start && pos < tk.end && n > 0;
    const bool has_nm = pass && P.want_nm && (aux & SNFB_AUX_NM);
    if (!(aux & SNFB_AUX_HP)) hp = 0;
    if (pass && hp > 2) { hp = 0; atomicAdd(&P.ctr->soft_errors, 1ULL); }
    RecScan s; s.cig8 = (uint32_t)(cigar_off >> 3); s.n_words = n; s.pos = pos;
    s.meta = ((uint32_t)task & 0xffffu) | (mapq << 16) | (pass ? RM_PASS : 0u) | (has_nm ? RM_HAS_NM : 0u) | ((pass && (aux & SNFB_AUX_SA)) ? RM_HAS_SA : 0u) | ((pass ? hp : 0u) << 27);
    *reinterpret_cast<uint4*>(P.scan + i) = *reinterpret_cast<const uint4*>(&s);
    RecClip c; c.alen = alen; c.qas = qas; c.clip_left = clip_left; c.clip_right = clip_right;
    *reinterpret_cast<int4*>(P.clip + i) = *reinterpret_cast<const int4*>(&c);
    P.rec_flags[i] = pass ? (uint8_t)(RF_PASS | (has_nm ? RF_HAS_NM : 0) | (hp << 2)) : (uint8_t)0;
    P.rec_nm[i] = has_nm ? (double)nm : -1.0;        // k_rec_post turns it into (nm - big) / (alen + 1)
    P.rec_end[i] = -1; P.rec_nlead[i] = 0;
    P.pass_chunks[i] = pass ? chunks_of(n) : 0u;
}

// one SV signature found by k_chunk_rare (an I / D / S op of at least minsvlen_screen inside the task's region): all k_emit needs to build its lead
struct Event { uint32_t rec; uint32_t len; uint32_t pos_q; int32_t pos_r; uint32_t k_cls; uint32_t fidx; uint32_t pad[2]; };   // k_cls: ordinal inside its flagged chunk | class << 16; fidx: that chunk's place in the flagged list

// ---- stage A streaming: chunks ----------------------------------------------------------------------------------------------
// A passing record's CIGAR16 groups (16 bytes = 8 words) are cut into CHUNKS of up to CH groups (256 bytes); a chunk belongs to one
// record.  The work splits into a dense streaming pass and a sparse one:
//   k_chunk_sum   one THREAD per chunk walks its groups sequentially: per 32-bit word (two ops) two masked sums (read / reference
//                 advance, both halves at once) and one OR (E / extension flags).  No cross-lane traffic at all: the instruction count
//                 per byte is what the masked sums cost, and the kernel runs at memory speed.  Writes (read advance, reference advance,
//                 "has a flagged word") per chunk.
//   k_rec_base    one thread per record: exclusive prefix over its chunks -> absolute (query, reference) position at every chunk start,
//                 reference_end of the record; the record's flagged chunks are appended to a dense list (contiguous per record, in order).
//   k_chunk_rare  one thread per FLAGGED chunk (0.4 % of the groups hold an I / D / S of event length or an extension word; ~6 % of
//                 the chunks): walks the chunk again with its base positions, decodes the flagged groups, applies the region test
//                 (leadprov.py:464-466) and appends one Event per SV signature with its ordinal inside the chunk.
//   k_rec_fin     one thread per record: prefix over its flagged chunks' event counts -> ordinal base per flagged chunk, lead count
//                 and "big indel" sum of the record (get_cigar_indels), SA list.  k_emit adds the base to an event's ordinal.
struct ChunkParams {
    const RecScan* scan; const uint32_t* choff; uint32_t n_rec;          // choff: exclusive scan of the records' chunk counts
    const uint16_t* cigar; const snfb_task* task;
    uint2* cdesc;            // per chunk: first group (index of the 16-byte group in the arena), record << 4 | groups - 1
    uint2* csum;             // per chunk: k_chunk_sum (read advance | flag << 31, reference advance) -> k_rec_base (query pos | flag << 31, reference pos) at the chunk start
    uint32_t* flist;         // flagged chunks, contiguous per record
    uint2* fcnt;             // per flagged chunk: (events, big) -> k_rec_fin: (ordinal base, -)
    uint32_t* rec_foff; uint32_t* rec_nf;
    const unsigned long long* n_chunks; unsigned long long* n_flag; unsigned long long chunk_cap;
    int32_t* rec_end; uint32_t* rec_nlead; int32_t* rec_big;
    Event* ev; unsigned long long ev_cap; unsigned long long* n_ev;
    uint32_t* sa_list; unsigned long long* n_sa;
    DevCounters* ctr;
    int minsv;
};

// thread per record: chunk descriptors of a passing record
__global__ void __launch_bounds__(256) k_cdesc(const ChunkParams P) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x; if (i >= P.n_rec) return;
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(P.scan + i));
    if (!(d.w & RM_PASS)) return;
    const uint32_t g = (d.y + 7u) >> 3, c0 = P.choff[i];
    for (uint32_t j = 0, left = g; left; ++j) {
        const uint32_t n = left < (uint32_t)CH ? left : (uint32_t)CH;
        if ((unsigned long long)c0 + j < P.chunk_cap) P.cdesc[c0 + j] = make_uint2(d.x + j * (uint32_t)CH, (i << 4) | (n - 1u));
        left -= n;
    }
}
// lowers the E-bit threshold of a CIGAR16 arena in place (a config that cares about shorter events than the block was packed for)
__global__ void k_reflag(uint16_t* __restrict__ cigar, unsigned long long n_words, unsigned evt_min) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_words; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned w = cigar[i]; if (w & 0x8000u) continue;
        const unsigned cls = c16_class(w);
        if (!c16_is_event(cls) || (w & C16_E)) continue;
        bool big = (w & C16_LEN_MASK) >= evt_min;
        if (!big && (i & 7) != 7) big = (cigar[i + 1] & 0x8000u) != 0;          // an extension word follows: the length is at least 2048
        if (big) cigar[i] = (uint16_t)(w | C16_E);
    }
}

// read / reference advance of a lane's eight words when one of them is an extension word
__device__ __noinline__ uint2 lane_sums_ext(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    const uint32_t ww[4] = { w0, w1, w2, w3 };
    unsigned cls[8], len[8]; c16_decode8(ww, cls, len);
    unsigned lq = 0, lr = 0;
    #pragma unroll
    for (int j = 0; j < 8; ++j) { lq += len[j] * (cls[j] & 1u); lr += len[j] * ((cls[j] >> 1) & 1u); }
    return make_uint2(lq, lr);
}

// per 32-bit word (two ops): the halves' lengths masked by "advances the read" (class bit 0 = word bit 11) / "advances the reference" (bit 12)
#define SNFB_WORD_BODY(w, aq, ar) { const uint32_t t_ = (w) >> 11; aq += (w) & ((t_ & 0x00010001u) * 0x7ffu); ar += (w) & (((t_ >> 1) & 0x00010001u) * 0x7ffu); }

// The streaming kernel: one thread per chunk.  A thread issues the loads of 8 groups (128 bytes) before it touches the first of them,
// so a warp keeps 4 KB in flight; groups behind the end of a short chunk read as pad words (all-zero words advance nothing).  The packed
// half-word sums hold 32 values of at most 2047 before they could carry into the neighbouring half: they are folded every 8 groups.
__global__ void __launch_bounds__(256, 4) k_chunk_sum(const __grid_constant__ ChunkParams P) {
    unsigned long long n = *P.n_chunks; if (n > P.chunk_cap) n = P.chunk_cap;
    const uint4* __restrict__ cig4 = reinterpret_cast<const uint4*>(P.cigar);
    for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 d = __ldg(P.cdesc + c);
        const int ng = (int)(d.y & 15u) + 1;
        const uint4* src = cig4 + d.x;
        uint32_t q = 0, r = 0, flags = 0;
        #pragma unroll 1
        for (int g0 = 0; g0 < ng; g0 += 8) {
            uint4 v[8];
            #pragma unroll
            for (int g = 0; g < 8; ++g) v[g] = g0 + g < ng ? __ldg(src + g0 + g) : make_uint4(0u, 0u, 0u, 0u);
            uint32_t aq = 0, ar = 0;
            #pragma unroll
            for (int g = 0; g < 8; ++g) {
                const uint32_t rb = (v[g].x | v[g].y | v[g].z | v[g].w) & 0xC000C000u;
                flags |= rb;
                if (rb & 0x80008000u) { const uint2 t = lane_sums_ext(v[g].x, v[g].y, v[g].z, v[g].w); q += t.x; r += t.y; }      // an extension word: the group is decoded op by op
                else { SNFB_WORD_BODY(v[g].x, aq, ar) SNFB_WORD_BODY(v[g].y, aq, ar) SNFB_WORD_BODY(v[g].z, aq, ar) SNFB_WORD_BODY(v[g].w, aq, ar) }
            }
            q += (aq & 0xffffu) + (aq >> 16); r += (ar & 0xffffu) + (ar >> 16);
        }
        P.csum[c] = make_uint2(q | (flags ? 0x80000000u : 0u), r);
    }
}

// thread per record: positions at every chunk start, reference_end, the record's flagged chunks (slots reserved once per block)
__global__ void __launch_bounds__(256) k_rec_base(const ChunkParams P) {
    __shared__ unsigned long long s_base;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    uint32_t c0 = 0, nch = 0, nf = 0; bool pass = false;
    if (i < P.n_rec) {
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(P.scan + i));
        if (d.w & RM_PASS) {
            pass = true; c0 = P.choff[i]; nch = chunks_of(d.y);
            if ((unsigned long long)c0 + nch > P.chunk_cap) nch = (unsigned long long)c0 < P.chunk_cap ? (uint32_t)(P.chunk_cap - c0) : 0u;
            uint32_t q = 0; int r = (int)d.z;
            for (uint32_t j = 0; j < nch; ++j) {
                const uint2 s = P.csum[c0 + j];
                P.csum[c0 + j] = make_uint2(q | (s.x & 0x80000000u), (uint32_t)r);
                q += s.x & 0x7fffffffu; r += (int)s.y; nf += s.x >> 31;
            }
            P.rec_end[i] = r;
        }
    }
    uint32_t tot; const uint32_t mine = prims::block_excl_scan(nf, &tot);
    if (threadIdx.x == 0) s_base = tot ? atomicAdd(P.n_flag, (unsigned long long)tot) : 0ull;
    __syncthreads();
    if (i < P.n_rec) {
        const uint32_t f0 = (uint32_t)s_base + mine;
        if (nf) { uint32_t k = 0; for (uint32_t j = 0; j < nch; ++j) if (P.csum[c0 + j].x >> 31) { if ((unsigned long long)f0 + k < P.chunk_cap) P.flist[f0 + k] = c0 + j; ++k; } }
        P.rec_foff[i] = f0; P.rec_nf[i] = pass ? nf : 0u;
    }
}

// thread per flagged chunk: SV signatures (read_iterindels, leadprov.py:583-670) and the indels get_cigar_indels counts (leadprov.py:198-224).
// The flagged groups are decoded twice — to count the chunk's signatures, then, with event slots reserved once per block, to write them.
__global__ void __launch_bounds__(128) k_chunk_rare(const __grid_constant__ ChunkParams P) {
    __shared__ unsigned long long s_base;
    unsigned long long n = *P.n_flag; if (n > P.chunk_cap) n = P.chunk_cap;
    const uint4* __restrict__ cig4 = reinterpret_cast<const uint4*>(P.cigar);
    for (unsigned long long f0 = (unsigned long long)blockIdx.x * 128; f0 < n; f0 += (unsigned long long)gridDim.x * 128) {
        const unsigned long long f = f0 + threadIdx.x; const bool valid = f < n;
        uint32_t pq[CH]; int pr[CH]; unsigned mask = 0, cnt = 0, big = 0; uint32_t rec = 0; const uint4* src = cig4; int tk_start = 0, tk_end = 0;
        if (valid) {
            const uint32_t c = P.flist[f];
            const uint2 d = __ldg(P.cdesc + c), base = P.csum[c];
            const int ng = (int)(d.y & 15u) + 1; rec = d.y >> 4; src = cig4 + d.x;
            const snfb_task tk = P.task[P.scan[rec].meta & 0xffffu]; tk_start = tk.start; tk_end = tk.end;
            // positions in front of every group, which groups hold a flagged word
            uint32_t q = base.x & 0x7fffffffu; int r = (int)base.y;
            #pragma unroll 1
            for (int g = 0; g < ng; ++g) {
                const uint4 v = __ldg(src + g);
                pq[g] = q; pr[g] = r;
                const uint32_t rb = (v.x | v.y | v.z | v.w) & 0xC000C000u;
                if (rb) mask |= 1u << g;
                if (rb & 0x80008000u) { const uint2 t = lane_sums_ext(v.x, v.y, v.z, v.w); q += t.x; r += (int)t.y; }
                else { uint32_t aq = 0, ar = 0; SNFB_WORD_BODY(v.x, aq, ar) SNFB_WORD_BODY(v.y, aq, ar) SNFB_WORD_BODY(v.z, aq, ar) SNFB_WORD_BODY(v.w, aq, ar)
                       q += (aq & 0xffffu) + (aq >> 16); r += (int)((ar & 0xffffu) + (ar >> 16)); }
            }
        }
        unsigned long long e = 0;
        #pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            unsigned k = 0, m = mask;
            while (m) {
                const int g = __ffs(m) - 1; m &= m - 1u;
                const uint4 v = __ldg(src + g);
                const uint32_t ww[4] = { v.x, v.y, v.z, v.w };
                unsigned cls[8], len[8]; c16_decode8(ww, cls, len);
                uint32_t q2 = pq[g]; int r2 = pr[g];
                #pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (pass == 0 && len[j] > 10u && (cls[j] == C16_I || cls[j] == C16_D)) big += len[j];            // get_cigar_indels, minoplen 10
                    if (c16_is_event(cls[j]) && (int)len[j] >= P.minsv) {
                        const int rs = cls[j] == C16_D ? r2 + (int)len[j] : r2;
                        if (rs >= tk_start && rs < tk_end) {                                       // the signature stays inside the task's region (leadprov.py:464-466)
                            if (pass == 1) {
                                if (e + k < P.ev_cap) { uint4* dst = reinterpret_cast<uint4*>(P.ev + e + k); dst[0] = make_uint4(rec, len[j], q2, (uint32_t)r2); dst[1] = make_uint4((k & 0xffffu) | (cls[j] << 16), (uint32_t)f, 0u, 0u); }
                                else atomicAdd(&P.ctr->lead_overflow, 1ULL);
                            }
                            ++k;
                        }
                    }
                    q2 += len[j] * (cls[j] & 1u); r2 += (int)(len[j] * ((cls[j] >> 1) & 1u));
                }
            }
            if (pass == 0) {
                cnt = k;
                // one reservation of event slots per block
                __shared__ uint32_t wsum[4];
                uint32_t inc = prims::warp_incl_scan(cnt);
                if (lane_id() == 31) wsum[threadIdx.x >> 5] = inc;
                __syncthreads();
                const uint32_t w = threadIdx.x >> 5; uint32_t before = 0, tot = 0;
                #pragma unroll
                for (int x = 0; x < 4; ++x) { const uint32_t t = wsum[x]; if ((uint32_t)x < w) before += t; tot += t; }
                if (threadIdx.x == 0) s_base = tot ? atomicAdd(P.n_ev, (unsigned long long)tot) : 0ull;
                __syncthreads();
                e = s_base + before + inc - cnt;
                __syncthreads();
            }
        }
        if (valid) P.fcnt[f] = make_uint2(cnt, big);
    }
}

// thread per record: ordinal base of every flagged chunk, lead count and big-indel sum of the record, the SA work list (slots reserved once per block)
__global__ void __launch_bounds__(256) k_rec_fin(const ChunkParams P) {
    __shared__ unsigned long long s_base;
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    uint32_t has_sa = 0;
    if (i < P.n_rec) {
        const uint32_t meta = P.scan[i].meta;
        if (meta & RM_PASS) {
            const uint32_t f0 = P.rec_foff[i]; uint32_t nf = P.rec_nf[i];
            if ((unsigned long long)f0 + nf > P.chunk_cap) nf = (unsigned long long)f0 < P.chunk_cap ? (uint32_t)(P.chunk_cap - f0) : 0u;
            uint32_t k = 0, big = 0;
            for (uint32_t j = 0; j < nf; ++j) { const uint2 e = P.fcnt[f0 + j]; P.fcnt[f0 + j] = make_uint2(k, 0u); k += e.x; big += e.y; }
            if (k > 0xffffu) atomicAdd(&P.ctr->ordinal_overflow, 1ULL);
            P.rec_nlead[i] = k; P.rec_big[i] = (int)big;
            has_sa = (meta & RM_HAS_SA) ? 1u : 0u;
        }
    }
    uint32_t tot; const uint32_t mine = prims::block_excl_scan(has_sa, &tot);
    if (threadIdx.x == 0) s_base = tot ? atomicAdd(P.n_sa, (unsigned long long)tot) : 0ull;
    __syncthreads();
    if (has_sa) P.sa_list[s_base + mine] = i;
}

// per read nm (leadprov.py:517-526) and the per-task bookkeeping of iter_region (read count, covered bases, longest span)
struct PostParams {
    const RecScan* scan; const RecClip* clip; const snfb_task* task; uint32_t n_rec;
    const int32_t* rec_end; const int32_t* rec_big; double* rec_nm;
    uint32_t* task_reads; unsigned long long* task_cov_bp; int32_t* task_maxspan;
};
__global__ void __launch_bounds__(256) k_rec_post(const PostParams P) {
    __shared__ unsigned s_reads; __shared__ unsigned long long s_bp; __shared__ int s_span; __shared__ int s_task;
    const uint32_t i0 = blockIdx.x * 256, i = i0 + threadIdx.x;
    if (threadIdx.x == 0) { s_reads = 0; s_bp = 0; s_span = 0; s_task = (int)(P.scan[i0].meta & 0xffffu); }
    __syncthreads();
    unsigned mine = 0; unsigned long long bp = 0; int span = 0;
    if (i < P.n_rec) {
        const RecScan s = P.scan[i];
        if (s.meta & RM_PASS) {
            const int task = (int)(s.meta & 0xffffu), ref_end = P.rec_end[i];
            if (s.meta & RM_HAS_NM) P.rec_nm[i] = __ddiv_rn(P.rec_nm[i] - (double)P.rec_big[i], (double)(P.clip[i].alen + 1));
            const int tk_len = P.task[task].contig_len;
            const int ce = ref_end < tk_len ? ref_end : tk_len;
            bp = ce > s.pos ? (unsigned long long)(ce - s.pos) : 0ull; span = ref_end - s.pos; if (span < 0) span = 0;
            if (task == s_task) mine = 1;
            else { atomicAdd(&P.task_reads[task], 1u); atomicAdd(&P.task_cov_bp[task], bp); atomicMax(&P.task_maxspan[task], span); bp = 0; span = 0; }   // block straddles a task boundary
        }
    }
    const unsigned wr = __reduce_add_sync(FULL, mine); const int ws = __reduce_max_sync(FULL, span);
    #pragma unroll
    for (int o = 16; o; o >>= 1) bp += __shfl_xor_sync(FULL, bp, o);
    if (lane_id() == 0 && wr) { atomicAdd(&s_reads, wr); atomicAdd(&s_bp, bp); atomicMax(&s_span, ws); }
    __syncthreads();
    if (threadIdx.x == 0 && s_reads) { atomicAdd(&P.task_reads[s_task], s_reads); atomicAdd(&P.task_cov_bp[s_task], s_bp); atomicMax(&P.task_maxspan[s_task], s_span); }
}

struct EmitParams {
    const snfb_rec* rec; const RecClip* clip; const uint8_t* var;
    const Event* ev; const unsigned long long* n_ev; unsigned long long ev_cap; const uint2* fcnt;      // fcnt[fidx].x: ordinal base of the event's chunk inside its record
    snfb_lead* leads; DevCounters* ctr;
    int maxlen, detect_large_ins; double longinslen;
};
// qname_hash_warp (common.cuh) evaluated by one thread
__device__ inline uint64_t qname_hash_thread(const uint8_t* s, int n) {
    uint64_t acc = 0;
    for (int l = 0; l * 8 < n; ++l) {
        uint64_t w = 0; const int m = n - l * 8 < 8 ? n - l * 8 : 8;
        for (int j = 0; j < m; ++j) w |= (uint64_t)s[l * 8 + j] << (8 * j);
        acc += qname_word(w, (uint32_t)l);
    }
    return qname_finish(acc + (0x9E3779B97F4A7C15ull ^ (uint64_t)n));
}
// one thread per event: read_iterindels' lead construction (leadprov.py:583-670).  Event e becomes lead slot e; the slots of
// the SA leads (k_sa) start behind the last event.
__global__ void __launch_bounds__(128) k_emit(const EmitParams P) {
    const unsigned long long n_all = *P.n_ev, n = n_all < P.ev_cap ? n_all : P.ev_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) P.ctr->n_slots = n_all;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x) {
        const uint4 e0 = __ldg(reinterpret_cast<const uint4*>(P.ev + e)); const uint2 kf = __ldg(reinterpret_cast<const uint2*>(P.ev + e) + 2); const uint32_t kc = kf.x;
        const uint32_t rec = e0.x; const int ln = (int)e0.y, pqi = (int)e0.z, pr = (int)e0.w; const unsigned op = kc >> 16;
        const uint4* core = reinterpret_cast<const uint4*>(P.rec + rec);
        const uint4 c0 = __ldg(core), c3 = __ldg(core + 3);
        const int r_task = (int)c0.x, r_pos = (int)c0.y; const unsigned flag = c0.z & 0xffffu, mapq = (c0.z >> 16) & 255u, aux = c0.z >> 24; const int l_qname = (int)((c0.w >> 8) & 255u);
        const unsigned long long var_off = (unsigned long long)c3.z | ((unsigned long long)c3.w << 32);
        const int alen = P.clip[rec].alen;
        const bool is_supp = flag & 2048u, rev = flag & 16u, has_sa = aux & SNFB_AUX_SA;
        unsigned hp = (aux & SNFB_AUX_HP) ? (c0.w & 255u) : 0u; if (hp > 2) hp = 0;
        const bool use_clips = P.detect_large_ins && !is_supp && !has_sa;
        snfb_lead L;
        L.rec = rec; L.qname_hash = qname_hash_thread(P.var + var_off, l_qname); L.read_len = alen; L.seq_off = -1; L.seq_len = 0; L.mate_contig = -1; L.mate_pos = 0; L.nm_sa = 0;
        L.task = (uint16_t)r_task; L.k = (uint16_t)(((kc & 0xffffu) + P.fcnt[kf.y].x) & 0xffffu);
        uint32_t f = (rev ? SNFB_LF_REVERSE : 0u) | (mapq << 16) | ((uint32_t)SNFB_SRC_INLINE << 3) | (hp << 24) | (is_supp ? SNFB_LF_IS_SA : 0u);
        if (op == C16_I) { f |= SNFB_INS; L.ref_start = pr; L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi + ln; L.svlen = ln;
            if (ln <= P.maxlen) { f |= SNFB_LF_HAS_SEQ; L.seq_off = pqi; L.seq_len = ln; } }
        else if (op == C16_D) { f |= SNFB_DEL; L.ref_start = pr + ln; L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi; L.svlen = -ln; }
        else if (use_clips && (double)ln >= P.longinslen) { f |= SNFB_INS | SNFB_LF_SVLEN_NONE; L.ref_start = L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi + ln; L.svlen = 0; }
        else { f |= (pr == r_pos) ? SNFB_SINGLE_LEFT : SNFB_SINGLE_RIGHT; L.ref_start = L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi + ln; L.svlen = 0; }
        L.flags = f;
        store_lead(P.leads + e, L);
    }
}

struct SaParams {
    const snfb_rec* rec; const RecClip* clip; const uint8_t* var; const snfb_task* task; const snfb_contig* contig; uint32_t n_contig;
    const uint32_t* sa_list; const unsigned long long* n_sa; const int32_t* rec_end; uint32_t* rec_nlead;
    snfb_lead* leads; unsigned long long lead_cap; DevCounters* ctr;
    Seg* seg_scratch;                 // MAXSEG segments per thread of the grid
    snfb_config cfg;
};
constexpr int SA_THREADS = 128, SA_BLOCKS = 148 * 8;
// one thread per record with an SA tag: the text parse is serial per record, so the parallelism is across records.
// Segments live in a per-thread slice of global scratch (a read has a handful of them; the touched part stays in L2).
__global__ void __launch_bounds__(SA_THREADS) k_sa(const SaParams P) {
    __shared__ snfb_config s_cfg;
    if (threadIdx.x < sizeof(snfb_config) / 4) reinterpret_cast<uint32_t*>(&s_cfg)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&P.cfg)[threadIdx.x];
    __syncthreads();
    const unsigned long long n = *P.n_sa;
    const unsigned long long tid = (unsigned long long)blockIdx.x * SA_THREADS + threadIdx.x, nthr = (unsigned long long)gridDim.x * SA_THREADS;
    Seg* sg = P.seg_scratch + tid * MAXSEG;
    unsigned long long soft = 0, overflow = 0;
    SlotState slots; slots.cur = 0; slots.end = 0;
    for (unsigned long long e = tid; e < n; e += nthr) {
        const uint32_t rec = P.sa_list[e];
        const uint4* core = reinterpret_cast<const uint4*>(P.rec + rec);
        const uint4 c0 = __ldg(core), c1 = __ldg(core + 1), c3 = __ldg(core + 3);
        const int r_task = (int)c0.x, r_pos = (int)c0.y; const unsigned flag = c0.z & 0xffffu, mapq = (c0.z >> 16) & 255u, aux = c0.z >> 24; const int l_qname = (int)((c0.w >> 8) & 255u);
        const int l_seq = (int)c1.w; const unsigned long long var_off = (unsigned long long)c3.z | ((unsigned long long)c3.w << 32);
        const uint32_t sa_len = __ldg(reinterpret_cast<const uint32_t*>(core + 2));
        const snfb_task tk = P.task[r_task];
        const RecClip rc = P.clip[rec];
        int hp = (aux & SNFB_AUX_HP) ? (int)(c0.w & 255u) : 0; if (hp > 2) hp = 0;
        const bool rev = flag & 16u;
        SaArgs a; a.rec = rec; a.qas = rc.qas; a.qae = rc.qas + rc.alen; a.alen = rc.alen; a.ref_end = P.rec_end[rec]; a.hp = hp;
        a.base_flags = (rev ? SNFB_LF_REVERSE : 0u) | (mapq << 16); a.qh = qname_hash_thread(P.var + var_off, l_qname); a.nlead = P.rec_nlead[rec]; a.rev = rev; a.is_supp = flag & 2048u;
        a.sa = P.var + var_off + l_qname; a.sa_len = (int)sa_len; a.clip_left = rc.clip_left; a.clip_right = rc.clip_right; a.pos = r_pos; a.l_seq = l_seq; a.mapq = (int)mapq;
        a.aux_flags = (int)aux; a.task = r_task; a.tk_contig = tk.contig; a.tk_start = tk.start; a.tk_end = tk.end; a.contig = P.contig; a.n_contig = P.n_contig;
        a.leads = P.leads; a.lead_cap = P.lead_cap; a.n_slots = &P.ctr->n_slots; a.slots = &slots;
        a.mapq_min = s_cfg.mapq; a.dev_keep_lowqual_splits = s_cfg.dev_keep_lowqual_splits; a.max_splits_base = s_cfg.max_splits_base; a.max_splits_kb = s_cfg.max_splits_kb;
        const unsigned added = process_sa(&s_cfg, sg, a, &soft, &overflow);
        if (added) { if (a.nlead + added > 0xffffu) atomicAdd(&P.ctr->ordinal_overflow, 1ULL); P.rec_nlead[rec] += added; }
    }
    for (unsigned long long sidx = slots.cur; sidx < slots.end; ++sidx) if (sidx < P.lead_cap) P.leads[sidx].rec = HOLE;   // retire the last chunk
    if (soft) atomicAdd(&P.ctr->soft_errors, soft);
    if (overflow) atomicAdd(&P.ctr->lead_overflow, overflow);
}

// deterministic per-task mean of the per-read nm values (config.average_regional_nm, leadprov.py:577).
// Fixed two-level reduction tree (4096-record chunks of each task, then the chunk partials in order): the result does
// not depend on scheduling; it can differ from the reference's sequential float accumulation in the last bits.
constexpr int NM_CHUNK = 4096;
__global__ void __launch_bounds__(256) k_nm_partial(const uint8_t* __restrict__ rec_flags, const double* __restrict__ rec_nm, const uint32_t* __restrict__ task_first,
                                                    const uint32_t* __restrict__ task_last, double* __restrict__ part_sum, unsigned* __restrict__ part_cnt) {
    __shared__ double ssum[256]; __shared__ unsigned scnt[256];
    const int t = blockIdx.y; const uint32_t lo = task_first[t] + blockIdx.x * NM_CHUNK, end = task_last[t];
    if (lo >= end) return;                                  // whole block exits together
    const uint32_t hi = lo + NM_CHUNK < end ? lo + NM_CHUNK : end;
    double s = 0; unsigned c = 0;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) if ((rec_flags[i] & (RF_PASS | RF_HAS_NM)) == (RF_PASS | RF_HAS_NM)) { s += rec_nm[i]; ++c; }
    ssum[threadIdx.x] = s; scnt[threadIdx.x] = c; __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; } __syncthreads(); }
    if (threadIdx.x == 0) { part_sum[(size_t)t * gridDim.x + blockIdx.x] = ssum[0]; part_cnt[(size_t)t * gridDim.x + blockIdx.x] = scnt[0]; }
}
__global__ void __launch_bounds__(256) k_task_nm(const uint32_t* __restrict__ task_first, const uint32_t* __restrict__ task_last, const double* __restrict__ part_sum,
                                                 const unsigned* __restrict__ part_cnt, int chunks_per_task, double* __restrict__ task_mean_nm) {
    __shared__ double ssum[256]; __shared__ unsigned long long scnt[256];
    const int t = blockIdx.x; const uint32_t lo = task_first[t], hi = task_last[t];
    const uint32_t nch = hi > lo ? (hi - lo + NM_CHUNK - 1) / NM_CHUNK : 0;
    double s = 0; unsigned long long c = 0;
    for (uint32_t ch = threadIdx.x; ch < nch; ch += 256) { s += part_sum[(size_t)t * chunks_per_task + ch]; c += part_cnt[(size_t)t * chunks_per_task + ch]; }
    ssum[threadIdx.x] = s; scnt[threadIdx.x] = c; __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) { ssum[threadIdx.x] += ssum[threadIdx.x + o]; scnt[threadIdx.x] += scnt[threadIdx.x + o]; } __syncthreads(); }
    if (threadIdx.x == 0) task_mean_nm[t] = ssum[0] / (double)(scnt[0] > 1 ? scnt[0] : 1);
}

}  // namespace extract
