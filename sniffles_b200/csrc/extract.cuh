// extract.cuh — stage A: alignment records -> SV leads.
//
// One warp per alignment record.  The warp streams the record's CIGAR with 16-byte loads
// (128 ops per iteration, coalesced), turns op lengths into running read/reference
// positions with a shuffle scan, and appends a 64-byte lead for every signature it meets.
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

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int MAXSEG = 40;     // primary + supplementary segments per read held in shared memory

struct Params {
    const snfb_rec* rec; const uint32_t* cigar; const uint8_t* var;
    const snfb_task* task; const snfb_contig* contig;
    uint32_t n_rec, n_task, n_contig;
    snfb_lead* leads; unsigned long long lead_cap;
    // per-record outputs
    int32_t* rec_pos; int32_t* rec_end; uint8_t* rec_flags; double* rec_nm; uint32_t* rec_nlead;
    // per-task outputs
    uint32_t* task_first; uint32_t* task_last; uint32_t* task_reads; unsigned long long* task_cov_bp; int32_t* task_maxspan;
    DevCounters* ctr;
    snfb_config cfg;
};

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

// ---- lead slot allocation.  A single global counter bumped once per lead serialises in L2; instead every warp reserves
//      SLOT_CHUNK slots at a time and hands them out locally.  Unused slots of a retired chunk are marked as holes
//      (rec == HOLE) and skipped later; canonical (record, k) order never depended on slot numbers.
constexpr unsigned SLOT_CHUNK = 64;
constexpr uint32_t HOLE = 0xffffffffu;
struct SlotState { unsigned long long cur, end; };
__device__ __forceinline__ unsigned long long alloc_slots_warp(SlotState& st, unsigned m, snfb_lead* leads, unsigned long long lead_cap, unsigned long long* n_slots) {
    if (st.cur + m > st.end) {          // warp-uniform
        for (unsigned long long s = st.cur + lane_id(); s < st.end; s += 32) if (s < lead_cap) leads[s].rec = HOLE;
        const unsigned chunk = m > SLOT_CHUNK ? m : SLOT_CHUNK;
        unsigned long long base = 0; if (lane_id() == 0) base = atomicAdd(n_slots, (unsigned long long)chunk);
        base = __shfl_sync(FULL, base, 0);
        st.cur = base; st.end = base + chunk;
    }
    const unsigned long long r = st.cur; st.cur += m; return r;
}
__device__ __forceinline__ unsigned long long alloc_slot_lane(SlotState& st, snfb_lead* leads, unsigned long long lead_cap, unsigned long long* n_slots) {   // one lane only
    if (st.cur + 1 > st.end) { const unsigned long long base = atomicAdd(n_slots, (unsigned long long)SLOT_CHUNK); st.cur = base; st.end = base + SLOT_CHUNK; }
    return st.cur++;
}

// ---- supplementary alignments (SA tag) of one record: Lead.for_bnd + read_itersplits, run by lane 0 ----
struct SaArgs {
    uint32_t rec; int qas, qae, alen, ref_end, hp; uint32_t base_flags; uint64_t qh; unsigned nlead; bool rev, is_supp;
    // the pieces of Params / snfb_rec / snfb_task the SA path needs, by value (keeps the caller's structs out of local memory)
    const uint8_t* sa; int sa_len; uint32_t c_first, c_last; int pos, l_seq, mapq, aux_flags, task;
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
        int left = 0, right = 0;
        { uint32_t c = a.c_first; int op = c & 15; if (op == 4 || op == 5) left = (int)(c >> 4); }
        { uint32_t c = a.c_last; int op = c & 15; if (op == 4 || op == 5) right = (int)(c >> 4); }
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

// record index: task boundaries, start positions and the coordinate-order check (one thread per record)
__global__ void __launch_bounds__(256) k_rec_index(const snfb_rec* __restrict__ rec, uint32_t n_rec, int32_t* __restrict__ rec_pos, uint32_t* __restrict__ task_first,
                                                   uint32_t* __restrict__ task_last, DevCounters* ctr) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x; if (i >= n_rec) return;
    const int2 tp = __ldg(reinterpret_cast<const int2*>(rec + i));            // (task, pos)
    rec_pos[i] = tp.y;
    if (i == 0) task_first[tp.x] = 0;
    else { const int2 pv = __ldg(reinterpret_cast<const int2*>(rec + i - 1)); if (pv.x != tp.x) { task_first[tp.x] = i; task_last[pv.x] = i; } else if (pv.y > tp.y) atomicAdd(&ctr->unsorted, 1ULL); }
    if (i + 1 == n_rec) task_last[tp.x] = n_rec;
}

// per-op flags for ops 0..7 (M I D N S H P =): bit0 advances the read, bit1 advances the reference
__device__ __forceinline__ unsigned op_flags(unsigned op) {
    const unsigned b = __byte_perm(0x02020103u, 0x03000001u, op) & 3u;
    return op == 8u ? 3u : b;                      // X (only with --eqx) behaves like M
}

// ================================================================================================
// Stage A is three kernels:
//   k_scan  (hot, HBM-bound): one warp per record streams the CIGAR once: filters, reference end, the NM correction,
//           coverage bookkeeping — and for every 128-op slice that holds an SV signature just one 32-byte "event slice"
//           entry.  No lead is built here, which keeps the streaming loop small (registers, I-cache).
//   k_emit  one warp per event slice: reload the slice (L2), prefix positions, write the 64-byte leads to exact slots.
//   k_sa    one warp per record with an SA tag: Lead.for_bnd + read_itersplits (lane-serial text parsing).
// ================================================================================================
struct EvSlice { uint32_t rec; int32_t base; uint32_t pos_q; int32_t pos_r; uint32_t k0; uint32_t count; uint32_t slot0; uint32_t pad; };

struct ScanParams {
    const snfb_rec* rec; const uint32_t* cigar; const snfb_task* task;
    uint32_t n_rec;
    int32_t* rec_end; uint8_t* rec_flags; double* rec_nm; uint32_t* rec_nlead;
    uint32_t* task_reads; unsigned long long* task_cov_bp; int32_t* task_maxspan;
    EvSlice* ev; unsigned long long ev_cap; unsigned long long* n_ev;
    uint32_t* sa_list; unsigned long long* n_sa;
    DevCounters* ctr;
    int minsv, mapq_min, alen_min, excl, want_nm;
};

// rare path of k_scan: a slice with at least one op longer than 10 that is not a match.  Returns (big << 32) | leads counted.
__device__ __noinline__ unsigned long long scan_rare(const ScanParams* __restrict__ P, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, unsigned lq, unsigned lr,
                                                      uint32_t rec, int base, unsigned pos_q, int pos_r, unsigned k0, int tk_start, int tk_end) {
    const int lane = lane_id();
    const uint32_t ww[4] = { w0, w1, w2, w3 };
    unsigned big = 0, evm = 0;
    #pragma unroll
    for (int j = 0; j < 4; ++j) { const unsigned op = ww[j] & 15u, len = ww[j] >> 4;
        if (len > 10u && (op == 1u || op == 2u)) big += len;                      // get_cigar_indels, minoplen 10
        if (((0x016u >> op) & 1u) && (int)len >= P->minsv) evm |= 1u << j; }
    big = __reduce_add_sync(FULL, big);
    unsigned count = 0;
    if (__any_sync(FULL, evm != 0)) {
        unsigned ir = lr;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned tr = __shfl_up_sync(FULL, ir, o); if (lane >= o) ir += tr; }
        int r2 = pos_r + (int)(ir - lr); unsigned cnt = 0;
        #pragma unroll
        for (int j = 0; j < 4; ++j) { const unsigned op = ww[j] & 15u, len = ww[j] >> 4;
            if (evm & (1u << j)) { const int rs = op == 2u ? r2 + (int)len : r2; cnt += (rs >= tk_start && rs < tk_end); }
            r2 += (int)(len * (op_flags(op) >> 1)); }
        count = __reduce_add_sync(FULL, cnt);
        if (count && lane == 0) {
            const unsigned long long e = atomicAdd(P->n_ev, 1ULL);
            if (e < P->ev_cap) { EvSlice s; s.rec = rec; s.base = base; s.pos_q = pos_q; s.pos_r = pos_r; s.k0 = k0; s.count = count; s.slot0 = 0; s.pad = 0;
                *reinterpret_cast<uint4*>(&P->ev[e]) = *reinterpret_cast<const uint4*>(&s); *(reinterpret_cast<uint4*>(&P->ev[e]) + 1) = *(reinterpret_cast<const uint4*>(&s) + 1); }
            else atomicAdd(&P->ctr->lead_overflow, 1ULL);
        }
    }
    return ((unsigned long long)big << 32) | count;
}

__global__ void __launch_bounds__(256, 4) k_scan(const __grid_constant__ ScanParams P) {
    const int lane = lane_id();
    const unsigned nwarps = gridDim.x * 8;
    int acc_task = -1; unsigned acc_reads = 0; unsigned long long acc_bp = 0; int acc_span = 0;
    int tk_id = -1, tk_start = 0, tk_end = 0, tk_len = 0;
    unsigned rec = blockIdx.x * 8 + (threadIdx.x >> 5);
    uint32_t wnext = (rec < P.n_rec && lane < 16) ? __ldg(reinterpret_cast<const uint32_t*>(P.rec + rec) + lane) : 0u;
    for (; rec < P.n_rec; rec += nwarps) {
        const uint32_t wcur = wnext;
        { const unsigned nrec2 = rec + nwarps; wnext = (nrec2 < P.n_rec && lane < 16) ? __ldg(reinterpret_cast<const uint32_t*>(P.rec + nrec2) + lane) : 0u; }
        const int r_task = (int)__shfl_sync(FULL, wcur, 0), r_pos = (int)__shfl_sync(FULL, wcur, 1);
        const uint32_t x2 = __shfl_sync(FULL, wcur, 2);
        const int r_flag = x2 & 0xffff, r_mapq = (x2 >> 16) & 255, r_aux = x2 >> 24;
        const int n = (int)__shfl_sync(FULL, wcur, 6), r_lseq = (int)__shfl_sync(FULL, wcur, 7);
        const uint64_t cigar_off = (uint64_t)__shfl_sync(FULL, wcur, 10) | ((uint64_t)__shfl_sync(FULL, wcur, 11) << 32);
        const uint32_t* __restrict__ cg = P.cigar + cigar_off;
        const int mis = (int)(cigar_off & 3);
        const uint4* __restrict__ cga = reinterpret_cast<const uint4*>(cg - mis) + lane;
        const int n_al = n + mis, li0 = lane * 4;
        #define LOAD_SLICE(base) (((base) + li0 < n_al) ? __ldg(cga + ((base) >> 2)) : make_uint4(0, 0, 0, 0))
        uint4 va = LOAD_SLICE(0), vb = LOAD_SLICE(128), vc = LOAD_SLICE(256);
        const uint32_t c_first = n > 0 ? __ldg(cg) : 0u, c_last = n > 0 ? __ldg(cg + n - 1) : 0u;
        if (r_task != tk_id) { const snfb_task t = P.task[r_task]; tk_id = r_task; tk_start = t.start; tk_end = t.end; tk_len = t.contig_len; }
        int qas = 0, qae = r_lseq;
        { int op = c_first & 15; if (op == 4) qas = (int)(c_first >> 4);
          if (op == 4 || op == 5) for (int k = 1; k < n; ++k) { const uint32_t c = __ldg(cg + k); const int o2 = c & 15; if (o2 == 4) qas += (int)(c >> 4); else if (o2 != 5) break; } }
        if (n > 1) { int op = c_last & 15; if (op == 4) qae -= (int)(c_last >> 4);
          if (op == 4 || op == 5) for (int k = n - 2; k >= 1; --k) { const uint32_t c = __ldg(cg + k); const int o2 = c & 15; if (o2 == 4) qae -= (int)(c >> 4); else if (o2 != 5) break; } }
        const int alen = qae - qas;
        const bool pass = !(r_mapq < P.mapq_min || (r_flag & 256) || alen < P.alen_min) && !(P.excl && (r_flag & P.excl)) && r_pos >= tk_start && r_pos < tk_end && n > 0;
        if (!pass) {
            if (lane == 0) { P.rec_end[rec] = -1; P.rec_flags[rec] = 0; P.rec_nm[rec] = -1.0; P.rec_nlead[rec] = 0; }
            continue;
        }
        unsigned pos_q = 0; int pos_r = r_pos; unsigned big = 0, nlead = 0;
        #define SLICE_BODY(v, base) { \
            const int li = (base) + li0; \
            uint32_t w0 = (v).x, w1 = (v).y, w2 = (v).z, w3 = (v).w; \
            if (li < mis || li + 3 >= n_al) { \
                if (li < mis || li >= n_al) w0 = 6u; if (li + 1 < mis || li + 1 >= n_al) w1 = 6u; if (li + 2 < mis || li + 2 >= n_al) w2 = 6u; if (li + 3 < mis || li + 3 >= n_al) w3 = 6u; } \
            unsigned lq = 0, lr = 0; bool rare = false; \
            { const unsigned op = w0 & 15u, len = w0 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); } \
            { const unsigned op = w1 & 15u, len = w1 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); } \
            { const unsigned op = w2 & 15u, len = w2 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); } \
            { const unsigned op = w3 & 15u, len = w3 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); } \
            const unsigned tot_q = __reduce_add_sync(FULL, lq), tot_r = __reduce_add_sync(FULL, lr); \
            if (__any_sync(FULL, rare)) { const unsigned long long rr = scan_rare(&P, w0, w1, w2, w3, lq, lr, rec, (base), pos_q, pos_r, nlead, tk_start, tk_end); \
                big += (unsigned)(rr >> 32); nlead += (unsigned)rr; } \
            pos_q += tot_q; pos_r += (int)tot_r; }
        for (int base = 0; base < n_al; base += 384) {
            { const uint4 v = va; va = LOAD_SLICE(base + 384); SLICE_BODY(v, base) }
            if (base + 128 < n_al) { const uint4 v = vb; vb = LOAD_SLICE(base + 512); SLICE_BODY(v, base + 128) }
            if (base + 256 < n_al) { const uint4 v = vc; vc = LOAD_SLICE(base + 640); SLICE_BODY(v, base + 256) }
        }
        #undef SLICE_BODY
        #undef LOAD_SLICE
        const int ref_end = pos_r;
        int hp = (r_aux & SNFB_AUX_HP) ? (int)(__shfl_sync(FULL, wcur, 3) & 255u) : 0;
        if (hp > 2) { hp = 0; if (lane == 0) atomicAdd(&P.ctr->soft_errors, 1ULL); }
        const bool has_nm = P.want_nm && (r_aux & SNFB_AUX_NM);
        const int r_nm = (int)__shfl_sync(FULL, wcur, 4);
        if (lane == 0) {
            P.rec_end[rec] = ref_end;
            P.rec_flags[rec] = (uint8_t)(RF_PASS | (has_nm ? RF_HAS_NM : 0) | (hp << 2));
            P.rec_nm[rec] = has_nm ? __ddiv_rn((double)((long long)r_nm - (long long)big), (double)(alen + 1)) : -1.0;      // leadprov.py:517-526
            P.rec_nlead[rec] = nlead;
            if (r_aux & SNFB_AUX_SA) { const unsigned long long e = atomicAdd(P.n_sa, 1ULL); P.sa_list[e] = rec; }
            if (acc_task != r_task) {
                if (acc_task >= 0) { atomicAdd(&P.task_reads[acc_task], acc_reads); atomicAdd(&P.task_cov_bp[acc_task], acc_bp); atomicMax(&P.task_maxspan[acc_task], acc_span); }
                acc_task = r_task; acc_reads = 0; acc_bp = 0; acc_span = 0;
            }
            ++acc_reads;
            const int ce = ref_end < tk_len ? ref_end : tk_len;
            if (ce > r_pos) acc_bp += (unsigned long long)(ce - r_pos);
            if (ref_end - r_pos > acc_span) acc_span = ref_end - r_pos;
        }
    }
    if (lane == 0 && acc_task >= 0) { atomicAdd(&P.task_reads[acc_task], acc_reads); atomicAdd(&P.task_cov_bp[acc_task], acc_bp); atomicMax(&P.task_maxspan[acc_task], acc_span); }
}

// ------------------------------------------------------------------------------------------------
// k_scan_tma: the same streaming pass with the CIGAR staged through shared memory by the bulk-copy engine
// (cp.async.bulk global -> shared, completion on an mbarrier).  Every warp owns a ring of NST 2-KB stages; lane 0 runs
// a fetch cursor up to NST chunks (and records) ahead of the consume cursor, so each warp keeps several KB in flight
// without holding them in registers — the loads are no longer tied to the warp's own issue slots.
// ------------------------------------------------------------------------------------------------
namespace tma {
constexpr int NST = 4;                 // stages per warp
constexpr int CH_OPS = 512;            // ops per stage (2 KB = four 128-op slices)
constexpr int WPB = 4;                 // warps per block (4 x 4 x 2 KB = 32 KB of stages)
struct Desc { uint32_t rec; int32_t pos; int32_t n_al; int32_t mis; int32_t alen; int32_t nchunks; uint32_t aux_hp_nm; int32_t task; int32_t nm; int32_t tk_start, tk_end, tk_len; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
}  // namespace tma

__global__ void __launch_bounds__(tma::WPB * 32, 6) k_scan_tma(const __grid_constant__ ScanParams P) {
    using namespace tma;
    __shared__ __align__(128) uint8_t s_buf[WPB][NST][CH_OPS * 4];
    __shared__ __align__(8) uint64_t s_bar[WPB][NST];
    __shared__ Desc s_desc[WPB][NST + 1];
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    const unsigned nwarps = gridDim.x * WPB;
    if (lane == 0) { for (int i = 0; i < NST; ++i) mbar_init(&s_bar[wib][i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    int acc_task = -1; unsigned acc_reads = 0; unsigned long long acc_bp = 0; int acc_span = 0;
    // ---- fetch cursor (warp-uniform state; lane 0 acts) ----
    unsigned f_rec = blockIdx.x * WPB + wib;                // record the cursor stands on
    int f_chunk = 0, f_nchunks = 0; bool f_open = false;    // f_open: descriptor of f_rec already published
    const uint32_t* f_base = nullptr; int f_nal = 0;
    unsigned n_fetch = 0, n_cons = 0, d_w = 0, d_r = 0;     // chunk / descriptor ring counters
    int tk_id = -1, tk_start = 0, tk_end = 0, tk_len = 0;
    // record core of the cursor and of the record after it (prefetched)
    uint32_t wF = (f_rec < P.n_rec && lane < 16) ? __ldg(reinterpret_cast<const uint32_t*>(P.rec + f_rec) + lane) : 0u;
    uint32_t wN = (f_rec + nwarps < P.n_rec && lane < 16) ? __ldg(reinterpret_cast<const uint32_t*>(P.rec + f_rec + nwarps) + lane) : 0u;
    // ---- consume cursor ----
    int c_chunk = 0; unsigned pos_q = 0; int pos_r = 0; unsigned big = 0, nlead = 0;
    for (;;) {
        // (1) keep the ring full
        while (n_fetch - n_cons < (unsigned)NST && f_rec < P.n_rec) {
            if (!f_open) {
                const int r_task = (int)__shfl_sync(FULL, wF, 0), r_pos = (int)__shfl_sync(FULL, wF, 1);
                const uint32_t x2 = __shfl_sync(FULL, wF, 2);
                const int r_flag = x2 & 0xffff, r_mapq = (x2 >> 16) & 255, r_aux = x2 >> 24;
                const int n = (int)__shfl_sync(FULL, wF, 6), r_lseq = (int)__shfl_sync(FULL, wF, 7);
                const uint64_t cigar_off = (uint64_t)__shfl_sync(FULL, wF, 10) | ((uint64_t)__shfl_sync(FULL, wF, 11) << 32);
                const uint32_t* cg = P.cigar + cigar_off; const int mis = (int)(cigar_off & 3);
                if (r_task != tk_id) { const snfb_task t = P.task[r_task]; tk_id = r_task; tk_start = t.start; tk_end = t.end; tk_len = t.contig_len; }
                // cheap rejects first (no CIGAR access), then the alignment length from the clip ops
                bool pass = !(r_mapq < P.mapq_min || (r_flag & 256)) && !(P.excl && (r_flag & P.excl)) && r_pos >= tk_start && r_pos < tk_end && n > 0;
                int alen = 0;
                if (pass) {
                    const uint32_t c_first = __ldg(cg), c_last = __ldg(cg + n - 1);
                    int qas = 0, qae = r_lseq;
                    { int op = c_first & 15; if (op == 4) qas = (int)(c_first >> 4);
                      if (op == 4 || op == 5) for (int k = 1; k < n; ++k) { const uint32_t c = __ldg(cg + k); const int o2 = c & 15; if (o2 == 4) qas += (int)(c >> 4); else if (o2 != 5) break; } }
                    if (n > 1) { int op = c_last & 15; if (op == 4) qae -= (int)(c_last >> 4);
                      if (op == 4 || op == 5) for (int k = n - 2; k >= 1; --k) { const uint32_t c = __ldg(cg + k); const int o2 = c & 15; if (o2 == 4) qae -= (int)(c >> 4); else if (o2 != 5) break; } }
                    alen = qae - qas; pass = alen >= P.alen_min;
                }
                if (!pass) {
                    if (lane == 0) { P.rec_end[f_rec] = -1; P.rec_flags[f_rec] = 0; P.rec_nm[f_rec] = -1.0; P.rec_nlead[f_rec] = 0; }
                    f_rec += nwarps; wF = wN; { const unsigned nx = f_rec + nwarps; wN = (nx < P.n_rec && lane < 16) ? __ldg(reinterpret_cast<const uint32_t*>(P.rec + nx) + lane) : 0u; }
                    continue;
                }
                f_base = cg - mis; f_nal = n + mis; f_nchunks = (f_nal + CH_OPS - 1) / CH_OPS; f_chunk = 0; f_open = true;
                if (lane == 0) {
                    Desc d; d.rec = f_rec; d.pos = r_pos; d.n_al = f_nal; d.mis = mis; d.alen = alen; d.nchunks = f_nchunks;
                    d.aux_hp_nm = (uint32_t)r_aux;                     // hp / nm are added right below (they need a warp shuffle)
                    d.task = r_task; d.nm = 0; d.tk_start = tk_start; d.tk_end = tk_end; d.tk_len = tk_len;
                    s_desc[wib][d_w % (NST + 1)] = d;
                }
                { const uint32_t x3 = __shfl_sync(FULL, wF, 3); const int r_nm = (int)__shfl_sync(FULL, wF, 4);
                  if (lane == 0) { Desc* dp = &s_desc[wib][d_w % (NST + 1)]; dp->aux_hp_nm = (uint32_t)r_aux | ((x3 & 255u) << 8); dp->nm = r_nm; } }
                ++d_w;
                __syncwarp();
            }
            {   // issue one chunk of the open record
                const int st = n_fetch % NST; const int ops = min(CH_OPS, f_nal - f_chunk * CH_OPS); const unsigned bytes = (unsigned)(((ops * 4) + 15) & ~15);
                if (lane == 0) { mbar_expect_tx(&s_bar[wib][st], bytes); bulk_g2s(s_buf[wib][st], f_base + (size_t)f_chunk * CH_OPS, bytes, &s_bar[wib][st]); }
                ++n_fetch; ++f_chunk;
                if (f_chunk == f_nchunks) { f_open = false; f_rec += nwarps; wF = wN; { const unsigned nx = f_rec + nwarps; wN = (nx < P.n_rec && lane < 16) ? __ldg(reinterpret_cast<const uint32_t*>(P.rec + nx) + lane) : 0u; } }
            }
        }
        if (n_cons == n_fetch) break;                     // nothing in flight and nothing left to fetch
        // (2) consume one chunk
        const int st = n_cons % NST;
        mbar_wait(&s_bar[wib][st], (n_cons / NST) & 1u);
        const Desc d = s_desc[wib][d_r % (NST + 1)];
        if (c_chunk == 0) { pos_q = 0; pos_r = d.pos; big = 0; nlead = 0; }
        const uint4* sb = reinterpret_cast<const uint4*>(s_buf[wib][st]) + lane;
        const int cbase = c_chunk * CH_OPS;
        #pragma unroll 1
        for (int sl = 0; sl < 4; ++sl) {
            const int base = cbase + sl * 128; if (base >= d.n_al) break;
            const uint4 v = sb[sl * 32];
            const int li = base + lane * 4;
            uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
            if (li < d.mis || li + 3 >= d.n_al) {
                if (li < d.mis || li >= d.n_al) w0 = 6u; if (li + 1 < d.mis || li + 1 >= d.n_al) w1 = 6u; if (li + 2 < d.mis || li + 2 >= d.n_al) w2 = 6u; if (li + 3 < d.mis || li + 3 >= d.n_al) w3 = 6u; }
            unsigned lq = 0, lr = 0; bool rare = false;
            { const unsigned op = w0 & 15u, len = w0 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); }
            { const unsigned op = w1 & 15u, len = w1 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); }
            { const unsigned op = w2 & 15u, len = w2 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); }
            { const unsigned op = w3 & 15u, len = w3 >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1); rare |= (len > 10u) & (op != 0u); }
            const unsigned tot_q = __reduce_add_sync(FULL, lq), tot_r = __reduce_add_sync(FULL, lr);
            if (__any_sync(FULL, rare)) { const unsigned long long rr = scan_rare(&P, w0, w1, w2, w3, lq, lr, d.rec, base, pos_q, pos_r, nlead, d.tk_start, d.tk_end);
                big += (unsigned)(rr >> 32); nlead += (unsigned)rr; }
            pos_q += tot_q; pos_r += (int)tot_r;
        }
        __syncwarp();                                      // all lanes are done with this stage before lane 0 may refill it
        ++n_cons; ++c_chunk;
        if (c_chunk == d.nchunks) {
            c_chunk = 0; ++d_r;
            if (lane == 0) {
                const int r_aux = d.aux_hp_nm & 255; int hp = (r_aux & SNFB_AUX_HP) ? (int)((d.aux_hp_nm >> 8) & 255u) : 0;
                if (hp > 2) { hp = 0; atomicAdd(&P.ctr->soft_errors, 1ULL); }
                const bool has_nm = P.want_nm && (r_aux & SNFB_AUX_NM); const int ref_end = pos_r;
                P.rec_end[d.rec] = ref_end;
                P.rec_flags[d.rec] = (uint8_t)(RF_PASS | (has_nm ? RF_HAS_NM : 0) | (hp << 2));
                P.rec_nm[d.rec] = has_nm ? __ddiv_rn((double)((long long)d.nm - (long long)big), (double)(d.alen + 1)) : -1.0;
                P.rec_nlead[d.rec] = nlead;
                if (r_aux & SNFB_AUX_SA) { const unsigned long long e = atomicAdd(P.n_sa, 1ULL); P.sa_list[e] = d.rec; }
                if (acc_task != d.task) {
                    if (acc_task >= 0) { atomicAdd(&P.task_reads[acc_task], acc_reads); atomicAdd(&P.task_cov_bp[acc_task], acc_bp); atomicMax(&P.task_maxspan[acc_task], acc_span); }
                    acc_task = d.task; acc_reads = 0; acc_bp = 0; acc_span = 0;
                }
                ++acc_reads;
                const int ce = ref_end < d.tk_len ? ref_end : d.tk_len;
                if (ce > d.pos) acc_bp += (unsigned long long)(ce - d.pos);
                if (ref_end - d.pos > acc_span) acc_span = ref_end - d.pos;
            }
        }
    }
    if (lane == 0 && acc_task >= 0) { atomicAdd(&P.task_reads[acc_task], acc_reads); atomicAdd(&P.task_cov_bp[acc_task], acc_bp); atomicMax(&P.task_maxspan[acc_task], acc_span); }
}

// exact lead slots of the event slices: exclusive prefix of their counts (slices of one record keep their order through k0)
__global__ void k_ev_counts(const EvSlice* __restrict__ ev, uint32_t* __restrict__ cnt, const unsigned long long* __restrict__ n_ev, unsigned long long bound) {
    const unsigned long long n = *n_ev < bound ? *n_ev : bound;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < bound; i += (unsigned long long)gridDim.x * blockDim.x) cnt[i] = i < n ? ev[i].count : 0u;
}

struct EmitParams {
    const snfb_rec* rec; const uint32_t* cigar; const uint8_t* var; const snfb_task* task;
    const EvSlice* ev; const uint32_t* ev_slot; const unsigned long long* n_ev; unsigned long long ev_cap;
    snfb_lead* leads; unsigned long long lead_cap; DevCounters* ctr;
    int minsv, maxlen, detect_large_ins; double longinslen;
};
// one warp per event slice: read_iterindels' lead construction (leadprov.py:583-670)
__global__ void __launch_bounds__(256) k_emit(const EmitParams P) {
    const int lane = lane_id();
    const unsigned long long n = *P.n_ev < P.ev_cap ? *P.n_ev : P.ev_cap;
    const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long e = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += nw) {
        const EvSlice s = P.ev[e];
        const snfb_rec r = P.rec[s.rec];
        const snfb_task tk = P.task[r.task];
        const uint32_t* cg = P.cigar + r.cigar_off; const int n_ops = (int)r.n_cigar;
        const int mis = (int)(r.cigar_off & 3); const int n_al = n_ops + mis;
        const int li = s.base + lane * 4;
        uint4 v = (li < n_al) ? __ldg(reinterpret_cast<const uint4*>(cg - mis) + (s.base >> 2) + lane) : make_uint4(0, 0, 0, 0);
        uint32_t ww[4] = { v.x, v.y, v.z, v.w };
        #pragma unroll
        for (int j = 0; j < 4; ++j) if (li + j < mis || li + j >= n_al) ww[j] = 6u;
        unsigned lq = 0, lr = 0, evm = 0;
        #pragma unroll
        for (int j = 0; j < 4; ++j) { const unsigned op = ww[j] & 15u, len = ww[j] >> 4, fl = op_flags(op); lq += len * (fl & 1u); lr += len * (fl >> 1);
            if (((0x016u >> op) & 1u) && (int)len >= P.minsv) evm |= 1u << j; }
        unsigned iq = lq, ir = lr;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { unsigned tq = __shfl_up_sync(FULL, iq, o), tr = __shfl_up_sync(FULL, ir, o); if (lane >= o) { iq += tq; ir += tr; } }
        unsigned pq = s.pos_q + iq - lq; int pr = s.pos_r + (int)(ir - lr);
        // which signatures stay inside the task's region (leadprov.py:464-466), and their rank inside the slice
        unsigned emm = 0; int cnt = 0; { int r2 = pr;
            #pragma unroll
            for (int j = 0; j < 4; ++j) { const unsigned op = ww[j] & 15u, len = ww[j] >> 4;
                if (evm & (1u << j)) { const int rs = op == 2u ? r2 + (int)len : r2; if (rs >= tk.start && rs < tk.end) { emm |= 1u << j; ++cnt; } }
                r2 += (int)(len * (op_flags(op) >> 1)); } }
        int inc = cnt;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        int mine = inc - cnt;
        // per-record facts
        int qas = 0, qae = r.l_seq;
        for (int k = 0; k < n_ops; ++k) { const uint32_t c = __ldg(cg + k); const int op = c & 15; if (op == 4) qas += (int)(c >> 4); else if (op != 5) break; }
        for (int k = n_ops - 1; k >= 1; --k) { const uint32_t c = __ldg(cg + k); const int op = c & 15; if (op == 4) qae -= (int)(c >> 4); else if (op != 5) break; }
        const int alen = qae - qas;
        const bool is_supp = r.flag & 2048, rev = r.flag & 16, has_sa = r.aux_flags & SNFB_AUX_SA;
        int hp = (r.aux_flags & SNFB_AUX_HP) ? r.hp : 0; if (hp > 2) hp = 0;
        const bool use_clips = P.detect_large_ins && !is_supp && !has_sa;
        const uint32_t inl_flags = (rev ? SNFB_LF_REVERSE : 0u) | ((uint32_t)r.mapq << 16) | ((uint32_t)SNFB_SRC_INLINE << 3) | ((uint32_t)hp << 24) | (is_supp ? SNFB_LF_IS_SA : 0u);
        const uint64_t qh = qname_hash_warp(P.var + r.var_off, r.l_qname);
        const unsigned long long slot0 = P.ev_slot[e];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned op = ww[j] & 15u; const int len = (int)(ww[j] >> 4); const unsigned fl = op_flags(op);
            if (emm & (1u << j)) {
                snfb_lead L;
                L.rec = s.rec; L.qname_hash = qh; L.read_len = alen; L.seq_off = -1; L.seq_len = 0; L.mate_contig = -1; L.mate_pos = 0; L.nm_sa = 0;
                L.task = (uint16_t)r.task; L.k = (uint16_t)(s.k0 + mine);
                uint32_t f = inl_flags; const int pqi = (int)pq;
                if (op == 1u) { f |= SNFB_INS; L.ref_start = pr; L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi + len; L.svlen = len;
                    if (len <= P.maxlen) { f |= SNFB_LF_HAS_SEQ; L.seq_off = pqi; L.seq_len = len; } }
                else if (op == 2u) { f |= SNFB_DEL; L.ref_start = pr + len; L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi; L.svlen = -len; }
                else if (use_clips && (double)len >= P.longinslen) { f |= SNFB_INS | SNFB_LF_SVLEN_NONE; L.ref_start = L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi + len; L.svlen = 0; }
                else { f |= (pr == r.pos) ? SNFB_SINGLE_LEFT : SNFB_SINGLE_RIGHT; L.ref_start = L.ref_end = pr; L.qry_start = pqi; L.qry_end = pqi + len; L.svlen = 0; }
                L.flags = f;
                const unsigned long long slot = slot0 + (unsigned long long)mine;
                if (slot < P.lead_cap) store_lead(P.leads + slot, L); else atomicAdd(&P.ctr->lead_overflow, 1ULL);
                ++mine;
            }
            pq += (unsigned)len * (fl & 1u); pr += (int)((unsigned)len * (fl >> 1));
        }
    }
}

struct SaParams {
    const snfb_rec* rec; const uint32_t* cigar; const uint8_t* var; const snfb_task* task; const snfb_contig* contig; uint32_t n_contig;
    const uint32_t* sa_list; const unsigned long long* n_sa; const int32_t* rec_end; uint32_t* rec_nlead;
    snfb_lead* leads; unsigned long long lead_cap; DevCounters* ctr;
    snfb_config cfg;
};
// one warp per record with an SA tag (lane 0 parses; the warp hashes the name)
__global__ void __launch_bounds__(THREADS) k_sa(const SaParams P) {
    __shared__ Seg segs[WARPS][MAXSEG];
    __shared__ snfb_config s_cfg;
    if (threadIdx.x < sizeof(snfb_config) / 4) reinterpret_cast<uint32_t*>(&s_cfg)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&P.cfg)[threadIdx.x];
    __syncthreads();
    const int lane = lane_id(), wib = threadIdx.x >> 5;
    const unsigned long long n = *P.n_sa;
    const unsigned long long nw = (unsigned long long)gridDim.x * WARPS;
    unsigned long long soft = 0, overflow = 0;
    SlotState slots; slots.cur = 0; slots.end = 0;
    for (unsigned long long e = (unsigned long long)blockIdx.x * WARPS + wib; e < n; e += nw) {
        const uint32_t rec = P.sa_list[e];
        const snfb_rec r = P.rec[rec];
        const snfb_task tk = P.task[r.task];
        const uint32_t* cg = P.cigar + r.cigar_off; const int n_ops = (int)r.n_cigar;
        int qas = 0, qae = r.l_seq;
        for (int k = 0; k < n_ops; ++k) { const uint32_t c = __ldg(cg + k); const int op = c & 15; if (op == 4) qas += (int)(c >> 4); else if (op != 5) break; }
        for (int k = n_ops - 1; k >= 1; --k) { const uint32_t c = __ldg(cg + k); const int op = c & 15; if (op == 4) qae -= (int)(c >> 4); else if (op != 5) break; }
        const uint64_t qh = qname_hash_warp(P.var + r.var_off, r.l_qname);
        if (lane == 0) {
            int hp = (r.aux_flags & SNFB_AUX_HP) ? r.hp : 0; if (hp > 2) hp = 0;
            const bool rev = r.flag & 16;
            SaArgs a; a.rec = rec; a.qas = qas; a.qae = qae; a.alen = qae - qas; a.ref_end = P.rec_end[rec]; a.hp = hp;
            a.base_flags = (rev ? SNFB_LF_REVERSE : 0u) | ((uint32_t)r.mapq << 16); a.qh = qh; a.nlead = P.rec_nlead[rec]; a.rev = rev; a.is_supp = r.flag & 2048;
            a.sa = P.var + r.var_off + r.l_qname; a.sa_len = (int)r.sa_len; a.c_first = __ldg(cg); a.c_last = __ldg(cg + n_ops - 1); a.pos = r.pos; a.l_seq = r.l_seq; a.mapq = r.mapq;
            a.aux_flags = r.aux_flags; a.task = r.task; a.tk_contig = tk.contig; a.tk_start = tk.start; a.tk_end = tk.end; a.contig = P.contig; a.n_contig = P.n_contig;
            a.leads = P.leads; a.lead_cap = P.lead_cap; a.n_slots = &P.ctr->n_slots; a.slots = &slots;
            a.mapq_min = s_cfg.mapq; a.dev_keep_lowqual_splits = s_cfg.dev_keep_lowqual_splits; a.max_splits_base = s_cfg.max_splits_base; a.max_splits_kb = s_cfg.max_splits_kb;
            const unsigned added = process_sa(&s_cfg, segs[wib], a, &soft, &overflow);
            if (added) P.rec_nlead[rec] += added;
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (unsigned long long sidx = slots.cur; sidx < slots.end; ++sidx) if (sidx < P.lead_cap) P.leads[sidx].rec = HOLE;   // retire the last chunk
        if (soft) atomicAdd(&P.ctr->soft_errors, soft);
        if (overflow) atomicAdd(&P.ctr->lead_overflow, overflow);
    }
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
