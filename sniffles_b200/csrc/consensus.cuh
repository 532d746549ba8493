// consensus.cuh — stage C: INS ALT sequences.
//   postprocessing.annotate_sv INS branch (best read selection)        postprocessing.py:33-66
//   consensus.novel_from_reads (k-mer anchored pile-up polish, k = 6)  consensus.py:280-394
// k_plan picks the best read per candidate and sizes the work; k_prep unpacks the best read and builds its anchor
// table; k_align aligns one (candidate, supporting read) pair per warp; k_vote takes the column vote per 4096-column tile.
#pragma once
#include "common.cuh"

namespace consensus {

constexpr int THREADS = 128;
constexpr int TAB = 2048;                  // > 4 x the at most ~500 strided k-mers of the best read
constexpr uint8_t DASH = 0xff;

struct C {
    const snfb_cand* cand; snfb_cand* cand_rw; const snfb_lead* cand_leads;
    const uint32_t* out_plo; const uint32_t* out_pn;     // per candidate lead: the run of `ord` entries (merge_inner parts) it was folded from
    const uint32_t* ord; const snfb_lead* kleads;         // kept-lead indices in merge_inner order, the kept leads
    const snfb_rec* rec; const uint8_t* seq;
    const uint32_t* arena_off;       // seq on demand: per kept lead, 16-byte unit offset of its bytes in `seq` (which then is the compact arena); nullptr = full arena
    uint32_t* plan_best; uint32_t* plan_nother; uint32_t* plan_otot; uint32_t* alt_len; uint32_t* scr_len; uint32_t* alt_off; uint32_t* scr_off;   // scr in units of 16 bytes
    uint8_t* alt; uint8_t* scr; unsigned long long alt_cap, scr_cap16, cand_cap;
    uint32_t* work_big; uint32_t* work_small; uint32_t* work_ctr;      // work_ctr: [0] n_big, [1] n_small, [2],[3] queue positions, [4] n_items_big, [5] n_items_small, [6],[7] item queue positions, [8] n_tiles, [9] tile queue position
    // item pipeline: one (candidate, supporting read) pair per warp, one (candidate, column tile) per block
    struct Item { uint32_t cand; uint32_t k; uint32_t row; uint32_t rd_off; };
    Item* items_big; Item* items_small; uint2* tiles; unsigned long long item_cap, tile_cap;
    unsigned long long* dbg;          // optional (SNFB_DEBUG): per warp of k_align [busy cycles, end time, items, longest item cycles, its L, its Lo]
    DevCounters* ctr; snfb_config cfg;
};

__device__ __forceinline__ uint8_t seq_code(const uint8_t* sq, long long q) { const uint8_t v = sq[q >> 1]; return (q & 1) ? (v & 15) : (v >> 4); }

// choose the best read, size the outputs
__global__ void k_plan(C c) {
    const unsigned long long nc = c.ctr->n_cand < c.cand_cap ? c.ctr->n_cand : c.cand_cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nc; i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t al = 0, sl = 0;
        if (c.cand[i].svtype == SNFB_INS && !c.cfg.symbolic) {
            const snfb_cand* cd = &c.cand[i]; long nm = 0, bi = -1; long long bd = 0; long long tot = 0;
            for (int k = 0; k < cd->lead_n; ++k) { const snfb_lead* l = &c.cand_leads[cd->lead_off + k]; if (!(l->flags & SNFB_LF_HAS_SEQ)) continue;
                // abs(len(seq) - svlen) + abs(ref_start - pos) * 1.5, compared exactly in halves
                long long a = (long long)l->seq_len - cd->svlen; if (a < 0) a = -a; long long p = (long long)l->ref_start - cd->pos; if (p < 0) p = -p;
                const long long d = 2 * a + 3 * p; if (nm == 0 || d < bd) { bd = d; bi = k; } ++nm; tot += l->seq_len; }
            if (nm > 0) {
                const uint32_t L = (uint32_t)c.cand_leads[cd->lead_off + bi].seq_len; al = L;
                const bool cons = (nm - 1 >= c.cfg.consensus_min_reads) && !c.cfg.no_consensus;
                c.plan_best[i] = (uint32_t)bi; c.plan_nother[i] = cons ? (uint32_t)(nm - 1) : 0u; c.plan_otot[i] = (uint32_t)(tot - L);
                // scratch: best codes | others' codes | (4-byte aligned) one row of align4(L) per other read | accept flags | anchor table; 16-byte units
                const unsigned long long bytes = cons ? (unsigned long long)tot + 4 + (unsigned long long)(nm - 1) * ((L + 3u) & ~3u) + (unsigned long long)(nm - 1) * 16 + 64 + 16 + TAB * 8 : (unsigned long long)L + 16;
                sl = (uint32_t)((bytes + 15) / 16);
                c.cand_rw[i].alt_len = (int)L;
                // work queue: the heavy tail (long insertions with many reads) is scheduled first
                const unsigned long long work = (unsigned long long)L * (unsigned long long)nm;
                if (work > 60000ull) c.work_big[atomicAdd(&c.work_ctr[0], 1u)] = (uint32_t)i; else c.work_small[atomicAdd(&c.work_ctr[1], 1u)] = (uint32_t)i;
                if (cons) {
                    // one work item per supporting read (heavy rows first) and one per 4096-column tile of the vote
                    const bool heavy = L > 4000u; C::Item* dst = heavy ? c.items_big : c.items_small;
                    const uint32_t base = atomicAdd(&c.work_ctr[heavy ? 4 : 5], (uint32_t)(nm - 1));
                    uint32_t row = 0, ro = 0;
                    for (int k = 0; k < cd->lead_n; ++k) { const snfb_lead* l = &c.cand_leads[cd->lead_off + k]; if (!(l->flags & SNFB_LF_HAS_SEQ) || k == bi) continue;
                        if ((unsigned long long)base + row < c.item_cap) { C::Item it; it.cand = (uint32_t)i; it.k = (uint32_t)k; it.row = row; it.rd_off = ro; dst[base + row] = it; }
                        ++row; ro += (uint32_t)l->seq_len; }
                    const uint32_t nt = (L + 4095u) / 4096u; const uint32_t tb = atomicAdd(&c.work_ctr[8], nt);
                    for (uint32_t t = 0; t < nt; ++t) if ((unsigned long long)tb + t < c.tile_cap) c.tiles[tb + t] = make_uint2((uint32_t)i, t);
                }
            }
        }
        c.alt_len[i] = al; c.scr_len[i] = sl;
    }
}

// the candidate records are final once the ALT offsets are known (stage C only fills the ALT bytes)
__global__ void k_plan_finish(C c) {
    const unsigned long long nc = c.ctr->n_cand < c.cand_cap ? c.ctr->n_cand : c.cand_cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nc; i += (unsigned long long)gridDim.x * blockDim.x)
        if (c.scr_len[i]) c.cand_rw[i].alt_off = (int)c.alt_off[i];
}

// unpack `len` bases starting at nibble `off` of sq into dst (one code per byte); `tid`/`nthr` = cooperating threads.
// Each thread takes 8 consecutive bases per step and four steps are kept in flight so that the 4-bit arena is
// streamed from HBM with enough loads outstanding.
__device__ __forceinline__ void unpack_span(const uint8_t* __restrict__ sq, long long off, int len, uint8_t* __restrict__ dst, int tid, int nthr) {
    const int stride = nthr * 8;
    for (int j0 = tid * 8; j0 < len; j0 += 4 * stride) {
        uint8_t v[4][5];
        #pragma unroll
        for (int u = 0; u < 4; ++u) { const int jb = j0 + u * stride; const long long q = off + jb; const uint8_t* p = sq + (q >> 1);
            #pragma unroll
            for (int t = 0; t < 5; ++t) v[u][t] = (jb + 2 * t - (int)(q & 1) < len + 1 && jb < len) ? __ldg(p + t) : (uint8_t)0; }
        #pragma unroll
        for (int u = 0; u < 4; ++u) { const int jb = j0 + u * stride; const int ph = (int)((off + jb) & 1);
            #pragma unroll
            for (int t = 0; t < 8; ++t) { const int nb = t + ph; const uint8_t by = v[u][nb >> 1]; if (jb + t < len) dst[jb + t] = (nb & 1) ? (by & 15) : (by >> 4); } }
    }
}
// unpack the (possibly merged) sequence of candidate lead `cl_index` as 4-bit codes, one byte per base
__device__ inline void unpack_lead(const C& c, uint32_t cl_index, uint8_t* dst) {
    const uint32_t plo = c.out_plo[cl_index], pn = c.out_pn[cl_index];
    long long o = 0;
    for (uint32_t p = 0; p < pn; ++p) {
        const uint32_t slot = c.ord[plo + p]; const snfb_lead* l = &c.kleads[slot];
        const uint8_t* sq = c.arena_off ? c.seq + (size_t)c.arena_off[slot] * 16 : c.seq + c.rec[l->rec].seq_off;
        unpack_span(sq, c.arena_off ? (l->seq_off & 1) : l->seq_off, l->seq_len, dst + o, threadIdx.x, blockDim.x);
        o += l->seq_len;
    }
}

__device__ __forceinline__ uint32_t kmer6(const uint8_t* s) { return (uint32_t)s[0] | ((uint32_t)s[1] << 4) | ((uint32_t)s[2] << 8) | ((uint32_t)s[3] << 12) | ((uint32_t)s[4] << 16) | ((uint32_t)s[5] << 20); }
__device__ __forceinline__ uint32_t kslot(uint32_t key) { return (key * 2654435761u) >> 21; }    // top 11 bits

// seq on demand: the base slices stage C will read, as (source byte offset in the host seq arena, bytes, destination unit)
struct SeqReq { unsigned long long src; uint32_t nbytes; uint32_t dst16; };
__global__ void k_seq_requests(C c, SeqReq* req, unsigned long long req_cap, uint32_t* arena_off_rw, unsigned long long* n_req, unsigned long long* n_units) {
    const unsigned long long nc = c.ctr->n_cand < c.cand_cap ? c.ctr->n_cand : c.cand_cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nc; i += (unsigned long long)gridDim.x * blockDim.x) {
        if (c.scr_len[i] == 0) continue;
        const snfb_cand* cd = &c.cand[i]; const bool cons = c.plan_nother[i] > 0;
        for (int k = 0; k < cd->lead_n; ++k) {
            const snfb_lead* cl = &c.cand_leads[cd->lead_off + k];
            if (!(cl->flags & SNFB_LF_HAS_SEQ) || (!cons && (uint32_t)k != c.plan_best[i])) continue;
            const uint32_t plo = c.out_plo[cd->lead_off + k], pn = c.out_pn[cd->lead_off + k];
            for (uint32_t p = 0; p < pn; ++p) {
                const uint32_t slot = c.ord[plo + p]; const snfb_lead* l = &c.kleads[slot];
                const unsigned long long b0 = (unsigned long long)l->seq_off >> 1, b1 = ((unsigned long long)l->seq_off + l->seq_len + 1) >> 1;
                const uint32_t nb = (uint32_t)(b1 - b0) + 1;                                  // +1: unpack_span may touch one byte past the last base
                const unsigned long long u = atomicAdd(n_units, (unsigned long long)((nb + 15) / 16));
                const unsigned long long r = atomicAdd(n_req, 1ULL);
                arena_off_rw[slot] = (uint32_t)u;
                if (r < req_cap) { req[r].src = c.rec[l->rec].seq_off + b0; req[r].nbytes = nb; req[r].dst16 = (uint32_t)u; }
            }
        }
    }
}

constexpr int MAXHIT = 512;       // strided k-mer hits of one read are bounded by (L + 6) / skip + 1 < 512 (skip = 3 + L / 500)

// the same with the calling warp only
__device__ inline void unpack_lead_warp(const C& c, uint32_t cl_index, uint8_t* dst) {
    const uint32_t plo = c.out_plo[cl_index], pn = c.out_pn[cl_index];
    long long o = 0;
    for (uint32_t p = 0; p < pn; ++p) {
        const uint32_t slot = c.ord[plo + p]; const snfb_lead* l = &c.kleads[slot];
        const uint8_t* sq = c.arena_off ? c.seq + (size_t)c.arena_off[slot] * 16 : c.seq + c.rec[l->rec].seq_off;
        unpack_span(sq, c.arena_off ? (l->seq_off & 1) : l->seq_off, l->seq_len, dst + o, lane_id(), 32);
        o += l->seq_len;
    }
}

// ================================================================================================
// k_prep (block per candidate: unpack the best read, build its anchor table in global
// scratch, or copy the best read to ALT when there is no consensus) -> k_align (one warp per (candidate, read)
// item from a heavy-first queue: no block barriers, the heaviest candidate's reads spread over the whole GPU)
// -> k_vote (one block per (candidate, 4096-column tile)).
// ================================================================================================
// ---- byte-string helpers on unaligned pointers, four bytes per step (sliding aligned 32-bit windows).  They may read up to
//      7 bytes past the last byte asked for; every buffer they are used on is followed by other scratch of the same candidate.
__device__ __forceinline__ int match_count(const uint8_t* a, const uint8_t* b, long n) {
    const uint32_t sa = ((uintptr_t)a & 3) * 8, sb = ((uintptr_t)b & 3) * 8;
    const uint32_t* wa = reinterpret_cast<const uint32_t*>((uintptr_t)a & ~(uintptr_t)3); const uint32_t* wb = reinterpret_cast<const uint32_t*>((uintptr_t)b & ~(uintptr_t)3);
    uint32_t alo = wa[0], blo = wb[0]; int mt = 0;
    long q = 0;
    #pragma unroll 2
    for (; q + 4 <= n; q += 4) {
        const uint32_t ahi = *++wa, bhi = *++wb;
        const uint32_t x = __funnelshift_r(alo, ahi, sa) ^ __funnelshift_r(blo, bhi, sb);
        mt += __popc(~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u);       // 0x80 in every byte that is equal
        alo = ahi; blo = bhi;
    }
    if (q < n) {
        const uint32_t ahi = *++wa, bhi = *++wb;
        const uint32_t x = __funnelshift_r(alo, ahi, sa) ^ __funnelshift_r(blo, bhi, sb);
        mt += __popc(~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u & ((1u << (8 * (n - q))) - 1u));
    }
    return mt;
}
__device__ __forceinline__ void copy_bytes(uint8_t* dst, const uint8_t* src, long n) {
    long q = 0;
    while (q < n && ((uintptr_t)(dst + q) & 3)) { dst[q] = src[q]; ++q; }
    if (q + 4 <= n) {
        const uint8_t* s0 = src + q; const uint32_t sh = ((uintptr_t)s0 & 3) * 8;
        const uint32_t* ws = reinterpret_cast<const uint32_t*>((uintptr_t)s0 & ~(uintptr_t)3); uint32_t lo = ws[0];
        uint32_t* wd = reinterpret_cast<uint32_t*>(dst + q);
        #pragma unroll 4
        for (; q + 4 <= n; q += 4) { const uint32_t hi = *++ws; *wd++ = __funnelshift_r(lo, hi, sh); lo = hi; }
    }
    for (; q < n; ++q) dst[q] = src[q];
}
__device__ __forceinline__ void fill_dash(uint8_t* dst, long n) {
    long q = 0;
    while (q < n && ((uintptr_t)(dst + q) & 3)) dst[q++] = DASH;
    for (; q + 4 <= n; q += 4) *reinterpret_cast<uint32_t*>(dst + q) = 0xffffffffu;
    for (; q < n; ++q) dst[q] = DASH;
}

// scratch layout of one consensus candidate: best[L] | other reads' codes [otot] | rows[no][Ls] (4-byte aligned, Ls = align4(L)) | accept[no] | ... | table
struct Layout { uint8_t* best; uint8_t* oth; uint8_t* rows; uint8_t* acc; uint32_t Ls; };
__device__ __forceinline__ Layout cand_layout(const C& c, uint32_t ci, uint32_t L, uint32_t no) {
    Layout y; y.best = c.scr + (size_t)c.scr_off[ci] * 16; y.oth = y.best + L;
    y.rows = reinterpret_cast<uint8_t*>(((uintptr_t)(y.oth + c.plan_otot[ci]) + 3) & ~(uintptr_t)3); y.Ls = (L + 3u) & ~3u; y.acc = y.rows + (size_t)no * y.Ls;
    return y;
}
__device__ __forceinline__ uint8_t* cand_table(const C& c, uint32_t ci, uint32_t** keys, int** pos) {
    uint8_t* scr = c.scr + (size_t)c.scr_off[ci] * 16;
    uint8_t* end = scr + (size_t)c.scr_len[ci] * 16;
    *keys = reinterpret_cast<uint32_t*>(end - TAB * 8); *pos = reinterpret_cast<int*>(end - TAB * 4);
    return scr;
}

__global__ void __launch_bounds__(128) k_prep(C c) {
    __shared__ uint32_t s_cand;
    static const char CODE[17] = "=ACMGRSVTWYHKDBN";
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { const uint32_t q = atomicAdd(&c.work_ctr[2], 1u); const uint32_t nb = c.work_ctr[0], ns = c.work_ctr[1];
            s_cand = q < nb ? c.work_big[q] : (q < nb + ns ? c.work_small[q - nb] : 0xffffffffu); }
        __syncthreads();
        const uint32_t ci = s_cand; if (ci == 0xffffffffu) break;
        const uint32_t L = c.alt_len[ci];
        if ((unsigned long long)c.alt_off[ci] + L > c.alt_cap || (unsigned long long)c.scr_off[ci] + c.scr_len[ci] > c.scr_cap16) { if (threadIdx.x == 0) atomicAdd(&c.ctr->scratch_overflow, 1ULL); continue; }
        const snfb_cand cd = c.cand[ci];
        uint32_t* tk; int* tp; uint8_t* best = cand_table(c, ci, &tk, &tp);
        unpack_lead(c, cd.lead_off + c.plan_best[ci], best);
        const uint32_t no = c.plan_nother[ci];
        if (no == 0 || L == 0) { __syncthreads(); uint8_t* out = c.alt + c.alt_off[ci]; for (uint32_t h = threadIdx.x; h < L; h += blockDim.x) out[h] = (uint8_t)CODE[best[h]]; continue; }
        for (int i = threadIdx.x; i < TAB; i += blockDim.x) { tk[i] = 0xffffffffu; tp[i] = -1; }
        __syncthreads();
        const long skip = c.cfg.consensus_kmer_skip_base + (long)__dmul_rn((double)L, c.cfg.consensus_kmer_skip_seqlen_mult);
        for (long i = (long)threadIdx.x * skip; i < (long)L - 6; i += (long)blockDim.x * skip) {
            const uint32_t key = kmer6(best + i); uint32_t s = kslot(key);
            for (;;) { const uint32_t old = atomicCAS(&tk[s], 0xffffffffu, key); if (old == 0xffffffffu || old == key) break; s = (s + 1) & (TAB - 1); }
            if (atomicCAS(&tp[s], -1, (int)i) != -1) tp[s] = -2;
        }
    }
}

constexpr int ALIGN_WARPS = 4;
__global__ void __launch_bounds__(ALIGN_WARPS * 32, 5) k_align(C c) {
    __shared__ int h_i[ALIGN_WARPS][MAXHIT], h_j[ALIGN_WARPS][MAXHIT], h_cl[ALIGN_WARPS][MAXHIT], h_a[ALIGN_WARPS][MAXHIT], h_b[ALIGN_WARPS][MAXHIT];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int* hi = h_i[warp]; int* hj = h_j[warp]; int* hcl = h_cl[warp]; int* run_st = h_a[warp];   /* per-run identity sum */ int* run_len = h_b[warp];
    const int klen = 6;
    unsigned long long d_busy = 0, d_items = 0, d_max = 0, d_L = 0, d_Lo = 0, d_cL = 0, d_cLo = 0; long long d_t0 = 0; const long long d_start = clock64();
    unsigned long long d_ph[5] = { 0, 0, 0, 0, 0 }; long long d_pt = 0;
    #define PHASE(k) if (c.dbg) { const long long n_ = clock64(); d_ph[k] += (unsigned long long)(n_ - d_pt); d_pt = n_; }
    for (;;) {
        if (c.dbg && d_t0) { const unsigned long long dt = (unsigned long long)(clock64() - d_t0); d_busy += dt; if (dt > d_max) { d_max = dt; d_L = d_cL; d_Lo = d_cLo; } d_t0 = 0; }
        uint32_t q = 0; if (lane == 0) q = atomicAdd(&c.work_ctr[6], 1u);
        q = __shfl_sync(FULL, q, 0);
        const uint32_t nb = (uint32_t)min((unsigned long long)c.work_ctr[4], c.item_cap), ns = (uint32_t)min((unsigned long long)c.work_ctr[5], c.item_cap);
        if (q >= nb + ns) break;
        const C::Item it = q < nb ? c.items_big[q] : c.items_small[q - nb];
        const uint32_t ci = it.cand; const uint32_t L = c.alt_len[ci];
        if (c.dbg) { d_t0 = clock64(); d_pt = d_t0; ++d_items; d_cL = L; }
        if ((unsigned long long)c.scr_off[ci] + c.scr_len[ci] > c.scr_cap16) continue;
        const snfb_cand* cd = &c.cand[ci];
        uint32_t* t_key; int* t_pos; cand_table(c, ci, &t_key, &t_pos);
        const uint32_t no = c.plan_nother[ci];
        const Layout y = cand_layout(c, ci, L, no);
        uint8_t* best = y.best; uint8_t* acc = y.acc;
        const snfb_lead* l = &c.cand_leads[cd->lead_off + it.k];
        const long Lo = l->seq_len; uint8_t* rd = y.oth + it.rd_off; uint8_t* row = y.rows + (size_t)it.row * y.Ls;
        d_cLo = (unsigned long long)Lo;
        const long skip = c.cfg.consensus_kmer_skip_base + (long)__dmul_rn((double)L, c.cfg.consensus_kmer_skip_seqlen_mult);
        unpack_lead_warp(c, cd->lead_off + it.k, rd);
        __syncwarp();
        PHASE(0)
        // (1) anchor hits in j order (table lives in global scratch, L2 resident)
        int nh = 0;
        const long nk = Lo - klen > 0 ? (Lo - klen + skip - 1) / skip : 0;
        for (long kb = 0; kb < nk; kb += 32) {
            const long kk = kb + lane; const long j = kk * skip; int ai = -1;
            if (kk < nk) { const uint32_t key = kmer6(rd + j); uint32_t s = kslot(key);
                for (;;) { const uint32_t tk = t_key[s]; if (tk == 0xffffffffu) break; if (tk == key) { ai = t_pos[s]; break; } s = (s + 1) & (TAB - 1); }
                if (ai >= 0) { long d = ai - j; if (d < 0) d = -d; if (d > klen) ai = -1; } }
            const unsigned hm = __ballot_sync(FULL, ai >= 0);
            if (ai >= 0) { const int p = nh + __popc(hm & lanemask_lt()); if (p < MAXHIT) { hi[p] = ai; hj[p] = (int)j; } }
            nh += __popc(hm);
        }
        if (nh > MAXHIT) nh = MAXHIT;
        __syncwarp();
        PHASE(1)
        // (2) anchor automaton (consensus.py:306-338) in closed form: a hit is accepted iff its i exceeds every earlier hit's i
        //     (the accepted hits are the left-to-right maxima), and len(conseq) before accepted hit m is
        //     min(L, c0 + j[m-1] - j[0]) because every step appends min(j step, room left).  Compacted in place.
        int na = 0, pm = -1;
        for (int hb = 0; hb < nh; hb += 32) {
            const int h = hb + lane; const int vi = h < nh ? hi[h] : -1, vj = h < nh ? hj[h] : 0;
            int inc = vi;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc = max(inc, t); }
            int exc = __shfl_up_sync(FULL, inc, 1); if (lane == 0) exc = -1; exc = max(exc, pm);
            const bool accp = h < nh && vi > exc;
            const unsigned am = __ballot_sync(FULL, accp);
            __syncwarp();
            if (accp) { const int p = na + __popc(am & lanemask_lt()); hi[p] = vi; hj[p] = vj; }
            na += __popc(am); pm = max(pm, __shfl_sync(FULL, inc, 31));
            __syncwarp();
        }
        const long j0 = na ? hj[0] : 0, c0 = (na && j0 > 0) ? hi[0] : 0;
        // (3a) lane per segment: agreement with the best read along the diagonal decides copy / dash; a copied segment also
        //      gets its column identity (the bases it shares with the best read at the columns it lands on)
        long span = 0;
        for (int m = 1 + lane; m < na; m += 32) {
            const long li = hi[m - 1], lj = hj[m - 1], i = hi[m], j = hj[m];
            long cs = c0 + lj - j0; if (cs > (long)L) cs = (long)L;
            const long d = j - lj; long fwd_j = d; if (cs + fwd_j > (long)L) fwd_j = (long)L - cs;
            int st = -1;
            if (i - li == fwd_j && fwd_j > 0) {
                span += d;
                long nc = (long)L - 1 - li; if (nc > d) nc = d;                       // positions past the end of the best read never match
                const int mt = nc > 0 ? match_count(rd + lj + 1, best + li + 1, nc) : 0;
                // column identity of the copied bases.  Without drift (cs == li, nothing clamped) it is the same sum shifted by one
                // position, and both end positions are anchor k-mer bases that match by construction: reuse mt
                if (__ddiv_rn((double)mt, (double)d) >= 0.5) st = (cs == li && fwd_j == d && nc == d) ? mt : match_count(rd + lj, best + cs, fwd_j);
            }
            hcl[m] = st;
        }
        span = (long)__reduce_add_sync(FULL, (unsigned)span);
        __syncwarp();
        PHASE(2)
        // (3b) dash-free runs (= chains of copied segments) survive only with identity > 0.5 and more than 5 matches
        //      (consensus.py:343-360); decided on the segment list before anything is written.  A non-empty dashed segment
        //      ends a run; run ids are prefix counts of those, the per-run sums are accumulated in shared memory.
        {
            int* hr = hi;                                   // the anchor i positions are no longer needed
            for (int m = lane; m < na; m += 32) { run_st[m] = 0; run_len[m] = 0; }
            __syncwarp();
            int run_base = 0;
            for (int mb = 1; mb < na; mb += 32) {
                const int m = mb + lane; int st = -1; long len = 0;
                if (m < na) { const long lj = hj[m - 1]; long cs = c0 + lj - j0; if (cs > (long)L) cs = (long)L; len = hj[m] - lj; if (cs + len > (long)L) len = (long)L - cs; st = hcl[m]; }
                const unsigned bm = __ballot_sync(FULL, m < na && st < 0 && len > 0);
                const int rid = run_base + __popc(bm & lanemask_lt());
                __syncwarp();
                if (m < na) { hr[m] = rid; if (st >= 0) { atomicAdd(&run_st[rid], st); atomicAdd(&run_len[rid], (int)len); } }
                run_base += __popc(bm);
            }
            __syncwarp();
            for (int m = 1 + lane; m < na; m += 32) {
                if (hcl[m] < 0) continue;
                const int r = hr[m], ident = run_st[r];
                if (!(__ddiv_rn((double)ident, (double)run_len[r]) > 0.5 && ident > 5)) hcl[m] = -1;
            }
            __syncwarp();
        }
        PHASE(3)
        // (3c) write the row once
        for (long q2 = lane; q2 < c0; q2 += 32) row[q2] = DASH;
        for (int m = 1 + lane; m < na; m += 32) {
            const long lj = hj[m - 1]; long cs = c0 + lj - j0; if (cs > (long)L) cs = (long)L;
            long len = hj[m] - lj; if (cs + len > (long)L) len = (long)L - cs;
            if (len > 0) { if (hcl[m] >= 0) copy_bytes(row + cs, rd + lj, len); else fill_dash(row + cs, len); }
        }
        { long cl = 0; if (na) { cl = c0 + hj[na - 1] - j0; if (cl > (long)L) cl = (long)L; }
          for (long q2 = cl + lane; q2 < (long)L; q2 += 32) row[q2] = DASH; }
        if (lane == 0) acc[it.row] = __ddiv_rn((double)span, (double)L) > 0.2;
        __syncwarp();
        PHASE(4)
    }
    if (c.dbg && lane == 0) { unsigned long long* o = c.dbg + (size_t)(blockIdx.x * ALIGN_WARPS + warp) * 8;
        o[0] = d_busy; o[1] = (unsigned long long)(clock64() - d_start); o[2] = d_items; o[3] = d_ph[0]; o[4] = d_ph[1]; o[5] = d_ph[2]; o[6] = d_ph[3]; o[7] = d_ph[4]; }
    #undef PHASE
}

// column vote (consensus.py:365-380), one block per (candidate, 4096-column tile); every thread takes four adjacent columns
// (rows are 4-byte aligned with a stride of align4(L)), eight rows in flight
constexpr int VOTE_LIST = 256;
// sixteen 16-bit counters (one per base code) in four registers; the selects keep them out of local memory
__device__ __forceinline__ void vote_add(unsigned long long (&cnt)[4], uint32_t code) {
    const unsigned long long inc = 1ull << (16 * (code & 3u)); const uint32_t sel = (code >> 2) & 3u;
    #pragma unroll
    for (int k = 0; k < 4; ++k) cnt[k] += sel == (uint32_t)k ? inc : 0ull;
}
constexpr int VOTE_THREADS = 64;      // most candidates are a few hundred columns: small blocks, many of them
__global__ void __launch_bounds__(VOTE_THREADS) k_vote(C c) {
    __shared__ uint2 s_tile; __shared__ int s_nacc, s_nlist; __shared__ uint16_t s_rows[VOTE_LIST];
    static const char CODE[17] = "=ACMGRSVTWYHKDBN";
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) { const uint32_t q = atomicAdd(&c.work_ctr[9], 1u); s_tile = q < c.work_ctr[8] && q < c.tile_cap ? c.tiles[q] : make_uint2(0xffffffffu, 0); }
        __syncthreads();
        const uint32_t ci = s_tile.x; if (ci == 0xffffffffu) break;
        const uint32_t L = c.alt_len[ci], no = c.plan_nother[ci];
        if ((unsigned long long)c.alt_off[ci] + L > c.alt_cap || (unsigned long long)c.scr_off[ci] + c.scr_len[ci] > c.scr_cap16) continue;
        const Layout y = cand_layout(c, ci, L, no);
        const uint8_t* best = y.best; const uint8_t* rows = y.rows; const uint8_t* acc = y.acc; const uint32_t Ls = y.Ls;
        if (threadIdx.x < 32) {                         // the accepted rows, in order (the first VOTE_LIST rows go through the list)
            int cnt = 0, extra = 0; const uint32_t lim = no < (uint32_t)VOTE_LIST ? no : (uint32_t)VOTE_LIST;
            for (uint32_t r0 = 0; r0 < lim; r0 += 32) { const uint32_t r = r0 + threadIdx.x; const bool a = r < lim && acc[r]; const unsigned bm = __ballot_sync(FULL, a);
                if (a) s_rows[cnt + __popc(bm & lanemask_lt())] = (uint16_t)r; cnt += __popc(bm); }
            for (uint32_t r = lim + threadIdx.x; r < no; r += 32) extra += acc[r] ? 1 : 0;
            extra = (int)__reduce_add_sync(FULL, (unsigned)extra);
            if (threadIdx.x == 0) { s_nlist = cnt; s_nacc = cnt + extra; }
        }
        __syncthreads();
        const double maxal = (double)(1 + s_nacc); const int nlist = s_nlist;
        uint8_t* out = c.alt + c.alt_off[ci];
        const uint32_t h_end = min(L, (s_tile.y + 1u) * 4096u);
        for (uint32_t h = s_tile.y * 4096u + threadIdx.x * 4u; h < h_end; h += blockDim.x * 4u) {
            unsigned long long cnt[4][4]; int nal[4];
            #pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) { nal[b2] = 0; cnt[b2][0] = cnt[b2][1] = cnt[b2][2] = cnt[b2][3] = 0; }
            for (int r0 = 0; r0 < nlist; r0 += 8) {
                uint32_t cv[8];
                #pragma unroll
                for (int u = 0; u < 8; ++u) cv[u] = r0 + u < nlist ? *reinterpret_cast<const uint32_t*>(rows + (size_t)s_rows[r0 + u] * Ls + h) : 0xffffffffu;
                #pragma unroll
                for (int u = 0; u < 8; ++u) {
                    #pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) { const uint32_t cc = (cv[u] >> (8 * b2)) & 255u; if (cc != DASH) { vote_add(cnt[b2], cc); ++nal[b2]; } }
                }
            }
            for (uint32_t r = VOTE_LIST; r < no; ++r) if (acc[r]) {          // more reads than the list holds: never with the default bins
                const uint32_t v = *reinterpret_cast<const uint32_t*>(rows + (size_t)r * Ls + h);
                #pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) { const uint32_t cc = (v >> (8 * b2)) & 255u; if (cc != DASH) { vote_add(cnt[b2], cc); ++nal[b2]; } }
            }
            const uint32_t bw = *reinterpret_cast<const uint32_t*>(best + h);
            #pragma unroll
            for (int b2 = 0; b2 < 4; ++b2) {
                if (h + b2 >= h_end) break;
                const uint32_t bc = (bw >> (8 * b2)) & 255u; uint32_t res = bc;
                if (!(nal[b2] < 2 || __ddiv_rn((double)nal[b2], maxal) < 0.25)) {
                    vote_add(cnt[b2], bc);
                    int t0 = -1, t1 = -1, c0 = 0, nd = 0;
                    #pragma unroll
                    for (int code = 0; code < 16; ++code) { const int v = (int)((cnt[b2][code >> 2] >> (16 * (code & 3))) & 0xffff); if (!v) continue; ++nd; if (v > t0) { t1 = t0; t0 = v; c0 = code; } else if (v > t1) t1 = v; }
                    if (nd > 1 && t0 - t1 >= 3) res = (uint32_t)c0;
                }
                out[h + b2] = (uint8_t)CODE[res];
            }
        }
    }
}

}  // namespace consensus
