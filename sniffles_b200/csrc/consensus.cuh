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
// strided k-mer hits of one read are bounded by (L + 6) / skip + 1 (skip = 3 + L / 500): < 512 for every L, < 384 for L <= HEAVY_L.
// The light items run with the smaller per-warp hit arrays, i.e. with more warps per SM (they are latency bound).
constexpr int MAXHIT = 512, MAXHIT_LIGHT = 384; constexpr uint32_t HEAVY_L = 3000;

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
                const long long d = 2 * a + 3 * p; if (nm == 0 || d < bd) { bd = d; bi = k; } ++nm; tot += ((long long)l->seq_len + 7) & ~7ll; }
            if (nm > 0) {
                const uint32_t L = (uint32_t)c.cand_leads[cd->lead_off + bi].seq_len; al = L;
                const bool cons = (nm - 1 >= c.cfg.consensus_min_reads) && !c.cfg.no_consensus;
                c.plan_best[i] = (uint32_t)bi; c.plan_nother[i] = cons ? (uint32_t)(nm - 1) : 0u; c.plan_otot[i] = (uint32_t)(tot - ((L + 7u) & ~7u));
                // scratch: best codes | others' codes (every read in a slot of align8(len) bytes) | one row of align4(L) per other read | accept flags | anchor table; 16-byte units
                const unsigned long long bytes = cons ? (unsigned long long)tot + 8 + (unsigned long long)(nm - 1) * ((L + 3u) & ~3u) + (unsigned long long)(nm - 1) * 16 + 64 + 16 + TAB * 8 : (unsigned long long)L + 24;
                sl = (uint32_t)((bytes + 15) / 16);
                c.cand_rw[i].alt_len = (int)L;
                // work queue: the heavy tail (long insertions with many reads) is scheduled first
                const unsigned long long work = (unsigned long long)L * (unsigned long long)nm;
                if (work > 60000ull) c.work_big[atomicAdd(&c.work_ctr[0], 1u)] = (uint32_t)i; else c.work_small[atomicAdd(&c.work_ctr[1], 1u)] = (uint32_t)i;
                if (cons) {
                    // one work item per supporting read (heavy rows first) and one per 4096-column tile of the vote
                    const bool heavy = L > HEAVY_L; C::Item* dst = heavy ? c.items_big : c.items_small;
                    const uint32_t base = atomicAdd(&c.work_ctr[heavy ? 4 : 5], (uint32_t)(nm - 1));
                    uint32_t row = 0, ro = 0;
                    for (int k = 0; k < cd->lead_n; ++k) { const snfb_lead* l = &c.cand_leads[cd->lead_off + k]; if (!(l->flags & SNFB_LF_HAS_SEQ) || k == bi) continue;
                        if ((unsigned long long)base + row < c.item_cap) { C::Item it; it.cand = (uint32_t)i; it.k = (uint32_t)k; it.row = row; it.rd_off = ro; dst[base + row] = it; }
                        ++row; ro += ((uint32_t)l->seq_len + 7u) & ~7u; }
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
// A thread takes 8 consecutive bases per step: two aligned words of the 4-bit arena (the second only when the bases reach
// into it), nibbles swapped into little-endian order, one funnel shift to the first base, two spreads, one 8-byte store.
__device__ __forceinline__ uint32_t nib_swap(uint32_t w) { return ((w & 0x0f0f0f0fu) << 4) | ((w >> 4) & 0x0f0f0f0fu); }
__device__ __forceinline__ uint32_t spread4(uint32_t x) { const uint32_t t = (x | (x << 8)) & 0x00ff00ffu; return (t | (t << 4)) & 0x0f0f0f0fu; }
__device__ __forceinline__ uint2 unpack8(const uint8_t* __restrict__ sq, long long q, int n) {
    const uintptr_t A = (uintptr_t)(sq + (q >> 1));
    const uint32_t* w = reinterpret_cast<const uint32_t*>(A & ~(uintptr_t)3);
    const int qn = (int)((A & 3) << 1) | (int)(q & 1);
    const uint32_t lo = nib_swap(__ldg(w)), hi = qn + n > 8 ? nib_swap(__ldg(w + 1)) : 0u;
    const uint32_t x = __funnelshift_r(lo, hi, 4 * qn);
    return make_uint2(spread4(x & 0xffffu), spread4(x >> 16));
}
__device__ __forceinline__ void unpack_span(const uint8_t* __restrict__ sq, long long off, int len, uint8_t* __restrict__ dst, int tid, int nthr) {
    const int full = ((uintptr_t)dst & 7) == 0 ? (len & ~7) : 0;          // whole 8-base groups go out as aligned 8-byte stores, eight loads in flight
    int jb = tid * 8;
    for (; jb + 7 * nthr * 8 < full; jb += 8 * nthr * 8) {
        uint2 v[8];
        #pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = unpack8(sq, off + jb + u * nthr * 8, 8);
        #pragma unroll
        for (int u = 0; u < 8; ++u) *reinterpret_cast<uint2*>(dst + jb + u * nthr * 8) = v[u];
    }
    for (; jb < full; jb += nthr * 8) *reinterpret_cast<uint2*>(dst + jb) = unpack8(sq, off + jb, 8);
    for (; jb < len; jb += nthr * 8) {
        const int n = min(8, len - jb); const uint2 v = unpack8(sq, off + jb, n);
        for (int t = 0; t < n; ++t) dst[jb + t] = (uint8_t)(((t < 4 ? v.x : v.y) >> (8 * (t & 3))) & 15u);
    }
}
// unpack the (possibly merged) sequence of candidate lead `cl_index` as 4-bit codes, one byte per base, by `nthr` cooperating threads
__device__ inline void unpack_lead(const C& c, uint32_t cl_index, uint8_t* dst, int tid, int nthr) {
    const uint32_t plo = c.out_plo[cl_index], pn = c.out_pn[cl_index];
    long long o = 0;
    for (uint32_t p = 0; p < pn; ++p) {
        const uint32_t slot = c.ord[plo + p]; const snfb_lead* l = &c.kleads[slot];
        const uint8_t* sq = c.arena_off ? c.seq + (size_t)c.arena_off[slot] * 16 : c.seq + c.rec[l->rec].seq_off;
        unpack_span(sq, c.arena_off ? (l->seq_off & 1) : l->seq_off, l->seq_len, dst + o, tid, nthr);
        o += l->seq_len;
    }
}

__device__ __forceinline__ uint32_t kmer6(const uint8_t* s) { return (uint32_t)s[0] | ((uint32_t)s[1] << 4) | ((uint32_t)s[2] << 8) | ((uint32_t)s[3] << 12) | ((uint32_t)s[4] << 16) | ((uint32_t)s[5] << 20); }
// the same from an unaligned pointer with aligned 32-bit loads (reads at most 3 bytes past s + 5)
__device__ __forceinline__ uint32_t kmer6_u(const uint8_t* s) {
    const uintptr_t A = (uintptr_t)s; const uint32_t* w = reinterpret_cast<const uint32_t*>(A & ~(uintptr_t)3); const uint32_t sh = (uint32_t)(A & 3) * 8;
    const uint32_t l0 = w[0], l1 = w[1], l2 = sh > 16 ? w[2] : 0u;
    const uint32_t x0 = __funnelshift_r(l0, l1, sh), x1 = __funnelshift_r(l1, l2, sh);
    uint32_t t = x0 & 0x0f0f0f0fu; t = (t | (t >> 4)) & 0x00ff00ffu; t = (t | (t >> 8)) & 0xffffu;
    return t | ((x1 & 15u) << 16) | (((x1 >> 8) & 15u) << 20);
}
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

__device__ __forceinline__ uint32_t load_u32_unaligned(const uint8_t* p) {
    const uintptr_t A = (uintptr_t)p; const uint32_t* w = reinterpret_cast<const uint32_t*>(A & ~(uintptr_t)3);
    return __funnelshift_r(w[0], w[1], (uint32_t)(A & 3) * 8);
}
// match_count with the words of [0, n) dealt round robin to the G lanes of a group; the caller adds the lanes up
template <int G>
__device__ __forceinline__ int group_match(const uint8_t* a, const uint8_t* b, long n, int sub) {
    int mt = 0;
    #pragma unroll 2
    for (long q = 4L * sub; q < n; q += 4L * G) {
        const uint32_t x = load_u32_unaligned(a + q) ^ load_u32_unaligned(b + q);
        uint32_t eq = ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
        if (n - q < 4) eq &= (1u << (8 * (n - q))) - 1u;
        mt += __popc(eq);
    }
    return mt;
}
template <int G> __device__ __forceinline__ int group_sum(int v) {
    #pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// ---- cooperating groups of k_align: one warp (light items) or the whole block (heavy items).  `tid` order = element order.
struct WarpGrp {
    static constexpr int NTHR = 32; int tid;
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    // exclusive prefix sum over the group in tid order; tot = group total
    __device__ __forceinline__ int excl(int v, int& tot) const {
        int inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (tid >= o) inc += t; }
        tot = __shfl_sync(FULL, inc, 31); return inc - v;
    }
    // max over the threads before this one (-1 when none); tot = group max
    __device__ __forceinline__ int exclmax(int v, int& tot) const {
        int inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (tid >= o) inc = max(inc, t); }
        tot = __shfl_sync(FULL, inc, 31); int e = __shfl_up_sync(FULL, inc, 1); if (tid == 0) e = -1; return e;
    }
    __device__ __forceinline__ long sum(long v) const { return (long)__reduce_add_sync(FULL, (unsigned)v); }
};
template <int WARPS>
struct BlockGrp {
    static constexpr int NTHR = WARPS * 32; int tid; int* red;          // red: 2 * WARPS ints of shared memory
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ int excl(int v, int& tot) const {
        const int lane = tid & 31, w = tid >> 5; int inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) red[w] = inc;
        __syncthreads();
        int base = 0, all = 0;
        #pragma unroll
        for (int k = 0; k < WARPS; ++k) { const int t = red[k]; if (k < w) base += t; all += t; }
        __syncthreads();
        tot = all; return base + inc - v;
    }
    __device__ __forceinline__ int exclmax(int v, int& tot) const {
        const int lane = tid & 31, w = tid >> 5; int inc = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc = max(inc, t); }
        int e = __shfl_up_sync(FULL, inc, 1); if (lane == 0) e = -1;
        if (lane == 31) red[w] = inc;
        __syncthreads();
        int all = -1;
        #pragma unroll
        for (int k = 0; k < WARPS; ++k) { const int t = red[k]; if (k < w) e = max(e, t); all = max(all, t); }
        __syncthreads();
        tot = all; return e;
    }
    __device__ __forceinline__ long sum(long v) const {
        const int lane = tid & 31, w = tid >> 5; const int sv = (int)__reduce_add_sync(FULL, (unsigned)v);
        if (lane == 0) red[w] = sv;
        __syncthreads();
        int all = 0;
        #pragma unroll
        for (int k = 0; k < WARPS; ++k) all += red[k];
        __syncthreads();
        return (long)all;
    }
};

// (3a) of k_align with G lanes per segment (1: short segments, a lane walks its own; 8: long segments, coalesced words); `tid` of
// `nthr` cooperating threads (whole warps)
template <int G>
__device__ __forceinline__ long segments_pass(const int* hi, const int* hj, int* hcl, int na, long c0, long j0, long L, const uint8_t* rd, const uint8_t* best, int tid, int nthr) {
    const int PER = nthr / G; const int grp = tid / G, sub = tid % G; long span = 0;
    for (int mb = 1; mb < na; mb += PER) {
        const int m = mb + grp; const bool v = m < na;
        long li = 0, lj = 0, i = 0, j = 0; if (v) { li = hi[m - 1]; lj = hj[m - 1]; i = hi[m]; j = hj[m]; }
        long cs = c0 + lj - j0; if (cs > L) cs = L;
        const long d = j - lj; long fwd_j = d; if (cs + fwd_j > L) fwd_j = L - cs;
        const bool el = v && i - li == fwd_j && fwd_j > 0;
        long nc = 0; if (el) { nc = L - 1 - li; if (nc > d) nc = d; if (nc < 0) nc = 0; }      // positions past the end of the best read never match
        const int mt = group_sum<G>(group_match<G>(rd + lj + 1, best + li + 1, nc, sub));
        // column identity of the copied bases.  Without drift (cs == li, nothing clamped) it is the same sum shifted by one
        // position, and both end positions are anchor k-mer bases that match by construction: reuse mt
        int st = -1; bool second = false;
        if (el) { if (sub == 0) span += d; if (__ddiv_rn((double)mt, (double)d) >= 0.5) { if (cs == li && fwd_j == d && nc == d) st = mt; else second = true; } }
        const int m2 = group_sum<G>(group_match<G>(rd + lj, best + cs, second ? fwd_j : 0, sub));
        if (second) st = m2;
        if (v && sub == 0) hcl[m] = st;
    }
    return span;
}

// scratch layout of one consensus candidate: best[align8(L)] | other reads' codes [otot] | rows[no][Ls] (Ls = align4(L)) | accept[no] | ... | table
struct Layout { uint8_t* best; uint8_t* oth; uint8_t* rows; uint8_t* acc; uint32_t Ls; };
__device__ __forceinline__ Layout cand_layout(const C& c, uint32_t ci, uint32_t L, uint32_t no) {
    Layout y; y.best = c.scr + (size_t)c.scr_off[ci] * 16; y.oth = y.best + ((L + 7u) & ~7u);
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
        unpack_lead(c, cd.lead_off + c.plan_best[ci], best, threadIdx.x, blockDim.x);
        const uint32_t no = c.plan_nother[ci];
        if (no == 0 || L == 0) { __syncthreads(); uint8_t* out = c.alt + c.alt_off[ci]; for (uint32_t h = threadIdx.x; h < L; h += blockDim.x) out[h] = (uint8_t)CODE[best[h]]; continue; }
        for (int i = threadIdx.x; i < TAB; i += blockDim.x) { tk[i] = 0xffffffffu; tp[i] = -1; }
        __syncthreads();
        const long skip = c.cfg.consensus_kmer_skip_base + (long)__dmul_rn((double)L, c.cfg.consensus_kmer_skip_seqlen_mult);
        for (long i = (long)threadIdx.x * skip; i < (long)L - 6; i += (long)blockDim.x * skip) {
            const uint32_t key = kmer6_u(best + i); uint32_t s = kslot(key);
            for (;;) { const uint32_t old = atomicCAS(&tk[s], 0xffffffffu, key); if (old == 0xffffffffu || old == key) break; s = (s + 1) & (TAB - 1); }
            if (atomicCAS(&tp[s], -1, (int)i) != -1) tp[s] = -2;
        }
    }
}

// one (candidate, supporting read) item by the cooperating group g; hi / hj / hcl / run_st / run_len hold `cap` ints each
template <typename G>
__device__ __forceinline__ void align_item(const C& c, const C::Item it, const G g, int* hi, int* hj, int* hcl, int* run_st, int* run_len, int cap, unsigned long long* d_ph) {
    const int tid = g.tid; constexpr int NT = G::NTHR; const int klen = 6;
    long long d_pt = c.dbg ? clock64() : 0;
    #define PHASE(k) if (c.dbg) { const long long n_ = clock64(); d_ph[k] += (unsigned long long)(n_ - d_pt); d_pt = n_; }
    const uint32_t ci = it.cand; const uint32_t L = c.alt_len[ci];
    if ((unsigned long long)c.scr_off[ci] + c.scr_len[ci] > c.scr_cap16) return;
    const snfb_cand* cd = &c.cand[ci];
    uint32_t* t_key; int* t_pos; cand_table(c, ci, &t_key, &t_pos);
    const uint32_t no = c.plan_nother[ci];
    const Layout y = cand_layout(c, ci, L, no);
    uint8_t* best = y.best; uint8_t* acc = y.acc;
    const snfb_lead* l = &c.cand_leads[cd->lead_off + it.k];
    const long Lo = l->seq_len; uint8_t* rd = y.oth + it.rd_off; uint8_t* row = y.rows + (size_t)it.row * y.Ls;
    const long skip = c.cfg.consensus_kmer_skip_base + (long)__dmul_rn((double)L, c.cfg.consensus_kmer_skip_seqlen_mult);
    unpack_lead(c, cd->lead_off + it.k, rd, tid, NT);
    g.sync();
    PHASE(0)
    // (1) anchor hits in j order (table lives in global scratch, L2 resident); four k-mers per thread in flight
    int nh = 0;
    const long nk = Lo - klen > 0 ? (Lo - klen + skip - 1) / skip : 0;
    for (long kb = 0; kb < nk; kb += 4 * NT) {
        uint32_t key[4], sl[4], tk[4]; int ai[4];
        #pragma unroll
        for (int u = 0; u < 4; ++u) { const long kk = kb + u * NT + tid; key[u] = kk < nk ? kmer6_u(rd + kk * skip) : 0u; sl[u] = kslot(key[u]); }
        #pragma unroll
        for (int u = 0; u < 4; ++u) tk[u] = kb + u * NT + tid < nk ? t_key[sl[u]] : 0xffffffffu;
        #pragma unroll
        for (int u = 0; u < 4; ++u) while (tk[u] != 0xffffffffu && tk[u] != key[u]) { sl[u] = (sl[u] + 1) & (TAB - 1); tk[u] = t_key[sl[u]]; }
        #pragma unroll
        for (int u = 0; u < 4; ++u) ai[u] = tk[u] != 0xffffffffu ? t_pos[sl[u]] : -1;
        #pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (kb + (long)u * NT >= nk) break;
            const long j = (kb + u * NT + tid) * skip;
            if (ai[u] >= 0) { long d = ai[u] - j; if (d < 0) d = -d; if (d > klen) ai[u] = -1; }
            int tot; const int off = g.excl(ai[u] >= 0 ? 1 : 0, tot);
            if (ai[u] >= 0) { const int p = nh + off; if (p < cap) { hi[p] = ai[u]; hj[p] = (int)j; } }
            nh += tot;
        }
    }
    if (nh > cap) nh = cap;
    g.sync();
    PHASE(1)
    // (2) anchor automaton (consensus.py:306-338) in closed form: a hit is accepted iff its i exceeds every earlier hit's i
    //     (the accepted hits are the left-to-right maxima), and len(conseq) before accepted hit m is
    //     min(L, c0 + j[m-1] - j[0]) because every step appends min(j step, room left).  Compacted in place.
    int na = 0, pm = -1;
    for (int hb = 0; hb < nh; hb += NT) {
        const int h = hb + tid; const int vi = h < nh ? hi[h] : -1, vj = h < nh ? hj[h] : 0;
        int tmax; const int exc = max(g.exclmax(vi, tmax), pm);
        const bool accp = h < nh && vi > exc;
        int tot; const int off = g.excl(accp ? 1 : 0, tot);
        g.sync();
        if (accp) { hi[na + off] = vi; hj[na + off] = vj; }
        na += tot; pm = max(pm, tmax);
        g.sync();
    }
    const long j0 = na ? hj[0] : 0, c0 = (na && j0 > 0) ? hi[0] : 0;
    // (3a) agreement with the best read along the diagonal decides copy / dash; a copied segment also gets its column identity
    //      (the bases it shares with the best read at the columns it lands on)
    long span = skip > 12 ? segments_pass<8>(hi, hj, hcl, na, c0, j0, (long)L, rd, best, tid, NT) : segments_pass<1>(hi, hj, hcl, na, c0, j0, (long)L, rd, best, tid, NT);
    span = g.sum(span);
    g.sync();
    PHASE(2)
    // (3b) dash-free runs (= chains of copied segments) survive only with identity > 0.5 and more than 5 matches
    //      (consensus.py:343-360); decided on the segment list before anything is written.  A non-empty dashed segment
    //      ends a run; run ids are prefix counts of those, the per-run sums are accumulated in shared memory.
    {
        int* hr = hi;                                   // the anchor i positions are no longer needed
        for (int m = tid; m < na; m += NT) { run_st[m] = 0; run_len[m] = 0; }
        g.sync();
        int run_base = 0;
        for (int mb = 1; mb < na; mb += NT) {
            const int m = mb + tid; int st = -1; long len = 0;
            if (m < na) { const long lj = hj[m - 1]; long cs = c0 + lj - j0; if (cs > (long)L) cs = (long)L; len = hj[m] - lj; if (cs + len > (long)L) len = (long)L - cs; st = hcl[m]; }
            int tot; const int rid = run_base + g.excl((m < na && st < 0 && len > 0) ? 1 : 0, tot);
            g.sync();
            if (m < na) { hr[m] = rid; if (st >= 0) { atomicAdd(&run_st[rid], st); atomicAdd(&run_len[rid], (int)len); } }
            run_base += tot;
        }
        g.sync();
        for (int m = 1 + tid; m < na; m += NT) {
            if (hcl[m] < 0) continue;
            const int r = hr[m], ident = run_st[r];
            if (!(__ddiv_rn((double)ident, (double)run_len[r]) > 0.5 && ident > 5)) hcl[m] = -1;
        }
        g.sync();
    }
    PHASE(3)
    // (3c) the row: every copied segment lands at row[cs + t] = rd[lj + t] with cs - lj = c0 - j0 for all of them, so the row is
    //      one shifted copy of the read between the first and the last anchor (coalesced words), dashes outside, and the
    //      rejected segments dashed afterwards
    {
        long cl = 0; if (na) { cl = c0 + hj[na - 1] - j0; if (cl > (long)L) cl = (long)L; }
        const uint8_t* src0 = rd + (j0 - c0);
        #pragma unroll 4
        for (long X = 4L * tid; X < (long)L; X += 4L * NT) {
            uint32_t v = 0xffffffffu;
            if (X + 4 > c0 && X < cl) {
                v = load_u32_unaligned(src0 + X);
                if (X < c0) v |= (1u << (8 * (c0 - X))) - 1u;                         // bytes before c0
                if (X + 4 > cl) v |= ~((1u << (8 * (cl - X))) - 1u);                    // bytes from cl on
            }
            *reinterpret_cast<uint32_t*>(row + X) = v;
        }
        g.sync();
        for (int m = 1 + tid; m < na; m += NT) {
            if (hcl[m] >= 0) continue;
            const long lj = hj[m - 1]; long cs = c0 + lj - j0; if (cs > (long)L) cs = (long)L;
            long len = hj[m] - lj; if (cs + len > (long)L) len = (long)L - cs;
            if (len > 0) fill_dash(row + cs, len);
        }
    }
    if (tid == 0) acc[it.row] = __ddiv_rn((double)span, (double)L) > 0.2;
    g.sync();
    PHASE(4)
    #undef PHASE
}

// Heavy items (long insertions) first, the whole block on one item (the longest insertion's reads are the tail of the step);
// then the light items, one per warp.  Per-warp hit arrays of MAXHIT_LIGHT entries; a block item uses all of them as one array.
constexpr int ALIGN_WARPS = 4;
__global__ void __launch_bounds__(ALIGN_WARPS * 32, 7) k_align(C c) {
    __shared__ int sm[5][ALIGN_WARPS * MAXHIT_LIGHT]; __shared__ int s_red[2 * ALIGN_WARPS]; __shared__ uint32_t s_q;
    static_assert(ALIGN_WARPS * MAXHIT_LIGHT >= MAXHIT, "a block item needs MAXHIT entries");
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    unsigned long long d_busy = 0, d_items = 0, d_ph[5] = { 0, 0, 0, 0, 0 }; const long long d_start = c.dbg ? clock64() : 0;
    const uint32_t nb = (uint32_t)min((unsigned long long)c.work_ctr[4], c.item_cap), ns = (uint32_t)min((unsigned long long)c.work_ctr[5], c.item_cap);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_q = atomicAdd(&c.work_ctr[6], 1u);
        __syncthreads();
        const uint32_t q = s_q; if (q >= nb) break;
        const long long t0 = c.dbg ? clock64() : 0;
        BlockGrp<ALIGN_WARPS> g; g.tid = threadIdx.x; g.red = s_red;
        align_item(c, c.items_big[q], g, sm[0], sm[1], sm[2], sm[3], sm[4], ALIGN_WARPS * MAXHIT_LIGHT, d_ph);
        if (c.dbg) { d_busy += (unsigned long long)(clock64() - t0); if (warp == 0) ++d_items; }
    }
    __syncthreads();
    for (;;) {
        uint32_t q = 0; if (lane == 0) q = atomicAdd(&c.work_ctr[7], 1u);
        q = __shfl_sync(FULL, q, 0);
        if (q >= ns) break;
        const long long t0 = c.dbg ? clock64() : 0;
        WarpGrp g; g.tid = lane; const int o = warp * MAXHIT_LIGHT;
        align_item(c, c.items_small[q], g, sm[0] + o, sm[1] + o, sm[2] + o, sm[3] + o, sm[4] + o, MAXHIT_LIGHT, d_ph);
        if (c.dbg) { d_busy += (unsigned long long)(clock64() - t0); ++d_items; }
    }
    if (c.dbg && lane == 0) { unsigned long long* o = c.dbg + (size_t)(blockIdx.x * ALIGN_WARPS + warp) * 8;
        o[0] = d_busy; o[1] = (unsigned long long)(clock64() - d_start); o[2] = d_items; o[3] = d_ph[0]; o[4] = d_ph[1]; o[5] = d_ph[2]; o[6] = d_ph[3]; o[7] = d_ph[4]; }
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
            const uint32_t bw = *reinterpret_cast<const uint32_t*>(best + h);
            // fast path: the four columns of a word are counted byte-parallel.  A, C, G, T are the one-hot codes 1, 2, 4, 8, so bit k
            // of a byte is that base's vote; dashes (0xff) are cleared first; anything else (ambiguity codes, '=') shows up as a byte
            // with two bits set or as a column whose votes and dashes do not add up to the rows, and sends the word down the exact path
            if (nlist == s_nacc && nlist < 255) {
                uint32_t cA = 0, cC = 0, cG = 0, cT = 0, cD = 0, multi = 0;
                for (int r0 = 0; r0 < nlist; r0 += 8) {
                    uint32_t cv[8];
                    #pragma unroll
                    for (int u = 0; u < 8; ++u) cv[u] = r0 + u < nlist ? *reinterpret_cast<const uint32_t*>(rows + (size_t)s_rows[r0 + u] * Ls + h) : 0xffffffffu;
                    #pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const uint32_t hb = cv[u] & 0x80808080u, d1 = hb >> 7, t = cv[u] & ~(hb | (hb - d1));
                        cD += d1; multi |= ((t | 0x10101010u) - 0x01010101u) & t;
                        cA += t & 0x01010101u; cC += (t >> 1) & 0x01010101u; cG += (t >> 2) & 0x01010101u; cT += (t >> 3) & 0x01010101u;
                    }
                }
                // padding rows of the last group of eight counted as dashes
                const uint32_t padded = (uint32_t)((nlist + 7) & ~7);
                const uint32_t tot = cA + cC + cG + cT + cD;                      // per byte: rows accounted for (<= 255 + 7 would overflow: bounded by padded <= 256 only when nlist <= 248)
                const bool best_ok = (((bw | 0x10101010u) - 0x01010101u) & bw) == 0 && ((bw - 0x01010101u) & ~bw & 0x80808080u) == 0;
                if (multi == 0 && best_ok && padded <= 248 && tot == padded * 0x01010101u) {
                    #pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) {
                        if (h + b2 >= h_end) break;
                        const uint32_t bc = (bw >> (8 * b2)) & 255u; uint32_t res = bc;
                        const int nal = (int)padded - (int)((cD >> (8 * b2)) & 255u);
                        if (!(nal < 2 || __ddiv_rn((double)nal, maxal) < 0.25)) {
                            int v4[4] = { (int)((cA >> (8 * b2)) & 255u), (int)((cC >> (8 * b2)) & 255u), (int)((cG >> (8 * b2)) & 255u), (int)((cT >> (8 * b2)) & 255u) };
                            int t0 = -1, t1 = -1, c0 = 0, nd = 0;
                            #pragma unroll
                            for (int k = 0; k < 4; ++k) { const int v = v4[k] + ((bc >> k) & 1u); if (!v) continue; ++nd; if (v > t0) { t1 = t0; t0 = v; c0 = 1 << k; } else if (v > t1) t1 = v; }
                            if (nd > 1 && t0 - t1 >= 3) res = (uint32_t)c0;
                        }
                        out[h + b2] = (uint8_t)CODE[res];
                    }
                    continue;
                }
            }
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
