// combine.cuh — SURVEY §8(f)1: the grouping of multi-sample combine mode.
//   cluster.resolve_block_groups      cluster.py:356-390   (greedy nearest-group assignment, candidates by support)
//   SVGroup.from_candidate/add_candidate  sv.py:265-321    (running means, updated with the reference's operation order)
//   CombineTask.execute, the chunk loop   parallel.py:518-563 (coverage of non-included samples, keep / call split)
// The grouping is sequential along one (task, svtype) chain — the groups kept at the end of a chunk are the first groups the next chunk
// sees, across blocks — and independent between chains: one warp per chain, lanes over the active groups (distance + first-minimum
// reduction) and over the samples (coverage update).  All float decisions are in double with explicit _rn operations, in the
// reference's order.
#pragma once
#include "common.cuh"

namespace combine {

struct P {
    const snfb_combine_chain* chains; const snfb_combine_chunk* chunks; uint32_t n_chain, n_chunk, n_cand, n_samples, words;
    const int32_t* pos; const int32_t* svlen; const uint32_t* sample; const int32_t* mate_contig; const int32_t* mate_pos;
    const long long* block_start; const int32_t* cov; int bins_per_block, cov_binsize;
    int combine_match, combine_match_max, cluster_merge_bnd, separate_intra, overlap_abs;
    // per group slot (slot = chain.cand_off + local group id; a chain never has more groups than candidates)
    double* g_pos; double* g_len; double* g_mate; uint32_t* g_n; int32_t* g_mc; uint32_t* g_incl;   // g_incl: [slot][words] sample bitset
    uint32_t* act;                                 // [n_cand] active group list of the chain, at chain.cand_off
    uint32_t* cand_group; int32_t* emit_chunk; uint32_t* emit_ord; int32_t* cov_non;
    unsigned int* next_chain;
    // group.align_call (sv.py:282-292): pctseq > 0 switches it on.  alt: the candidates' ALT strings; g_first: the candidate that opened the group;
    // ex_stamp[g] == c + 1: group g failed the alignment test for candidate c; hs: per warp, max_alt bytes of carries between passes
    double pctseq; const uint8_t* alt; const unsigned long long* alt_off; const uint32_t* alt_len; uint32_t* g_first; uint32_t* ex_stamp; int8_t* hs; uint32_t max_alt;
};

// Levenshtein distance (edlib.align(a, b) defaults: global alignment, task "distance") by one warp: Myers' bit-vector recurrence in 64-row blocks
// (Hyyro's block formulation, the one edlib implements), lane = block, the text streamed through the lanes as a wavefront (lane l works on column
// s - l at step s, its horizontal carry-in is lane l-1's carry-out of the step before); patterns of more than 32 blocks take several passes,
// the carries of a pass's last block wait in hs[].  The shorter string is the pattern.  The result is exact for any byte strings.
__device__ inline int edit_distance_warp(const uint8_t* a, int m, const uint8_t* b, int n, int8_t* hs) {
    const int lane = lane_id();
    if (m > n) { const uint8_t* t = a; a = b; b = t; const int ti = m; m = n; n = ti; }
    if (m == 0) return n;
    const int W = (m + 63) / 64; int score = m;
    for (int p0 = 0; p0 < W; p0 += 32) {
        const int blk = p0 + lane; const bool vb = blk < W; const bool last = blk == W - 1;
        unsigned long long eq[5] = { 0, 0, 0, 0, 0 };            // A C G T N
        const int rows = vb ? min(64, m - blk * 64) : 0;
        for (int k = 0; k < rows; ++k) { const uint8_t ch = a[blk * 64 + k]; const int idx = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : ch == 'N' ? 4 : -1;
            #pragma unroll
            for (int q = 0; q < 5; ++q) if (idx == q) eq[q] |= 1ull << k; }
        unsigned long long Pv = ~0ull, Mv = 0ull; int hout = 0;
        const unsigned long long top = last ? 1ull << ((m - 1) & 63) : 1ull << 63;
        // the text (and, after the first pass, the carries of the previous pass) reach the lanes through registers: lane l holds column
        // base + l of the current 32-column group and of the one before; the next group is loaded while this one is worked on
        int cprev = 0, ccur = lane < n ? (int)b[lane] : 0, hcur = (p0 && lane < n) ? (int)hs[lane] : 0;
        for (int base = 0; base < n + 31; base += 32) {
            const int nb = base + 32 + lane;
            const int cnext = nb < n ? (int)b[nb] : 0, hnext = (p0 && nb < n) ? (int)hs[nb] : 0;
            #pragma unroll 4
            for (int t = 0; t < 32; ++t) {
                const int s = base + t;
                const int hprev = __shfl_up_sync(FULL, hout, 1);
                const int j = s - lane;
                const int src = (t - lane) & 31;                                   // lane holding column j in its group
                const int c_a = __shfl_sync(FULL, ccur, src), c_b = __shfl_sync(FULL, cprev, src);
                const int h0 = __shfl_sync(FULL, hcur, t);                         // lane 0 works on column s
                const bool act = vb && j >= 0 && j < n;
                if (act) {
                    const uint8_t c = (uint8_t)(t >= lane ? c_a : c_b);
                    unsigned long long Eq;
                    if (c == 'A') Eq = eq[0]; else if (c == 'C') Eq = eq[1]; else if (c == 'G') Eq = eq[2]; else if (c == 'T') Eq = eq[3]; else if (c == 'N') Eq = eq[4];
                    else { Eq = 0; for (int k = 0; k < rows; ++k) if (a[blk * 64 + k] == c) Eq |= 1ull << k; }
                    const int hin = lane == 0 ? (p0 == 0 ? 1 : h0) : hprev;
                    const unsigned long long Xv = Eq | Mv;
                    if (hin < 0) Eq |= 1ull;
                    const unsigned long long Xh = (((Eq & Pv) + Pv) ^ Pv) | Eq;
                    unsigned long long Ph = Mv | ~(Xh | Pv), Mh = Pv & Xh;
                    hout = (Ph & top) ? 1 : ((Mh & top) ? -1 : 0);
                    Ph <<= 1; Mh <<= 1;
                    if (hin < 0) Mh |= 1ull; else if (hin > 0) Ph |= 1ull;
                    Pv = Mh | ~(Xv | Ph); Mv = Ph & Xv;
                    if (last) score += hout; else if (lane == 31) hs[j] = (int8_t)hout;
                }
            }
            cprev = ccur; ccur = cnext; hcur = hnext;
        }
        __syncwarp();
    }
    return __shfl_sync(FULL, score, (W - 1) & 31);
}


__global__ void __launch_bounds__(128) k_combine(const P p) {
    const int lane = lane_id();
    for (;;) {
        uint32_t ci = 0; if (lane == 0) ci = atomicAdd(p.next_chain, 1u);
        ci = __shfl_sync(FULL, ci, 0);
        if (ci >= p.n_chain) break;
        const snfb_combine_chain ch = p.chains[ci];
        uint32_t* act = p.act + ch.cand_off; uint32_t n_act = 0, n_groups = 0;
        const uint32_t W = p.words;
        for (uint32_t k = 0; k < ch.n_chunk; ++k) {
            const snfb_combine_chunk ck = p.chunks[ch.chunk_off + k];
            // ---- resolve_block_groups over the chunk's candidates (already in support order)
            for (uint32_t c = ck.cand_off; c < ck.cand_off + ck.n_cand; ++c) {
                const int cpos = p.pos[c], clen = p.svlen[c]; const uint32_t smp = p.sample[c];
                const int cmc = ch.is_bnd ? p.mate_contig[c] : 0, cmp = ch.is_bnd ? p.mate_pos[c] : 0;
                const double alen = (double)(clen < 0 ? -(long long)clen : (long long)clen);
                double bd; uint32_t bi;
                for (;;) {
                    bd = __longlong_as_double(0x7ff0000000000000ll); bi = 0xffffffffu;
                    for (uint32_t a0 = 0; a0 < n_act; a0 += 32) {
                        const uint32_t a = a0 + lane;
                        if (a < n_act) {
                            const uint32_t g = act[a]; const double gp = p.g_pos[g];
                            double dist; bool ok;
                            if (ch.is_bnd) {
                                dist = __dadd_rn(fabs(__dsub_rn(gp, (double)cpos)), fabs(__dsub_rn(p.g_mate[g], (double)cmp)));
                                ok = dist <= (double)(p.cluster_merge_bnd * 2) && p.g_mc[g] == cmc;
                            } else {
                                const double gl = fabs(p.g_len[g]);
                                dist = __dadd_rn(fabs(__dsub_rn(gp, (double)cpos)), fabs(__dsub_rn(gl, alen)));
                                const double minlen = gl < alen ? gl : alen;
                                ok = minlen > 0.0 && dist <= __dmul_rn((double)p.combine_match, __dsqrt_rn(minlen)) && dist <= (double)p.combine_match_max && p.ex_stamp[g] != c + 1u;
                            }
                            if (ok && dist < bd && (!p.separate_intra || !((p.g_incl[(size_t)g * W + (smp >> 5)] >> (smp & 31)) & 1u))) { bd = dist; bi = a; }
                        }
                    }
                    // first minimum in list order: smallest distance, then smallest list index
                    #pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double od = __shfl_xor_sync(FULL, bd, o); const uint32_t oi = __shfl_xor_sync(FULL, bi, o);
                        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
                    }
                    if (bi == 0xffffffffu || ch.is_bnd || !(p.pctseq != 0.0)) break;
                    // the nearest eligible group must also align (only groups that pass can become the best one, so testing them nearest first is the reference's result)
                    const uint32_t g = act[bi], fc = p.g_first[g];
                    const int d = edit_distance_warp(p.alt + p.alt_off[fc], (int)p.alt_len[fc], p.alt + p.alt_off[c], (int)p.alt_len[c], p.hs + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * p.max_alt);
                    const double lm = p.g_len[g];
                    if (__ddiv_rn(__dsub_rn(lm, (double)d), lm) > p.pctseq) break;
                    if (lane == 0) p.ex_stamp[g] = c + 1u;
                    __syncwarp();
                }
                uint32_t g;
                if (bi == 0xffffffffu) {                          // SVGroup.from_candidate
                    g = ch.cand_off + n_groups;
                    if (lane == 0) {
                        p.g_pos[g] = (double)cpos; p.g_len[g] = alen; p.g_mate[g] = (double)cmp; p.g_mc[g] = cmc; p.g_n[g] = 1u; act[n_act] = g; p.g_first[g] = c; p.ex_stamp[g] = 0u;
                        p.emit_chunk[g] = (int32_t)p.n_chunk; p.emit_ord[g] = 0u;
                    }
                    for (uint32_t w = lane; w < W; w += 32) p.g_incl[(size_t)g * W + w] = (w == (smp >> 5)) ? (1u << (smp & 31)) : 0u;
                    for (uint32_t s = lane; s < p.n_samples; s += 32) p.cov_non[(size_t)g * p.n_samples + s] = -1;
                    ++n_groups; ++n_act;
                } else {                                          // SVGroup.add_candidate
                    g = act[bi];
                    if (lane == 0) {
                        const uint32_t n = p.g_n[g]; const double dn = (double)n, dn1 = (double)(n + 1u);
                        p.g_pos[g] = __ddiv_rn(__dadd_rn(__dmul_rn(p.g_pos[g], dn), (double)cpos), dn1);
                        p.g_len[g] = __ddiv_rn(__dadd_rn(__dmul_rn(p.g_len[g], dn), alen), dn1);
                        if (ch.is_bnd) p.g_mate[g] = __ddiv_rn(__dadd_rn(__dmul_rn(p.g_mate[g], dn), (double)cmp), dn1);
                        p.g_n[g] = n + 1u;
                        p.g_incl[(size_t)g * W + (smp >> 5)] |= 1u << (smp & 31);
                    }
                }
                if (lane == 0) p.cand_group[c] = g;
                __syncwarp();
            }
            // ---- end of chunk: coverage of the samples a group does not include, then keep / call
            const double lim = fmax(__dmul_rn((double)ck.size, 0.5), (double)p.overlap_abs);
            uint32_t n_keep = 0, n_call = 0;
            for (uint32_t a0 = 0; a0 < n_act; a0 += 32) {
                const uint32_t a = a0 + lane; const bool v = a < n_act;
                const uint32_t g = v ? act[a] : 0u; const double gp = v ? p.g_pos[g] : 0.0;
                // coverage: the lanes of the warp take the samples of one group at a time
                for (uint32_t j = 0; j < 32u && a0 + j < n_act; ++j) {
                    const uint32_t gj = __shfl_sync(FULL, g, j); const double pj = __shfl_sync(FULL, gp, j);
                    const long long cb = (long long)__ddiv_rn(pj, (double)p.cov_binsize) * p.cov_binsize;
                    long long kbin = -1;
                    if (ck.cov_block >= 0) { const long long off = cb - p.block_start[ck.cov_block]; if (off >= 0 && off < (long long)p.bins_per_block * p.cov_binsize) kbin = off / p.cov_binsize; }
                    for (uint32_t s = lane; s < p.n_samples; s += 32) {
                        if ((p.g_incl[(size_t)gj * W + (s >> 5)] >> (s & 31)) & 1u) continue;
                        int cv = 0;
                        if (kbin >= 0) { const int t = p.cov[((size_t)ck.cov_block * p.n_samples + s) * p.bins_per_block + kbin]; if (t >= 0) cv = t; }
                        int32_t* d = &p.cov_non[(size_t)gj * p.n_samples + s]; if (cv > *d) *d = cv;
                    }
                }
                const bool keep = v && fabs(__dsub_rn(gp, (double)ck.curr_bin)) < lim;
                const unsigned km = __ballot_sync(FULL, keep), cm = __ballot_sync(FULL, v && !keep);
                __syncwarp();
                if (keep) act[n_keep + __popc(km & lanemask_lt())] = g;          // compaction in place: n_keep + rank <= a
                else if (v) { p.emit_chunk[g] = (int32_t)(ch.chunk_off + k); p.emit_ord[g] = n_call + __popc(cm & lanemask_lt()); }
                n_keep += __popc(km); n_call += __popc(cm);
                __syncwarp();
            }
            n_act = n_keep;
        }
        // groups still kept at the end of the chain are called last, in list order (parallel.py:565-566)
        for (uint32_t a = lane; a < n_act; a += 32) { const uint32_t g = act[a]; p.emit_chunk[g] = (int32_t)p.n_chunk; p.emit_ord[g] = a; }
        for (uint32_t g = ch.cand_off + n_groups + lane; g < ch.cand_off + ch.n_cand; g += 32) p.emit_chunk[g] = -1;     // unused slots
        __syncwarp();
    }
}

// self-check: one warp per pair
__global__ void k_edit_selftest(const uint8_t* bytes, const unsigned long long* a_off, const uint32_t* a_len, const unsigned long long* b_off, const uint32_t* b_len, uint32_t n_pairs, int8_t* hs, uint32_t max_len, int* out) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = w; i < n_pairs; i += nw) {
        const int d = edit_distance_warp(bytes + a_off[i], (int)a_len[i], bytes + b_off[i], (int)b_len[i], hs + (size_t)w * max_len);
        if (lane_id() == 0) out[i] = d;
        __syncwarp();
    }
}

}  // namespace combine
