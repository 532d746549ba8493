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
};

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
                double bd = __longlong_as_double(0x7ff0000000000000ll); uint32_t bi = 0xffffffffu;
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
                            ok = minlen > 0.0 && dist <= __dmul_rn((double)p.combine_match, __dsqrt_rn(minlen)) && dist <= (double)p.combine_match_max;
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
                uint32_t g;
                if (bi == 0xffffffffu) {                          // SVGroup.from_candidate
                    g = ch.cand_off + n_groups;
                    if (lane == 0) {
                        p.g_pos[g] = (double)cpos; p.g_len[g] = alen; p.g_mate[g] = (double)cmp; p.g_mc[g] = cmc; p.g_n[g] = 1u; act[n_act] = g;
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

}  // namespace combine
