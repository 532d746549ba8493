// cluster.cuh — stage B: leads -> bins -> clusters -> SV candidates.
//
// Data flow (all counts stay in device memory, no host round trips):
//   scatter canonical (record,k) order -> stable radix sort by (task,svtype,bin)        [A9 ordering]
//   bins -> kept bins (>= dev_min_leads_cluster non-"long" leads)                         cluster.py:248-275
//   chains of kept bins are cut at gaps no merge criterion can bridge; every piece runs
//   the reference's order-dependent merge automaton independently (one thread each)      cluster.py:278-308
//   per cluster: merge_inner, resplit / resplit_bnd                                       cluster.py:85-216
//   per sub-cluster: sv.call_from / resolve_bnd, phase aggregates                         sv.py:497-639
//   coverage probes without a per-base array                                              postprocessing.py:69-130
#pragma once
#include "common.cuh"
#include "extract.cuh"

namespace cluster {

constexpr int TYPE_SHIFT = 26;               // key = task << 29 | svtype << 26 | bin
constexpr int TASK_SHIFT = 29;
constexpr uint32_t NONE = 0xffffffffu;

struct B {
    // inputs
    snfb_lead* leads; const snfb_rec* rec; const snfb_task* task; const snfb_contig* contig; const int32_t* tr; const int32_t* tr_pmax;
    const int32_t* rec_pos; const int32_t* rec_end; const uint8_t* rec_flags; const double* rec_nm; const uint32_t* rec_nlead; const uint32_t* rec_lead_off;
    const uint32_t* task_first; const uint32_t* task_last; const int32_t* task_maxspan;
    const int32_t* mask; const uint32_t* mask_task_off;      // reference 'N' runs (may be null)
    uint32_t n_task; unsigned long long n_bound;     // upper bound on the number of leads (launch size)
    DevCounters* ctr;
    snfb_config cfg;
    // sort buffers
    uint64_t* key0; uint32_t* val0; uint64_t* key1; uint32_t* val1;
    const uint64_t* skey; const uint32_t* sval;       // sorted result
    // bins
    uint32_t* flag; uint32_t* scan;                   // generic flag / scan arrays (n_bound)
    uint32_t* bin_start; uint32_t* bin_nl; uint32_t* bin_nlong; uint32_t* bin_kept; uint32_t* bin_hap;   // hap packed 3 x 16 saturating -> two words
    uint32_t* kl_off; uint32_t* kll_off; uint32_t* kb_idx;
    // kept leads and kept bins
    uint32_t* kl; uint32_t* kll;
    uint32_t* kb_bin; uint32_t* kb_lead_off; uint32_t* kb_lead_n; uint32_t* kb_long_off; uint32_t* kb_long_n; int32_t* kb_seed; uint32_t* kb_chain; uint8_t* kb_repeat;
    // segments and clusters
    uint32_t* seg_start;
    uint32_t* c_next; uint32_t* c_last; double* c_sd; double* c_mean; uint8_t* c_rep;
    double* seg_sd_last; double* seg_maxsd_first;
    uint32_t* cl_first; uint32_t* cl_last; uint8_t* cl_rep;
    // per-cluster post-processing scratch, indexed in kept-lead space
    uint64_t* s_hi; uint64_t* s_lo; uint32_t* s_a; uint32_t* s_b; uint32_t* s_c; uint32_t* s_d; uint32_t* s_e;
    uint32_t* ord;                  // slots in merge_inner iteration order
    uint32_t* ml_slot; int32_t* ml_svlen; int32_t* ml_seqlen; uint32_t* ml_plo; uint32_t* ml_pn; uint8_t* ml_has;
    uint32_t* subl;                 // ml indices (absolute) in final sub-cluster order
    uint32_t* sub_cnt; uint32_t* sub_off;
    uint32_t* t_lo; uint32_t* t_n; int32_t* t_bin;    // per-cluster-region sub descriptors
    uint32_t* sub_cluster; uint32_t* sub_lo; uint32_t* sub_n; int32_t* sub_bin;
    // candidates
    snfb_cand* cand_tmp; uint32_t* cand_valid; uint32_t* cand_id; uint32_t* cand_nlead; uint32_t* cand_lead_off; uint32_t* cand_nrn; uint32_t* cand_rn_off;
    snfb_cand* cand; snfb_lead* cand_leads; uint32_t* cand_lead_ml; uint64_t* rnames; uint32_t* rn_off_out;
    unsigned long long cand_cap, cand_lead_cap, rn_cap;
    uint32_t* scan_tmp;
};

__device__ __forceinline__ int lf_type(uint32_t f) { return (int)(f & 7u); }

// ---------------------------------------------------------------- in-thread heap sorts
__device__ inline bool lt2(uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl) { return ah < bh || (ah == bh && al < bl); }
__device__ inline void hsort2(uint64_t* hi, uint64_t* lo, long n) {   // ascending by (hi, lo)
    if (n < 2) return;
    for (long start = n / 2 - 1; start >= 0; --start) {
        long r = start; uint64_t vh = hi[r], vl = lo[r];
        for (;;) { long c = 2 * r + 1; if (c >= n) break; if (c + 1 < n && lt2(hi[c], lo[c], hi[c + 1], lo[c + 1])) ++c; if (!lt2(vh, vl, hi[c], lo[c])) break; hi[r] = hi[c]; lo[r] = lo[c]; r = c; }
        hi[r] = vh; lo[r] = vl;
    }
    for (long end = n - 1; end > 0; --end) {
        uint64_t vh = hi[end], vl = lo[end]; hi[end] = hi[0]; lo[end] = lo[0];
        long r = 0;
        for (;;) { long c = 2 * r + 1; if (c >= end) break; if (c + 1 < end && lt2(hi[c], lo[c], hi[c + 1], lo[c + 1])) ++c; if (!lt2(vh, vl, hi[c], lo[c])) break; hi[r] = hi[c]; lo[r] = lo[c]; r = c; }
        hi[r] = vh; lo[r] = vl;
    }
}
__device__ inline void hsort1(uint64_t* a, long n) {                  // ascending, unsigned
    if (n < 2) return;
    for (long start = n / 2 - 1; start >= 0; --start) {
        long r = start; uint64_t v = a[r];
        for (;;) { long c = 2 * r + 1; if (c >= n) break; if (c + 1 < n && a[c] < a[c + 1]) ++c; if (!(v < a[c])) break; a[r] = a[c]; r = c; }
        a[r] = v;
    }
    for (long end = n - 1; end > 0; --end) {
        uint64_t v = a[end]; a[end] = a[0]; long r = 0;
        for (;;) { long c = 2 * r + 1; if (c >= end) break; if (c + 1 < end && a[c] < a[c + 1]) ++c; if (!(v < a[c])) break; a[r] = a[c]; r = c; }
        a[r] = v;
    }
}
__device__ __forceinline__ uint64_t bias64(long long v) { return (uint64_t)v ^ 0x8000000000000000ull; }
__device__ __forceinline__ long long unbias64(uint64_t v) { return (long long)(v ^ 0x8000000000000000ull); }

// ---------------------------------------------------------------- canonical order + sort keys
__global__ void k_scatter_keys(B b, unsigned long long n_slots) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_slots; i += (unsigned long long)gridDim.x * blockDim.x) {
        const snfb_lead* l = &b.leads[i];
        if (l->rec == extract::HOLE) continue;         // unused slot of a retired allocation chunk
        const uint32_t r = b.rec_lead_off[l->rec] + l->k;
        const uint64_t bin = (uint64_t)(l->ref_start / b.cfg.cluster_binsize);
        b.key0[r] = ((uint64_t)l->task << TASK_SHIFT) | ((uint64_t)lf_type(l->flags) << TYPE_SHIFT) | bin;
        b.val0[r] = (uint32_t)i;
    }
}

__global__ void k_bin_heads(B b) {
    const unsigned long long n = b.ctr->n_leads < b.n_bound ? b.ctr->n_leads : b.n_bound;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < b.n_bound; i += (unsigned long long)gridDim.x * blockDim.x)
        b.flag[i] = (i < n && (i == 0 || b.skey[i] != b.skey[i - 1])) ? 1u : 0u;
}
// flag/scan -> start index of every bin; bin_start[n_bins] = n
__global__ void k_bin_build(B b) {
    const unsigned long long n = b.ctr->n_leads < b.n_bound ? b.ctr->n_leads : b.n_bound;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
        if (b.flag[i]) b.bin_start[b.scan[i]] = (uint32_t)i;
    if (blockIdx.x == 0 && threadIdx.x == 0) b.bin_start[b.ctr->n_bins] = (uint32_t)n;
}
// per-bin statistics: hap counts, split into leads / leads_long, seq dropping (leadprov.py:400-418)
__global__ void k_bin_stats(B b) {
    const unsigned long long nb = b.ctr->n_bins;
    for (unsigned long long bi = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; bi < b.n_bound; bi += (unsigned long long)gridDim.x * blockDim.x) {
        if (bi >= nb) { b.bin_nl[bi] = 0; b.bin_nlong[bi] = 0; b.bin_kept[bi] = 0; continue; }
        const uint32_t lo = b.bin_start[bi], hi = b.bin_start[bi + 1];
        uint32_t hc[3] = { 0, 0, 0 }; uint32_t nl = 0, nlong = 0;
        for (uint32_t i = lo; i < hi; ++i) {
            snfb_lead* l = &b.leads[b.sval[i]]; uint32_t f = l->flags;
            if ((int)(i - lo) >= b.cfg.consensus_max_reads_bin && (f & SNFB_LF_HAS_SEQ)) { f &= ~SNFB_LF_HAS_SEQ; l->flags = f; l->seq_off = -1; l->seq_len = 0; }
            hc[SNFB_LF_HAP(f)]++;
            if (lf_type(f) == SNFB_INS && (f & SNFB_LF_SVLEN_NONE)) ++nlong; else ++nl;
        }
        const bool kept = (int)nl >= b.cfg.dev_min_leads_cluster;
        b.bin_kept[bi] = kept; b.bin_nl[bi] = kept ? nl : 0; b.bin_nlong[bi] = kept ? nlong : 0;
        for (int h = 0; h < 3; ++h) b.bin_hap[bi * 3 + h] = hc[h] > 65535u ? 65535u : hc[h];
    }
}
// tandem repeat flag of a seed (cluster.py:240-246): first interval whose running max end reaches the seed
__device__ inline bool within_tr(const B& b, const snfb_task& tk, int seed) {
    if (tk.tr_n <= 0 || b.tr == nullptr) return false;
    const int32_t* pm = b.tr_pmax + tk.tr_off; int lo = 0, hi = tk.tr_n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (pm[mid] >= seed) hi = mid; else lo = mid + 1; }
    int j = lo < tk.tr_n ? lo : tk.tr_n - 1;
    const int32_t* t = b.tr + 2 * (size_t)(tk.tr_off + j);
    return t[0] < seed && seed < t[1];
}
__global__ void k_kbin_build(B b) {
    const unsigned long long nb = b.ctr->n_bins;
    for (unsigned long long bi = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; bi < nb; bi += (unsigned long long)gridDim.x * blockDim.x) {
        if (!b.bin_kept[bi]) continue;
        const uint32_t kb = b.kb_idx[bi]; const uint32_t lo = b.bin_start[bi], hi = b.bin_start[bi + 1];
        uint32_t a = b.kl_off[bi], c = b.kll_off[bi];
        b.kb_bin[kb] = (uint32_t)bi; b.kb_lead_off[kb] = a; b.kb_lead_n[kb] = b.bin_nl[bi]; b.kb_long_off[kb] = c; b.kb_long_n[kb] = b.bin_nlong[bi];
        for (uint32_t i = lo; i < hi; ++i) { const uint32_t s = b.sval[i]; const uint32_t f = b.leads[s].flags;
            if (lf_type(f) == SNFB_INS && (f & SNFB_LF_SVLEN_NONE)) b.kll[c++] = s; else b.kl[a++] = s; }
        const uint64_t key = b.skey[lo]; const int seed = (int)(key & ((1ull << TYPE_SHIFT) - 1)) * b.cfg.cluster_binsize;
        const uint32_t chain = (uint32_t)(key >> TYPE_SHIFT);
        b.kb_seed[kb] = seed; b.kb_chain[kb] = chain;
        b.kb_repeat[kb] = (within_tr(b, b.task[chain >> 3], seed) || b.cfg.repeat) ? 1 : 0;
    }
}

// ---------------------------------------------------------------- chain segmentation
__device__ __forceinline__ int break_gap(const snfb_config& cfg) {
    double g = cfg.cluster_repeat_h_max > (double)cfg.cluster_merge_bnd ? cfg.cluster_repeat_h_max : (double)cfg.cluster_merge_bnd;
    return (int)g;
}
__global__ void k_seg_heads(B b) {
    const unsigned long long nk = b.ctr->n_kbins;
    const int bg = break_gap(b.cfg);
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < b.n_bound; k += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t f = 0;
        if (k < nk) f = (k == 0 || b.kb_chain[k] != b.kb_chain[k - 1] || (b.kb_seed[k] - (b.kb_seed[k - 1] + b.cfg.cluster_binsize)) > bg) ? 1u : 0u;
        b.flag[k] = f;
    }
}
__global__ void k_seg_build(B b) {
    const unsigned long long nk = b.ctr->n_kbins;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < nk; k += (unsigned long long)gridDim.x * blockDim.x)
        if (b.flag[k]) b.seg_start[b.scan[k]] = (uint32_t)k;
    if (blockIdx.x == 0 && threadIdx.x == 0) b.seg_start[b.ctr->n_segs] = (uint32_t)nk;
}

// Cluster.compute_metrics over the kept leads [lo,hi) (cluster.py:48-61)
__device__ inline void compute_metrics(const B& b, uint32_t lo, uint32_t hi, double* mean_svlen, double* sd) {
    const long len = (long)hi - lo; const long n = len < 100 ? len : 100;
    if (n == 0) { *mean_svlen = 0; *sd = 0; return; }
    if (n == 1) { *mean_svlen = (double)b.leads[b.kl[lo]].svlen; *sd = 0; return; }
    const long step = len / n;      // int(len / n)
    long long sum = 0; long m = 0; const long long base = b.leads[b.kl[lo]].ref_start; u128 sxx = 0; __int128 sx = 0;
    for (long i = 0; i < len; i += step) { const snfb_lead* l = &b.leads[b.kl[lo + i]]; sum += l->svlen; const __int128 d = (__int128)((long long)l->ref_start - base); sx += d; sxx += (u128)(d * d); ++m; }
    *mean_svlen = __ddiv_rn((double)sum, (double)n);
    *sd = sqrt_frac_rn((u128)m * sxx - (u128)(sx * sx), (uint64_t)m * (uint64_t)(m - 1));
}

// the merge automaton of cluster.resolve on one chain piece (cluster.py:278-308)
__global__ void k_merge(B b) {
    const unsigned long long ns = b.ctr->n_segs;
    const snfb_config& cfg = b.cfg;
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < ns; s += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t k0 = b.seg_start[s], k1 = b.seg_start[s + 1];
        const uint32_t chain = b.kb_chain[k0]; const int svtype = (int)(chain & 7u);
        const bool first_piece = (k0 == 0) || b.kb_chain[k0 - 1] != chain;     // global index 0 of the chain lives here
        for (uint32_t k = k0; k < k1; ++k) {
            b.c_next[k] = k + 1 < k1 ? k + 1 : NONE; b.c_last[k] = k; b.c_rep[k] = b.kb_repeat[k];
            compute_metrics(b, b.kb_lead_off[k], b.kb_lead_off[k] + b.kb_lead_n[k], &b.c_mean[k], &b.c_sd[k]);
        }
        double maxsd_first = b.c_sd[k0];
        // i walks the linked list; `prev` is tracked so that i-1 is available; idx is the position inside the piece
        uint32_t cur = k0, prev = NONE; long idx = 0;
        while (b.c_next[cur] != NONE) {
            const uint32_t nx = b.c_next[cur];
            const int cur_end = b.kb_seed[b.c_last[cur]] + cfg.cluster_binsize, nx_end = b.kb_seed[b.c_last[nx]] + cfg.cluster_binsize;
            const long long inner = (long long)b.kb_seed[nx] - cur_end, outer = (long long)nx_end - b.kb_seed[cur];
            const double msd = b.c_sd[cur] < b.c_sd[nx] ? b.c_sd[cur] : b.c_sd[nx];
            bool merge = (double)inner <= __dmul_rn(msd, cfg.cluster_r);
            if (!merge && (cfg.repeat || b.c_rep[cur] || b.c_rep[nx])) {
                double lim = __dmul_rn(__dadd_rn(fabs(b.c_mean[cur]), fabs(b.c_mean[nx])), cfg.cluster_repeat_h);
                if (cfg.cluster_repeat_h_max < lim) lim = cfg.cluster_repeat_h_max;
                merge = (double)outer <= lim;
            }
            if (!merge) merge = svtype == SNFB_BND && inner <= cfg.cluster_merge_bnd;
            if (merge) {
                b.c_next[cur] = b.c_next[nx]; b.c_last[cur] = b.c_last[nx]; b.c_rep[cur] = b.c_rep[cur] | b.c_rep[nx];
                compute_metrics(b, b.kb_lead_off[cur], b.kb_lead_off[b.c_last[cur]] + b.kb_lead_n[b.c_last[cur]], &b.c_mean[cur], &b.c_sd[cur]);
                if (cur == k0 && b.c_sd[cur] > maxsd_first) maxsd_first = b.c_sd[cur];
                // i = max(0, i-2) + 1 in chain-global indices.  Pieces after the first sit at global index >= 1,
                // where the rule reduces to "step back one cluster if there is one inside the piece".
                if (first_piece) { if (idx >= 2) { /* i-1 */ cur = prev; --idx; prev = NONE; if (idx > 0) { uint32_t p = k0; while (b.c_next[p] != cur) p = b.c_next[p]; prev = p; } }
                                   else if (idx == 0) { prev = cur; cur = b.c_next[cur]; idx = 1; if (cur == NONE) break; }
                                   /* idx == 1 stays */ }
                else { if (idx >= 1) { cur = prev; --idx; prev = NONE; if (idx > 0) { uint32_t p = k0; while (b.c_next[p] != cur) p = b.c_next[p]; prev = p; } } }
            } else { prev = cur; cur = nx; ++idx; }
        }
        // verification data for the assumed chain cuts
        uint32_t last = k0; while (b.c_next[last] != NONE) last = b.c_next[last];
        b.seg_sd_last[s] = b.c_sd[last]; b.seg_maxsd_first[s] = maxsd_first;
        // mark surviving cluster heads
        for (uint32_t k = k0; k < k1; ++k) b.flag[k] = 0;
        for (uint32_t k = k0; k != NONE; k = b.c_next[k]) b.flag[k] = 1;
    }
}
// a cut between pieces is only valid if the stdev criterion could never have bridged it
__global__ void k_verify_cuts(B b) {
    const unsigned long long ns = b.ctr->n_segs;
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1; s < ns; s += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t k0 = b.seg_start[s];
        if (b.kb_chain[k0] != b.kb_chain[k0 - 1]) continue;
        const long long gap = (long long)b.kb_seed[k0] - (b.kb_seed[k0 - 1] + b.cfg.cluster_binsize);
        const double a = b.seg_sd_last[s - 1], c = b.seg_maxsd_first[s]; const double m = a < c ? a : c;
        if ((double)gap <= __dmul_rn(m, b.cfg.cluster_r)) atomicAdd(&b.ctr->unverified_breaks, 1ULL);
    }
}
__global__ void k_cluster_build(B b) {
    const unsigned long long nk = b.ctr->n_kbins;
    for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < nk; k += (unsigned long long)gridDim.x * blockDim.x)
        if (b.flag[k]) { const uint32_t c = b.scan[k]; b.cl_first[c] = (uint32_t)k; b.cl_last[c] = b.c_last[k]; b.cl_rep[c] = b.c_rep[k]; }
}

// The per-cluster / per-candidate kernels below run one thread per item, and a heavy item (hundreds of leads in a tandem
// repeat) is one long chain of dependent gathers of 64-byte leads.  Requesting all of an item's leads up front turns the
// chain's DRAM / L2 latencies into L1 hits.
__device__ __forceinline__ void prefetch_lead(const snfb_lead* l) { asm volatile("prefetch.global.L1 [%0];" :: "l"(l)); }
constexpr long PREFETCH_MIN = 8;

// ---------------------------------------------------------------- per-cluster post-processing
// cluster.merge_inner (cluster.py:85-122), cluster.resplit (125-161), cluster.resplit_bnd (164-216)
__global__ void k_cluster_post(B b) {
    const unsigned long long ncl = b.ctr->n_clusters;
    const snfb_config& cfg = b.cfg;
    for (unsigned long long c = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; c < b.n_bound; c += (unsigned long long)gridDim.x * blockDim.x) {
        if (c >= ncl) { b.sub_cnt[c] = 0; continue; }
        const uint32_t kf = b.cl_first[c], kl_ = b.cl_last[c];
        const uint32_t lo = b.kb_lead_off[kf], hi = b.kb_lead_off[kl_] + b.kb_lead_n[kl_];
        const long n = (long)hi - lo; const int svtype = (int)(b.kb_chain[kf] & 7u);
        uint64_t* khi = b.s_hi + lo; uint64_t* klo = b.s_lo + lo;
        if (n >= PREFETCH_MIN) for (long i = 0; i < n; ++i) prefetch_lead(&b.leads[b.kl[lo + i]]);
        long nm = 0;                    // number of merged leads
        if ((svtype == SNFB_INS || svtype == SNFB_DEL)) {
            const int thr = b.cl_rep[c] ? -1 : cfg.cluster_merge_pos;
            // groups by qname in first-seen order: sort (hash, idx), run heads carry the first idx
            for (long i = 0; i < n; ++i) { khi[i] = b.leads[b.kl[lo + i]].qname_hash; klo[i] = (uint64_t)i; }
            hsort2(khi, klo, n);
            uint32_t* first = b.s_a + lo;
            for (long i = 0; i < n;) { long j = i; while (j < n && khi[j] == khi[i]) ++j; for (long q = i; q < j; ++q) first[klo[q]] = (uint32_t)klo[i]; i = j; }
            for (long i = 0; i < n; ++i) { const int rs = b.leads[b.kl[lo + i]].ref_start; khi[i] = ((uint64_t)first[i] << 32) | (uint32_t)(rs ^ 0x80000000); klo[i] = (uint64_t)i; }
            hsort2(khi, klo, n);
            for (long i = 0; i < n; ++i) b.ord[lo + i] = b.kl[lo + klo[i]];
            // fold consecutive leads of a read
            for (long i = 0; i < n;) {
                long j = i; const uint32_t g = (uint32_t)(khi[i] >> 32); while (j < n && (uint32_t)(khi[j] >> 32) == g) ++j;
                const snfb_lead* t0 = &b.leads[b.ord[lo + i]];
                long curi = i; long long sv = t0->svlen; bool hs = t0->flags & SNFB_LF_HAS_SEQ; long long sl = hs ? t0->seq_len : 0; long pn = 1; const bool crev0 = t0->flags & SNFB_LF_REVERSE; bool crev = crev0;
                int lre = t0->ref_end, lqe = t0->qry_end, lrs = t0->ref_start, lqs = t0->qry_start;
                for (long q = i + 1; q < j; ++q) {
                    const snfb_lead* to = &b.leads[b.ord[lo + q]];
                    const bool trev = to->flags & SNFB_LF_REVERSE;
                    const bool mg = thr == -1 || (((abs(to->ref_start - lre) < thr || abs(to->ref_start - lrs) < thr) && (abs(to->qry_start - lqe) < thr || abs(to->qry_start - lqs) < thr)) && crev == trev);
                    if (mg) { sv += to->svlen; if (!(to->flags & SNFB_LF_HAS_SEQ) || !hs) { hs = false; sl = 0; } else sl += to->seq_len; ++pn; }
                    else {
                        const long m = lo + nm++; b.ml_slot[m] = b.ord[lo + curi]; b.ml_svlen[m] = (int)sv; b.ml_has[m] = hs; b.ml_seqlen[m] = (int)sl; b.ml_plo[m] = (uint32_t)(lo + curi); b.ml_pn[m] = (uint32_t)pn;
                        curi = q; sv = to->svlen; hs = to->flags & SNFB_LF_HAS_SEQ; sl = hs ? to->seq_len : 0; pn = 1; crev = trev;
                    }
                    lre = to->ref_end; lqe = to->qry_end; lrs = to->ref_start; lqs = to->qry_start;
                }
                const long m = lo + nm++; b.ml_slot[m] = b.ord[lo + curi]; b.ml_svlen[m] = (int)sv; b.ml_has[m] = hs; b.ml_seqlen[m] = (int)sl; b.ml_plo[m] = (uint32_t)(lo + curi); b.ml_pn[m] = (uint32_t)pn;
                i = j;
            }
        } else {
            for (long i = 0; i < n; ++i) { const uint32_t s = b.kl[lo + i]; const snfb_lead* l = &b.leads[s]; const long m = lo + i; b.ord[m] = s;
                b.ml_slot[m] = s; b.ml_svlen[m] = l->svlen; b.ml_has[m] = (l->flags & SNFB_LF_HAS_SEQ) != 0; b.ml_seqlen[m] = l->seq_len; b.ml_plo[m] = (uint32_t)m; b.ml_pn[m] = 1; }
            nm = n;
        }
        uint32_t nsub = 0;
        if (svtype == SNFB_BND) {
            if (cfg.dev_no_resplit || nm <= 1) { for (long i = 0; i < nm; ++i) b.subl[lo + i] = (uint32_t)(lo + i); b.t_lo[lo] = lo; b.t_n[lo] = (uint32_t)nm; b.t_bin[lo] = -1; nsub = 1; }
            else {
                const int thr = cfg.cluster_merge_bnd;
                // groups by (mate_contig, is_first) in first-seen order
                for (long i = 0; i < nm; ++i) { const snfb_lead* l = &b.leads[b.ml_slot[lo + i]]; khi[i] = ((uint64_t)(uint32_t)(l->mate_contig + 2) << 1) | ((l->flags & SNFB_LF_BND_FIRST) ? 1u : 0u); klo[i] = (uint64_t)i; }
                hsort2(khi, klo, nm);
                uint32_t* first = b.s_a + lo;
                for (long i = 0; i < nm;) { long j = i; while (j < nm && khi[j] == khi[i]) ++j; for (long q = i; q < j; ++q) first[klo[q]] = (uint32_t)klo[i]; i = j; }
                for (long i = 0; i < nm; ++i) { const int mp = b.leads[b.ml_slot[lo + i]].mate_pos; const int pb = thr > 0 ? (mp / thr) * thr : 0; khi[i] = ((uint64_t)first[i] << 32) | (uint32_t)(pb ^ 0x80000000); klo[i] = (uint64_t)i; }
                hsort2(khi, klo, nm);
                long start = 0;
                for (long i = 0; i < nm; ++i) {
                    b.subl[lo + i] = (uint32_t)(lo + klo[i]);
                    const bool newgrp = i > 0 && (uint32_t)(khi[i] >> 32) != (uint32_t)(khi[i - 1] >> 32);
                    const long long pbc = (int)((uint32_t)khi[i] ^ 0x80000000), pbp = i > 0 ? (int)((uint32_t)khi[i - 1] ^ 0x80000000) : 0;
                    if (i > 0 && (newgrp || (pbc != pbp && pbc - pbp > thr))) { b.t_lo[lo + nsub] = (uint32_t)(lo + start); b.t_n[lo + nsub] = (uint32_t)(i - start); b.t_bin[lo + nsub] = -1; ++nsub; start = i; }
                }
                b.t_lo[lo + nsub] = (uint32_t)(lo + start); b.t_n[lo + nsub] = (uint32_t)(nm - start); b.t_bin[lo + nsub] = -1; ++nsub;
            }
        } else if (cfg.dev_no_resplit_repeat || cfg.dev_no_resplit) {
            for (long i = 0; i < nm; ++i) b.subl[lo + i] = (uint32_t)(lo + i); b.t_lo[lo] = lo; b.t_n[lo] = (uint32_t)nm; b.t_bin[lo] = -1; nsub = 1;
        } else {
            // resplit: distinct svlen bins ascending, then the order-dependent merge with python's negative index
            for (long i = 0; i < nm; ++i) { const int sv = b.ml_svlen[lo + i]; const int a = sv < 0 ? -sv : sv; khi[i] = (uint64_t)((a / cfg.cluster_resplit_binsize) * cfg.cluster_resplit_binsize); klo[i] = (uint64_t)i; }
            hsort2(khi, klo, nm);
            uint32_t* seg_first = b.s_a + lo; uint32_t* seg_end = b.s_b + lo; uint32_t* seg_next = b.s_c + lo; uint32_t* nc = b.s_d + lo; uint32_t* tail = b.s_e + lo;
            long nb = 0;
            for (long i = 0; i < nm;) { long j = i; while (j < nm && khi[j] == khi[i]) ++j; seg_first[nb] = (uint32_t)i; seg_end[nb] = (uint32_t)j; seg_next[nb] = NONE; tail[nb] = (uint32_t)nb; nc[nb] = (uint32_t)nb; ++nb; i = j; }
            long len = nb, i = 1;
            while (len > 1 && i < len) {
                const long li = i - 1 < 0 ? len - 1 : i - 1;
                const long long last = (long long)khi[seg_first[nc[li]]], curr = (long long)khi[seg_first[nc[i]]];
                double thr = __dmul_rn((double)(curr < last ? curr : last), cfg.cluster_merge_len); if ((double)cfg.minsvlen > thr) thr = (double)cfg.minsvlen;
                const long long d = curr - last < 0 ? last - curr : curr - last;
                if ((double)d <= thr) {
                    seg_next[tail[nc[i]]] = nc[li]; tail[nc[i]] = tail[nc[li]];          // bins[curr].extend(bins[last])
                    for (long k = li; k + 1 < len; ++k) nc[k] = nc[k + 1]; --len;        // pop(i-1)
                    i = i - 2 > 0 ? i - 2 : 0;
                } else ++i;
            }
            long w = 0;
            for (long k = 0; k < len; ++k) {
                const long start = w;
                for (uint32_t sg = nc[k]; sg != NONE; sg = seg_next[sg]) for (uint32_t q = seg_first[sg]; q < seg_end[sg]; ++q) b.subl[lo + w++] = (uint32_t)(lo + klo[q]);
                b.t_lo[lo + nsub] = (uint32_t)(lo + start); b.t_n[lo + nsub] = (uint32_t)(w - start); b.t_bin[lo + nsub] = (int)khi[seg_first[nc[k]]]; ++nsub;
            }
        }
        b.sub_cnt[c] = nsub;
    }
}
__global__ void k_sub_build(B b) {
    const unsigned long long ncl = b.ctr->n_clusters;
    for (unsigned long long c = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; c < ncl; c += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t lo = b.kb_lead_off[b.cl_first[c]]; const uint32_t o = b.sub_off[c];
        for (uint32_t k = 0; k < b.sub_cnt[c]; ++k) { b.sub_cluster[o + k] = (uint32_t)c; b.sub_lo[o + k] = b.t_lo[lo + k]; b.sub_n[o + k] = b.t_n[lo + k]; b.sub_bin[o + k] = b.t_bin[lo + k]; }
    }
}

// util.center = median_modes over a sorted (biased) array (util.py:49-58)
__device__ inline long long center_sorted(const uint64_t* a, long n) {
    long maxc = 0; for (long i = 0; i < n;) { long j = i; while (j < n && a[j] == a[i]) ++j; if (j - i > maxc) maxc = j - i; i = j; }
    long m = 0; for (long i = 0; i < n;) { long j = i; while (j < n && a[j] == a[i]) ++j; if (maxc - (j - i) < 3) ++m; i = j; }
    const long want = m / 2; long k = 0;
    for (long i = 0; i < n;) { long j = i; while (j < n && a[j] == a[i]) ++j; if (maxc - (j - i) < 3) { if (k == want) return unbias64(a[i]); ++k; } i = j; }
    return unbias64(a[0]);
}
// util.stdev(util.trim(v)) on a sorted (biased) array (util.py:25-27, 82-88)
__device__ inline double stdev_trim_sorted(const uint64_t* a, long n) {
    const long trim_n = (long)__dmul_rn(__ddiv_rn((double)n, 100.0), 25.0);
    const long lo = trim_n > 0 ? trim_n : 0, m = trim_n > 0 ? n - 2 * trim_n : n;
    return stdev_ints(m, [&](long i) { return unbias64(a[lo + i]); });
}
__device__ inline int cmp_decstr(long long a, long long b) {     // strcmp(str(a), str(b)) for the PS tie break
    char x[24], y[24]; int nx = 0, ny = 0;
    { unsigned long long v = a < 0 ? (unsigned long long)(-a) : (unsigned long long)a; char t[24]; int k = 0; do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v); if (a < 0) x[nx++] = '-'; while (k) x[nx++] = t[--k]; }
    { unsigned long long v = b < 0 ? (unsigned long long)(-b) : (unsigned long long)b; char t[24]; int k = 0; do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v); if (b < 0) y[ny++] = '-'; while (k) y[ny++] = t[--k]; }
    for (int i = 0; i < nx && i < ny; ++i) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
    return nx == ny ? 0 : (nx < ny ? -1 : 1);
}

// sv.call_from + resolve_bnd + get_sa_count + phase aggregates for one sub-cluster
__global__ void k_call(B b) {
    const unsigned long long nsub = b.ctr->n_sub;
    const snfb_config& cfg = b.cfg;
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < b.n_bound; s += (unsigned long long)gridDim.x * blockDim.x) {
        if (s >= nsub) { b.cand_valid[s] = 0; b.cand_nlead[s] = 0; b.cand_nrn[s] = 0; continue; }
        const uint32_t c = b.sub_cluster[s]; const uint32_t slo = b.sub_lo[s]; long n = b.sub_n[s];
        const uint32_t kf = b.cl_first[c], kl_ = b.cl_last[c];
        const uint32_t chain = b.kb_chain[kf]; const int svtype = (int)(chain & 7u); const int t = (int)(chain >> 3);
        const uint32_t llo = b.kb_long_off[kf], lhi = b.kb_long_off[kl_] + b.kb_long_n[kl_];
        const bool has_long = svtype == SNFB_INS; const long nlong = has_long ? (long)lhi - llo : 0;
        uint64_t* w = b.s_hi + slo;      // this sub-cluster's private scratch (n entries)
        uint64_t* w2 = b.s_lo + slo;
        b.cand_valid[s] = 0; b.cand_nlead[s] = 0; b.cand_nrn[s] = 0;
        if (n == 0) continue;
        if (n >= PREFETCH_MIN) for (long i = 0; i < n; ++i) prefetch_lead(&b.leads[b.ml_slot[b.subl[slo + i]]]);
        // svlen = center(svlens)
        for (long i = 0; i < n; ++i) w[i] = bias64(b.ml_svlen[b.subl[slo + i]]);
        hsort1(w, n);
        const long long svlen = center_sorted(w, n);
        const bool single = svtype == SNFB_SINGLE_LEFT || svtype == SNFB_SINGLE_RIGHT;
        if (!single && svtype != SNFB_BND && (svlen < 0 ? -svlen : svlen) < cfg.minsvlen_screen) continue;
        snfb_cand cd; memset(&cd, 0, sizeof cd);
        double sd_len = __longlong_as_double(0x7ff8000000000000LL);
        if (svtype != SNFB_BND) sd_len = stdev_trim_sorted(w, n);
        for (long i = 0; i < n; ++i) w[i] = bias64(b.leads[b.ml_slot[b.subl[slo + i]]].ref_start);
        hsort1(w, n);
        const long long ref_start = center_sorted(w, n);
        const double sd_pos = stdev_trim_sorted(w, n);
        const bool precise = svtype != SNFB_BND ? (__dadd_rn(sd_pos, sd_len) < (double)cfg.precise) : (sd_pos < (double)cfg.precise);
        long long svstart, svend;
        if (svtype == SNFB_INS) { svstart = svend = ref_start; }
        else if (svtype == SNFB_DEL) { svstart = ref_start + svlen; svend = ref_start; }
        else { svstart = ref_start; svend = svstart + (svlen < 0 ? -svlen : svlen); }
        long long mq = 0; long fwd = 0, sa = 0, nsplit = 0;
        for (long i = 0; i < n; ++i) { const uint32_t f = b.leads[b.ml_slot[b.subl[slo + i]]].flags; mq += SNFB_LF_MAPQ(f); fwd += !(f & SNFB_LF_REVERSE); sa += (f & SNFB_LF_IS_SA) != 0; nsplit += SNFB_LF_SOURCE(f) != SNFB_SRC_INLINE; }
        long sa_all = sa; for (long i = 0; i < nlong; ++i) sa_all += (b.leads[b.kll[llo + i]].flags & SNFB_LF_IS_SA) != 0;
        cd.sa_count = (int)sa_all; cd.sa_total = (int)(n + nlong);
        // support = distinct qnames (+ leads_long for long insertions)
        for (long i = 0; i < n; ++i) w[i] = b.leads[b.ml_slot[b.subl[slo + i]]].qname_hash;
        hsort1(w, n);
        long nq = 0; for (long i = 0; i < n; ++i) if (i == 0 || w[i] != w[i - 1]) w[nq++] = w[i];
        long support = nq, support_long = 0, extra = 0;
        const bool use_long = svtype == SNFB_INS && svlen >= cfg.long_ins_length;
        if (use_long) {
            for (long i = 0; i < nlong; ++i) { const uint64_t h = b.leads[b.kll[llo + i]].qname_hash; bool dup = false;
                for (long j = 0; j < i; ++j) if (b.leads[b.kll[llo + j]].qname_hash == h) { dup = true; break; }
                if (dup) continue; ++support_long;
                long lo2 = 0, hi2 = nq; while (lo2 < hi2) { long mid = (lo2 + hi2) >> 1; if (w[mid] < h) lo2 = mid + 1; else hi2 = mid; }
                if (!(lo2 < nq && w[lo2] == h)) ++extra; }
            support += extra;
        }
        cd.task = t; cd.svtype = svtype; cd.pos = (int)svstart; cd.end = (int)svend; cd.svlen = (int)svlen;
        cd.qual = (int)__ddiv_rn((double)mq, (double)n); cd.precise = precise; cd.fwd = (int)fwd; cd.rev = (int)(n - fwd);
        cd.stdev_pos = sd_pos; cd.stdev_len = sd_len; cd.support_long = (int)support_long; cd.bnd_mate_contig = -1; cd.nm_mean = -1.0;
        if (svtype == SNFB_DEL) cd.support_sa = (int)nsplit;
        if (cfg.qc_nm_measure) {       // util.mean(v.nm): python's sum() is Neumaier-compensated (bltinmodule.c)
            double sm = 0.0, cc = 0.0;
            for (long i = 0; i < n; ++i) { const snfb_lead* l = &b.leads[b.ml_slot[b.subl[slo + i]]]; const double x = lf_type(l->flags) == SNFB_BND ? (double)l->nm_sa : b.rec_nm[l->rec];
                const double tt = __dadd_rn(sm, x); if (fabs(sm) >= fabs(x)) cc = __dadd_rn(cc, __dadd_rn(__dadd_rn(sm, -tt), x)); else cc = __dadd_rn(cc, __dadd_rn(__dadd_rn(x, -tt), sm)); sm = tt; }
            if (cc != 0.0 && isfinite(cc)) sm = __dadd_rn(sm, cc);
            cd.nm_mean = __ddiv_rn(sm, (double)n);
        }
        long nfinal = n;
        if (svtype == SNFB_BND) {      // resolve_bnd: keep the leads of the modal mate contig
            int best = -2; long bestc = 0; int bestrank = 0;
            for (long i = 0; i < n; ++i) { const int mc = b.leads[b.ml_slot[b.subl[slo + i]]].mate_contig; long k = 0; for (long j = 0; j < n; ++j) k += b.leads[b.ml_slot[b.subl[slo + j]]].mate_contig == mc;
                const int rk = mc >= 0 ? b.contig[mc].lex_rank : 1 << 30; if (k > bestc || (k == bestc && rk < bestrank)) { best = mc; bestc = k; bestrank = rk; } }
            long m = 0, nf = 0, nr = 0;
            for (long i = 0; i < n; ++i) { const uint32_t mi = b.subl[slo + i]; const snfb_lead* l = &b.leads[b.ml_slot[mi]]; if (l->mate_contig != best) continue;
                b.subl[slo + m] = mi; w[m] = bias64(l->mate_pos); w2[m] = l->qname_hash; nf += (l->flags & SNFB_LF_BND_FIRST) != 0; nr += (l->flags & SNFB_LF_BND_REVERSE) != 0; ++m; }
            hsort1(w, m); cd.bnd_mate_contig = best; cd.bnd_mate_pos = (int)center_sorted(w, m);
            cd.bnd_is_first = nf > m - nf; cd.bnd_is_reverse = nr > m - nr;
            hsort1(w2, m); long q = 0; for (long i = 0; i < m; ++i) if (i == 0 || w2[i] != w2[i - 1]) ++q; support = q; nfinal = m;
        }
        cd.support = (int)support;
        { const uint32_t bi = b.kb_bin[kf]; for (int h = 0; h < 3; ++h) cd.hap_counts[h] = (int)b.bin_hap[(size_t)bi * 3 + h]; }   // REF part is filled by k_coverage
        cd.cluster_seed = b.kb_seed[kf]; cd.resplit_bin = b.sub_bin[s];
        cd.lead_n = (int)nfinal; cd.long_n = (has_long && svtype != SNFB_BND) ? (int)nlong : 0;
        b.cand_tmp[s] = cd; b.cand_valid[s] = 1; b.cand_nlead[s] = (uint32_t)(nfinal + cd.long_n); b.cand_nrn[s] = (uint32_t)support;
    }
}

// coverage restricted by haplotype at one position: reads with start <= p < end (leadprov.py:510).
// Warp-cooperative: records are coordinate sorted, so the covering reads start inside
// (p - longest reference span, p]; the lanes split that window.
__device__ inline void cover_count_warp(const B& b, int t, long long p, uint32_t out[3]) {
    out[0] = out[1] = out[2] = 0;
    const uint32_t lo = b.task_first[t], hi = b.task_last[t];
    if (lo >= hi) return;
    uint32_t a = lo, z = hi;                     // first record with pos > p
    while (a < z) { uint32_t mid = a + ((z - a) >> 1); if ((long long)b.rec_pos[mid] <= p) a = mid + 1; else z = mid; }
    const long long span = b.task_maxspan[t];
    uint32_t c0 = 0, c1 = 0, c2 = 0;
    for (long long i = (long long)a - 1 - lane_id(); i >= (long long)lo; i -= 32) {
        const long long ps = b.rec_pos[i]; if (ps + span <= p) break;
        const uint8_t f = b.rec_flags[i];
        if ((f & extract::RF_PASS) && (long long)b.rec_end[i] > p) { const int h = (f >> 2) & 3; c0 += h == 0; c1 += h == 1; c2 += h == 2; }
    }
    out[0] = __reduce_add_sync(FULL, c0); out[1] = __reduce_add_sync(FULL, c1); out[2] = __reduce_add_sync(FULL, c2);
}
__device__ inline void cov_at_warp(const B& b, int t, long long idx, int* out) {     // numpy indexing of the uint16 coverage vector
    const long long L = b.task[t].contig_len;
    if (idx < 0) idx += L;
    if (idx < 0 || idx >= L) return;             // IndexError: the field keeps its default 0
    if (b.mask) {                                 // _mask_N_coverage: positions inside a reference 'N' run read 0 (leadprov.py:439)
        uint32_t lo = b.mask_task_off[t], hi = b.mask_task_off[t + 1];
        while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if ((long long)b.mask[2 * mid + 1] <= idx) lo = mid + 1; else hi = mid; }
        if (lo < b.mask_task_off[t + 1] && (long long)b.mask[2 * lo] <= idx && idx >= b.task[t].start && idx < b.task[t].end) { *out = 0; return; }   // the mask is fetched per region (leadprov.py:436-438)
    }
    uint32_t c[3]; cover_count_warp(b, t, idx, c); *out = (int)((c[0] + c[1] + c[2]) & 0xffffu);
}

// final candidate records in reference order + their leads, read names, phase aggregates, coverage
__global__ void k_cand_finish(B b) {
    const unsigned long long nsub = b.ctr->n_sub;
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < nsub; s += (unsigned long long)gridDim.x * blockDim.x) {
        if (!b.cand_valid[s]) continue;
        const uint32_t id = b.cand_id[s]; if (id >= b.cand_cap) { atomicAdd(&b.ctr->scratch_overflow, 1ULL); continue; }
        snfb_cand cd = b.cand_tmp[s];
        const uint32_t c = b.sub_cluster[s]; const uint32_t slo = b.sub_lo[s]; const long n = cd.lead_n;
        const uint32_t kf = b.cl_first[c], kl_ = b.cl_last[c]; const uint32_t llo = b.kb_long_off[kf];
        const uint32_t lo_out = b.cand_lead_off[s]; const uint32_t rn_out = b.cand_rn_off[s];
        (void)kl_;
        cd.lead_off = (int)lo_out; cd.long_off = (int)(lo_out + n); cd.alt_off = -1; cd.alt_len = 0;
        if ((unsigned long long)lo_out + n + cd.long_n > b.cand_lead_cap || (unsigned long long)rn_out + cd.support > b.rn_cap) { atomicAdd(&b.ctr->scratch_overflow, 1ULL); continue; }
        uint64_t* w = b.s_hi + slo; uint64_t* w2 = b.s_lo + slo;
        if (n >= PREFETCH_MIN) for (long i = 0; i < n; ++i) prefetch_lead(&b.leads[b.ml_slot[b.subl[slo + i]]]);
        int nf = 0, nr = 0; long ninl = 0;
        for (long i = 0; i < n; ++i) {
            const uint32_t mi = b.subl[slo + i]; snfb_lead L = b.leads[b.ml_slot[mi]];
            L.svlen = b.ml_svlen[mi];
            if (b.ml_has[mi]) { L.flags |= SNFB_LF_HAS_SEQ; L.seq_len = b.ml_seqlen[mi]; } else { L.flags &= ~SNFB_LF_HAS_SEQ; L.seq_len = 0; L.seq_off = -1; }
            b.cand_leads[lo_out + i] = L;
            if (L.flags & SNFB_LF_REVERSE) nr = 1; else nf = 1;
            w[i] = L.qname_hash;
            if (SNFB_LF_SOURCE(L.flags) == SNFB_SRC_INLINE) w2[ninl++] = L.qname_hash;
        }
        for (long i = 0; i < cd.long_n; ++i) b.cand_leads[lo_out + n + i] = b.leads[b.kll[llo + i]];
        cd.n_strands = nf + nr;
        hsort1(w2, ninl); { long q = 0; for (long i = 0; i < ninl; ++i) if (i == 0 || w2[i] != w2[i - 1]) ++q; cd.support_inline = (int)q; }
        // read names: distinct hashes, ascending
        hsort1(w, n); long nq = 0; for (long i = 0; i < n; ++i) if (i == 0 || w[i] != w[i - 1]) w[nq++] = w[i];
        long o = 0;
        for (long i = 0; i < nq && o < cd.support; ++i) b.rnames[rn_out + o++] = w[i];
        if (cd.svtype == SNFB_INS && cd.svlen >= b.cfg.long_ins_length) {
            // union with leads_long: insert the extra hashes, then keep the list sorted
            for (long i = 0; i < cd.long_n && o < cd.support; ++i) { const uint64_t h = b.leads[b.kll[llo + i]].qname_hash; bool dup = false; for (long j = 0; j < o; ++j) if (b.rnames[rn_out + j] == h) { dup = true; break; } if (!dup) b.rnames[rn_out + o++] = h; }
            hsort1(b.rnames + rn_out, o);
        }
        b.rn_off_out[id] = rn_out;
        // phase aggregates: reads_phases = {read_id: (hap, phase_set)}, the last lead of a record wins
        {
            long hc[3] = { 0, 0, 0 };
            for (long i = 0; i < n; ++i) { const snfb_lead* l = &b.cand_leads[lo_out + i]; w[i] = ((uint64_t)l->rec << 32) | (uint32_t)i; }
            hsort1(w, n);
            long np = 0;       // (ps value, isnull) of the distinct records -> w2 as packed keys
            for (long i = 0; i < n; ++i) {
                if (i + 1 < n && (w[i + 1] >> 32) == (w[i] >> 32)) continue;        // not the last lead of this record
                const snfb_lead* l = &b.cand_leads[lo_out + (uint32_t)w[i]];
                const bool bnd = lf_type(l->flags) == SNFB_BND;
                hc[bnd ? 0 : SNFB_LF_HAP(l->flags)]++;
                const snfb_rec* r = &b.rec[l->rec]; const bool isnull = bnd || !(r->aux_flags & SNFB_AUX_PS);
                w2[np++] = isnull ? 0xffffffffffffffffull : bias64(r->ps);
            }
            int ht = 0; for (int h = 1; h < 3; ++h) if (hc[h] > 0 && hc[h] >= hc[ht]) ht = h;
            cd.hp_top = ht; cd.hp_support = (int)hc[ht]; cd.hp_other = (int)(hc[0] + hc[1] + hc[2] - hc[ht]);
            hsort1(w2, np);
            long bc = 0; uint64_t bv = 0; bool have = false; long nonnull = 0;
            for (long i = 0; i < np;) { long j = i; while (j < np && w2[j] == w2[i]) ++j; const long cnt = j - i; const bool isnull = w2[i] == 0xffffffffffffffffull; if (!isnull) nonnull += cnt;
                bool gt;
                if (!have) gt = true; else if (cnt != bc) gt = cnt > bc; else { const bool bn = bv == 0xffffffffffffffffull; if (isnull != bn) gt = isnull; else gt = cmp_decstr(unbias64(w2[i]), unbias64(bv)) > 0; }
                if (gt) { bc = cnt; bv = w2[i]; have = true; } i = j; }
            cd.ps_top_null = bv == 0xffffffffffffffffull; cd.ps_top = cd.ps_top_null ? 0 : (int)unbias64(bv); cd.ps_support = (int)bc; cd.ps_other = (int)(nonnull - (cd.ps_top_null ? 0 : bc));
        }
        b.cand[id] = cd;
    }
}
// postprocessing.coverage (postprocessing.py:69-130) including the `end` that leaks from the previous call,
// plus the hap-REF counts of the cluster's first bin (cluster.py:255-260).  One warp per candidate.
__global__ void k_coverage(B b) {
    const unsigned long long nc = b.ctr->n_cand < b.cand_cap ? b.ctr->n_cand : b.cand_cap;
    const long long bs = b.cfg.coverage_binsize, ud = (long long)b.cfg.coverage_binsize * b.cfg.coverage_updown_bins;
    const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nc; i += nw) {
        snfb_cand* c = &b.cand[i]; long long start = c->pos, end; const int t = c->task; const int svtype = c->svtype;
        if (svtype == SNFB_INS) end = start + 1;
        else if (svtype == SNFB_BND) {
            if (c->bnd_is_first) start -= 1;
            long long j = (long long)i - 1; while (j >= 0 && b.cand[j].task == t && b.cand[j].svtype == SNFB_BND) --j;
            if (j >= 0 && b.cand[j].task == t) { const snfb_cand* p = &b.cand[j]; end = p->svtype == SNFB_INS ? (long long)p->pos + 1 : (long long)p->pos + (p->svlen < 0 ? -(long long)p->svlen : p->svlen); }
            else { end = start; if (lane_id() == 0) atomicAdd(&b.ctr->soft_errors, 1ULL); }
        } else end = (long long)c->pos + (c->svlen < 0 ? -(long long)c->svlen : c->svlen);
        int v[5] = { 0, 0, 0, 0, 0 };          // upstream, start, center, end, downstream
        if (svtype == SNFB_INS || svtype == SNFB_BND) { cov_at_warp(b, t, start - bs, &v[1]); cov_at_warp(b, t, start, &v[2]); cov_at_warp(b, t, end + bs, &v[3]); }
        else { cov_at_warp(b, t, start, &v[1]); cov_at_warp(b, t, (long long)__ddiv_rn((double)(start + end), 2.0), &v[2]); cov_at_warp(b, t, end - bs, &v[3]); }
        cov_at_warp(b, t, start - ud, &v[0]); cov_at_warp(b, t, end + ud, &v[4]);
        uint32_t hr[3]; const int cb = b.cfg.cluster_binsize; cover_count_warp(b, t, (long long)(c->cluster_seed / cb) * cb + cb - 1, hr);
        if (lane_id() == 0) {
            c->cov_upstream = v[0]; c->cov_start = v[1]; c->cov_center = v[2]; c->cov_end = v[3]; c->cov_downstream = v[4];
            for (int h = 0; h < 3; ++h) c->hap_counts[3 + h] = (int)(hr[h] > 65535u ? 65535u : hr[h]);
        }
    }
}

// _mask_N_coverage for the contig mean: subtract the read bases that fall inside reference 'N' runs.  One warp per run.
__global__ void k_mask_bp(B b, const uint32_t* __restrict__ mask_task, uint32_t n_mask, unsigned long long* task_cov_bp) {
    const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long m = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < n_mask; m += nw) {
        const int t = (int)mask_task[m]; const long long L = b.task[t].contig_len;
        long long a = b.mask[2 * m], e = b.mask[2 * m + 1]; if (a < b.task[t].start) a = b.task[t].start; if (e > b.task[t].end) e = b.task[t].end; if (a < 0) a = 0; if (e > L) e = L;
        const uint32_t lo = b.task_first[t], hi = b.task_last[t];
        unsigned long long sum = 0;
        if (a < e && lo < hi) {
            uint32_t x = lo, z = hi;                 // first record with pos >= e
            while (x < z) { const uint32_t mid = x + ((z - x) >> 1); if ((long long)b.rec_pos[mid] < e) x = mid + 1; else z = mid; }
            const long long span = b.task_maxspan[t];
            for (long long i = (long long)x - 1 - lane_id(); i >= (long long)lo; i -= 32) {
                const long long ps = b.rec_pos[i]; if (ps + span <= a) break;
                if (!(b.rec_flags[i] & extract::RF_PASS)) continue;
                long long re = b.rec_end[i]; if (re > L) re = L;
                const long long o0 = ps > a ? ps : a, o1 = re < e ? re : e;
                if (o1 > o0) sum += (unsigned long long)(o1 - o0);
            }
        }
        #pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
        if (lane_id() == 0 && sum) atomicAdd(&task_cov_bp[t], 0ull - sum);
    }
}

}  // namespace cluster
