// api.cu — C ABI of libsnfb200.so (include/snfb.h): context, memory, stage drivers.
// The product path has no CPU implementation: every stage below launches CUDA kernels.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <omp.h>

#include "common.cuh"
#include "prims.cuh"
#include "extract.cuh"
#include "cluster.cuh"
#include "consensus.cuh"

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; return 1; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
struct HostBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; return 1; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

constexpr int MAX_TIMINGS = 64;

struct snfb_ctx {
    int device = 0; cudaStream_t st = nullptr, st_copy = nullptr; cudaEvent_t ev_copy = nullptr; bool want_cand_prefetch = false, cand_prefetched = false; std::string err;
    snfb_config cfg{}; bool have_cfg = false;
    // records
    bool loaded = false, on_device = false, seq_on_demand = false; const uint8_t* h_seq = nullptr;
    uint64_t n_rec = 0, n_cigar = 0, n_var = 0, n_seq = 0; uint32_t n_task = 0, n_contig = 0, n_tr = 0;
    const snfb_rec* d_rec = nullptr; const uint16_t* d_cigar = nullptr; const uint8_t* d_var = nullptr; const uint8_t* d_seq = nullptr;
    DevBuf b_rec, b_cigar, b_var, b_seq, b_task, b_contig, b_tr, b_trp, b_mask, b_mask_off, b_mask_task; uint32_t n_mask = 0;
    // stage A outputs
    DevBuf b_ctr, b_leads, b_rec_pos, b_rec_end, b_rec_flags, b_rec_nm, b_rec_nlead, b_rec_lead_off, b_task_first, b_task_last, b_task_reads, b_task_cov, b_task_span, b_task_nm, b_nm_part, b_nm_cnt, b_ev, b_sa_list, b_scanrec, b_clip, b_rec_big, b_sa_seg;
    HostBuf h_c16, h_rec16;        // BAM32 host input converted to CIGAR16 before the upload
    unsigned long long lead_cap = 0;
    DevCounters h_ctr{};
    // stage B
    DevBuf b_key0, b_val0, b_key1, b_val1, b_flag, b_scan, b_hist, b_scan_tmp;
    DevBuf b_bin_start, b_bin_nl, b_bin_nlong, b_bin_kept, b_bin_hap, b_kl_off, b_kll_off, b_kb_idx, b_kl, b_kll;
    DevBuf b_kb_bin, b_kb_lead_off, b_kb_lead_n, b_kb_long_off, b_kb_long_n, b_kb_seed, b_kb_chain, b_kb_repeat;
    DevBuf b_seg_start, b_c_next, b_c_last, b_c_sd, b_c_mean, b_c_rep, b_seg_sd_last, b_seg_maxsd, b_cl_first, b_cl_last, b_cl_rep;
    DevBuf b_s_hi, b_s_lo, b_s_a, b_s_b, b_s_c, b_s_d, b_s_e, b_ord, b_ml_slot, b_ml_svlen, b_ml_seqlen, b_ml_plo, b_ml_pn, b_ml_has, b_subl;
    DevBuf b_sub_cnt, b_sub_off, b_t_lo, b_t_n, b_t_bin, b_sub_cluster, b_sub_lo, b_sub_n, b_sub_bin;
    DevBuf b_cand_tmp, b_cand_valid, b_cand_id, b_cand_nlead, b_cand_lead_off, b_cand_nrn, b_cand_rn_off, b_cand, b_cand_leads, b_cand_lead_ml, b_rnames, b_rn_off_out;
    unsigned long long n_bound = 0, cand_cap = 0, cand_lead_cap = 0, rn_cap = 0;
    bool sorted_in_first = true, stage_a_done = false, stage_b_done = false;
    // stage C
    DevBuf b_plan_best, b_plan_nother, b_plan_otot, b_alt_len, b_scr_len, b_alt_off, b_scr_off, b_alt, b_scr, b_sorted_leads, b_work_big, b_work_small, b_work_ctr, b_seq_req, b_arena_off, b_seq_arena, b_items_big, b_items_small, b_tiles;
    HostBuf h_seq_req, h_seq_arena; uint64_t seq_h2d_bytes = 0;
    // host staging
    HostBuf h_leads, h_task_reads, h_task_nm, h_rec_nm, h_cand, h_cand_leads, h_rnames, h_rn_off, h_task_cov, h_alt;
    std::vector<double> task_cov_mean; std::vector<snfb_task> tasks;
    // timings
    cudaEvent_t ev[MAX_TIMINGS + 1]; const char* ev_name[MAX_TIMINGS + 1]; uint64_t ev_bytes[MAX_TIMINGS + 1]; int n_ev = 0; int n_ev_load = 0; uint64_t launches = 0;
};

static void ctx_fail(snfb_ctx* ctx, const char* what, const char* msg) { ctx->err = std::string(what) + ": " + msg; }
static int fail(snfb_ctx* ctx, const std::string& m) { ctx->err = m; return 1; }

// timing marks: every call records an event on the ctx stream; a named mark opens an interval that the
// next mark (named or not) closes, so host-side gaps between stages are never attributed to a kernel
static void mark(snfb_ctx* ctx, const char* name, uint64_t bytes = 0) {
    if (ctx->n_ev >= MAX_TIMINGS) return;
    cudaEventRecord(ctx->ev[ctx->n_ev], ctx->st);
    ctx->ev_name[ctx->n_ev] = name; ctx->ev_bytes[ctx->n_ev] = bytes; ++ctx->n_ev;
}

extern "C" {

int snfb_version(void) { return SNFB_ABI_VERSION; }
size_t snfb_sizeof(int which) {
    switch (which) { case 0: return sizeof(snfb_rec); case 1: return sizeof(snfb_task); case 2: return sizeof(snfb_contig); case 3: return sizeof(snfb_records);
                     case 4: return sizeof(snfb_config); case 5: return sizeof(snfb_lead); case 6: return sizeof(snfb_cand); default: return 0; }
}

uint64_t snfb_hash_name(const char* s, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull; for (size_t i = 0; i < n; ++i) { h ^= (uint8_t)s[i]; h *= 0x100000001b3ull; } return h;
}

int snfb_ctx_create(int device, snfb_ctx** out) {
    if (!out) return 1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return 2;   // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return 3;
    snfb_ctx* ctx = new snfb_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return 4; }
    cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming);
    for (int i = 0; i <= MAX_TIMINGS; ++i) cudaEventCreate(&ctx->ev[i]);
    *out = ctx; return 0;
}

void snfb_ctx_destroy(snfb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    DevBuf* bufs[] = { &ctx->b_rec, &ctx->b_cigar, &ctx->b_var, &ctx->b_seq, &ctx->b_task, &ctx->b_contig, &ctx->b_tr, &ctx->b_trp, &ctx->b_mask, &ctx->b_mask_off, &ctx->b_mask_task, &ctx->b_ctr, &ctx->b_leads, &ctx->b_rec_pos, &ctx->b_rec_end,
        &ctx->b_rec_flags, &ctx->b_rec_nm, &ctx->b_rec_nlead, &ctx->b_rec_lead_off, &ctx->b_task_first, &ctx->b_task_last, &ctx->b_task_reads, &ctx->b_task_cov, &ctx->b_task_span, &ctx->b_task_nm, &ctx->b_nm_part, &ctx->b_nm_cnt, &ctx->b_ev, &ctx->b_sa_list, &ctx->b_scanrec, &ctx->b_clip, &ctx->b_rec_big, &ctx->b_sa_seg,
        &ctx->b_key0, &ctx->b_val0, &ctx->b_key1, &ctx->b_val1, &ctx->b_flag, &ctx->b_scan, &ctx->b_hist, &ctx->b_scan_tmp, &ctx->b_bin_start, &ctx->b_bin_nl, &ctx->b_bin_nlong, &ctx->b_bin_kept,
        &ctx->b_bin_hap, &ctx->b_kl_off, &ctx->b_kll_off, &ctx->b_kb_idx, &ctx->b_kl, &ctx->b_kll, &ctx->b_kb_bin, &ctx->b_kb_lead_off, &ctx->b_kb_lead_n, &ctx->b_kb_long_off, &ctx->b_kb_long_n,
        &ctx->b_kb_seed, &ctx->b_kb_chain, &ctx->b_kb_repeat, &ctx->b_seg_start, &ctx->b_c_next, &ctx->b_c_last, &ctx->b_c_sd, &ctx->b_c_mean, &ctx->b_c_rep, &ctx->b_seg_sd_last, &ctx->b_seg_maxsd,
        &ctx->b_cl_first, &ctx->b_cl_last, &ctx->b_cl_rep, &ctx->b_s_hi, &ctx->b_s_lo, &ctx->b_s_a, &ctx->b_s_b, &ctx->b_s_c, &ctx->b_s_d, &ctx->b_s_e, &ctx->b_ord, &ctx->b_ml_slot, &ctx->b_ml_svlen,
        &ctx->b_ml_seqlen, &ctx->b_ml_plo, &ctx->b_ml_pn, &ctx->b_ml_has, &ctx->b_subl, &ctx->b_sub_cnt, &ctx->b_sub_off, &ctx->b_t_lo, &ctx->b_t_n, &ctx->b_t_bin, &ctx->b_sub_cluster, &ctx->b_sub_lo,
        &ctx->b_sub_n, &ctx->b_sub_bin, &ctx->b_cand_tmp, &ctx->b_cand_valid, &ctx->b_cand_id, &ctx->b_cand_nlead, &ctx->b_cand_lead_off, &ctx->b_cand_nrn, &ctx->b_cand_rn_off, &ctx->b_cand,
        &ctx->b_cand_leads, &ctx->b_cand_lead_ml, &ctx->b_rnames, &ctx->b_rn_off_out, &ctx->b_plan_best, &ctx->b_plan_nother, &ctx->b_plan_otot, &ctx->b_alt_len, &ctx->b_scr_len, &ctx->b_alt_off, &ctx->b_scr_off,
        &ctx->b_alt, &ctx->b_scr, &ctx->b_sorted_leads, &ctx->b_work_big, &ctx->b_work_small, &ctx->b_work_ctr, &ctx->b_seq_req, &ctx->b_arena_off, &ctx->b_seq_arena, &ctx->b_items_big, &ctx->b_items_small, &ctx->b_tiles };
    for (DevBuf* b : bufs) b->release();
    HostBuf* hb[] = { &ctx->h_leads, &ctx->h_task_reads, &ctx->h_task_nm, &ctx->h_rec_nm, &ctx->h_cand, &ctx->h_cand_leads, &ctx->h_rnames, &ctx->h_rn_off, &ctx->h_task_cov, &ctx->h_alt, &ctx->h_seq_req, &ctx->h_seq_arena, &ctx->h_c16, &ctx->h_rec16 };
    for (HostBuf* b : hb) b->release();
    for (int i = 0; i <= MAX_TIMINGS; ++i) cudaEventDestroy(ctx->ev[i]);
    cudaEventDestroy(ctx->ev_copy); cudaStreamDestroy(ctx->st_copy); cudaStreamDestroy(ctx->st);
    delete ctx;
}

const char* snfb_last_error(snfb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int snfb_set_config(snfb_ctx* ctx, const snfb_config* cfg) {
    if (!ctx || !cfg) return 1;
    if (cfg->consensus_kmer_len != 6) return fail(ctx, "consensus_kmer_len must be 6 (the reference fixes it, config.py:550)");
    if (cfg->cluster_binsize <= 0 || cfg->cluster_resplit_binsize <= 0 || cfg->coverage_binsize <= 0) return fail(ctx, "bin sizes must be positive");
    ctx->cfg = *cfg; ctx->have_cfg = true; return 0;
}

// ---- BAM CIGAR words -> CIGAR16 (include/snfb.h).  Host code; the only place the 32-bit form is read. ----
static inline int c16_group_words(uint32_t len) { return len < (1u << 12) ? 1 : (len < (1u << 24) ? 2 : 3); }
static const uint8_t C16_CLASS[9] = { 3, 1, 2, 6, 5, 4, 0, 3, 3 };     // M I D N S H P = X
// number of 16-bit words of one record, pad words included (a group never straddles an 8-word boundary); 0 = bad op
static inline uint64_t c16_count(const uint32_t* cg, uint32_t n, bool* bad) {
    uint64_t k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if ((cg[i] & 15u) > 8u) { *bad = true; return 0; }
        const int g = c16_group_words(cg[i] >> 4);
        if ((k & 7) + g > 8) k = (k + 7) & ~7ull;
        k += g;
    }
    return k;
}
static inline void c16_write(const uint32_t* cg, uint32_t n, uint16_t* out) {
    uint64_t k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t len = cg[i] >> 4; const int g = c16_group_words(len);
        if ((k & 7) + g > 8) { while (k & 7) out[k++] = 0; }
        out[k++] = (uint16_t)((C16_CLASS[cg[i] & 15u] << 12) | (len & 0xfffu));
        if (g >= 2) out[k++] = (uint16_t)(0x8000u | (1u << 12) | ((len >> 12) & 0xfffu));
        if (g >= 3) out[k++] = (uint16_t)(0x8000u | (2u << 12) | ((len >> 24) & 0xfffu));
    }
}
uint64_t snfb_pack_cigar16(const snfb_rec* rec_in, uint64_t n_rec, const uint32_t* cigar32, snfb_rec* rec_out, uint16_t* out16, uint64_t out_cap) {
    if (n_rec && (!rec_in || !cigar32)) return UINT64_MAX;
    std::vector<uint64_t> off(n_rec + 1, 0);
    bool bad = false;
    #pragma omp parallel for schedule(static) reduction(|| : bad)
    for (long long i = 0; i < (long long)n_rec; ++i) { bool b = false; const uint64_t w = c16_count(cigar32 + rec_in[i].cigar_off, rec_in[i].n_cigar, &b); bad = bad || b; off[i + 1] = (w + 7) & ~7ull; }
    if (bad) return UINT64_MAX;
    for (uint64_t i = 0; i < n_rec; ++i) off[i + 1] += off[i];
    const uint64_t total = off[n_rec] + 8;                       // one zero group of slack after the last record
    if (!out16) return total;
    if (!rec_out || out_cap < total) return UINT64_MAX;
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n_rec; ++i) {
        uint16_t* dst = out16 + off[i]; const uint64_t span = off[i + 1] - off[i];
        memset(dst, 0, 2 * span);
        c16_write(cigar32 + rec_in[i].cigar_off, rec_in[i].n_cigar, dst);
        bool b = false;
        snfb_rec r = rec_in[i]; r.n_cigar = (uint32_t)c16_count(cigar32 + rec_in[i].cigar_off, rec_in[i].n_cigar, &b); r.cigar_off = off[i];
        rec_out[i] = r;
    }
    memset(out16 + off[n_rec], 0, 16);
    return total;
}

int snfb_load_records(snfb_ctx* ctx, const snfb_records* R) {
    if (!ctx || !R) return 1;
    cudaSetDevice(ctx->device);
    ctx->loaded = false; ctx->stage_a_done = ctx->stage_b_done = false;
    if (R->n_task == 0 || R->n_task > 65535) return fail(ctx, "n_task must be in 1..65535");
    if (R->n_rec > 0xfffffff0ull) return fail(ctx, "too many records in one block");
    ctx->n_rec = R->n_rec; ctx->n_cigar = R->n_cigar; ctx->n_var = R->n_var; ctx->n_seq = R->n_seq;
    ctx->n_task = R->n_task; ctx->n_contig = R->n_contig; ctx->n_tr = R->n_tr; ctx->on_device = R->on_device == SNFB_MEM_DEVICE; ctx->seq_on_demand = R->on_device == SNFB_MEM_HOST_SEQ_ON_DEMAND; ctx->h_seq = ctx->seq_on_demand ? R->seq : nullptr;
    ctx->n_ev = 0;
    const snfb_rec* src_rec = R->rec; const uint16_t* src_cigar = reinterpret_cast<const uint16_t*>(R->cigar); uint64_t n_words = R->n_cigar;
    if (R->cigar_fmt == SNFB_CIGAR_BAM32) {
        if (R->on_device == SNFB_MEM_DEVICE) return fail(ctx, "device-resident records must carry CIGAR16 (convert with snfb_pack_cigar16)");
        // host conversion: the kernels only read CIGAR16
        const uint64_t need = snfb_pack_cigar16(R->rec, R->n_rec, reinterpret_cast<const uint32_t*>(R->cigar), nullptr, nullptr, 0);
        if (need == UINT64_MAX) return fail(ctx, "a CIGAR holds an operation the path does not know");
        if (ctx->h_c16.ensure(2 * need + 16) || ctx->h_rec16.ensure(sizeof(snfb_rec) * (R->n_rec + 1))) return fail(ctx, "out of pinned memory for the CIGAR16 conversion");
        if (snfb_pack_cigar16(R->rec, R->n_rec, reinterpret_cast<const uint32_t*>(R->cigar), ctx->h_rec16.as<snfb_rec>(), ctx->h_c16.as<uint16_t>(), need) != need) return fail(ctx, "CIGAR16 conversion failed");
        src_rec = ctx->h_rec16.as<snfb_rec>(); src_cigar = ctx->h_c16.as<uint16_t>(); n_words = need;
    } else if (R->cigar_fmt != SNFB_CIGAR_16) return fail(ctx, "unknown cigar_fmt");
    if (n_words & 7) return fail(ctx, "a CIGAR16 arena must be padded to a multiple of 8 words");
    ctx->n_cigar = n_words;
    mark(ctx, "h2d_records", sizeof(snfb_rec) * R->n_rec + 2 * n_words + R->n_var + (R->on_device == SNFB_MEM_HOST_SEQ_ON_DEMAND ? 0 : R->n_seq));
    if (ctx->on_device) {
        ctx->d_rec = R->rec; ctx->d_cigar = src_cigar; ctx->d_var = R->var; ctx->d_seq = R->seq;     // caller keeps them alive
    } else {
        if (ctx->b_rec.ensure(sizeof(snfb_rec) * (R->n_rec + 1)) || ctx->b_cigar.ensure(2 * (n_words + 16)) || ctx->b_var.ensure(R->n_var + 16) || (!ctx->seq_on_demand && ctx->b_seq.ensure(R->n_seq + 16)))
            return fail(ctx, "out of device memory for the record block");
        CUDA_TRY(cudaMemcpyAsync(ctx->b_rec.p, src_rec, sizeof(snfb_rec) * R->n_rec, cudaMemcpyHostToDevice, ctx->st));
        CUDA_TRY(cudaMemcpyAsync(ctx->b_cigar.p, src_cigar, 2 * n_words, cudaMemcpyHostToDevice, ctx->st));
        CUDA_TRY(cudaMemcpyAsync(ctx->b_var.p, R->var, R->n_var, cudaMemcpyHostToDevice, ctx->st));
        if (!ctx->seq_on_demand) CUDA_TRY(cudaMemcpyAsync(ctx->b_seq.p, R->seq, R->n_seq, cudaMemcpyHostToDevice, ctx->st));
        ctx->d_rec = ctx->b_rec.as<snfb_rec>(); ctx->d_cigar = ctx->b_cigar.as<uint16_t>(); ctx->d_var = ctx->b_var.as<uint8_t>(); ctx->d_seq = ctx->b_seq.as<uint8_t>();
    }
    ctx->tasks.assign(R->task, R->task + R->n_task);
    if (ctx->b_task.ensure(sizeof(snfb_task) * R->n_task) || ctx->b_contig.ensure(sizeof(snfb_contig) * (R->n_contig + 1)) || ctx->b_tr.ensure(8 * ((size_t)R->n_tr + 1)) || ctx->b_trp.ensure(4 * ((size_t)R->n_tr + 1)))
        return fail(ctx, "out of device memory for the task tables");
    CUDA_TRY(cudaMemcpyAsync(ctx->b_task.p, R->task, sizeof(snfb_task) * R->n_task, cudaMemcpyHostToDevice, ctx->st));
    if (R->n_contig) CUDA_TRY(cudaMemcpyAsync(ctx->b_contig.p, R->contig, sizeof(snfb_contig) * R->n_contig, cudaMemcpyHostToDevice, ctx->st));
    if (R->n_tr) {
        // running maximum of the interval ends per task: makes the reference's forward-only scan (cluster.py:240-246) a binary search
        std::vector<int32_t> pm(R->n_tr);
        for (uint32_t t = 0; t < R->n_task; ++t) { int32_t m = INT32_MIN; for (int k = 0; k < R->task[t].tr_n; ++k) { const int idx = R->task[t].tr_off + k; if (R->tr[2 * idx + 1] > m) m = R->tr[2 * idx + 1]; pm[idx] = m; } }
        CUDA_TRY(cudaMemcpyAsync(ctx->b_tr.p, R->tr, 8 * (size_t)R->n_tr, cudaMemcpyHostToDevice, ctx->st));
        CUDA_TRY(cudaMemcpyAsync(ctx->b_trp.p, pm.data(), 4 * (size_t)R->n_tr, cudaMemcpyHostToDevice, ctx->st));
        CUDA_TRY(cudaStreamSynchronize(ctx->st));     // pm is a stack-owned staging vector
    }
    ctx->n_mask = (R->mask && R->mask_task_off) ? R->n_mask : 0;
    if (ctx->n_mask) {
        std::vector<uint32_t> mt(ctx->n_mask);
        for (uint32_t t = 0; t < R->n_task; ++t) for (uint32_t m = R->mask_task_off[t]; m < R->mask_task_off[t + 1] && m < ctx->n_mask; ++m) mt[m] = t;
        if (ctx->b_mask.ensure(8 * (size_t)ctx->n_mask) || ctx->b_mask_off.ensure(4 * ((size_t)R->n_task + 1)) || ctx->b_mask_task.ensure(4 * (size_t)ctx->n_mask)) return fail(ctx, "out of device memory (N mask)");
        CUDA_TRY(cudaMemcpyAsync(ctx->b_mask.p, R->mask, 8 * (size_t)ctx->n_mask, cudaMemcpyHostToDevice, ctx->st));
        CUDA_TRY(cudaMemcpyAsync(ctx->b_mask_off.p, R->mask_task_off, 4 * ((size_t)R->n_task + 1), cudaMemcpyHostToDevice, ctx->st));
        CUDA_TRY(cudaMemcpyAsync(ctx->b_mask_task.p, mt.data(), 4 * (size_t)ctx->n_mask, cudaMemcpyHostToDevice, ctx->st));
        CUDA_TRY(cudaStreamSynchronize(ctx->st));
    }
    mark(ctx, nullptr);
    ctx->n_ev_load = ctx->n_ev;
    ctx->loaded = true; return 0;
}

#define LAUNCHED(ctx, k) ((ctx)->launches += (k))
static int grid_for(unsigned long long n, int threads) { unsigned long long g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > 148ull * 64) g = 148ull * 64; return (int)g; }

static int fetch_counters(snfb_ctx* ctx) {
    CUDA_TRY(cudaMemcpyAsync(&ctx->h_ctr, ctx->b_ctr.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    return 0;
}

static int ensure_stage_b(snfb_ctx* ctx, unsigned long long nb) {
    // every stage-B array is bounded by the number of leads
    const size_t n = (size_t)nb + 8;
    int bad = 0;
    bad |= ctx->b_key0.ensure(8 * n) | ctx->b_val0.ensure(4 * n) | ctx->b_key1.ensure(8 * n) | ctx->b_val1.ensure(4 * n) | ctx->b_flag.ensure(4 * n) | ctx->b_scan.ensure(4 * n);
    bad |= ctx->b_hist.ensure(4 * prims::radix_hist_elems(nb)) | ctx->b_scan_tmp.ensure(4 * (prims::scan_tmp_elems(std::max<unsigned long long>(nb, (unsigned long long)prims::radix_hist_elems(nb))) + 16));
    bad |= ctx->b_bin_start.ensure(4 * n) | ctx->b_bin_nl.ensure(4 * n) | ctx->b_bin_nlong.ensure(4 * n) | ctx->b_bin_kept.ensure(4 * n) | ctx->b_bin_hap.ensure(12 * n) | ctx->b_kl_off.ensure(4 * n) | ctx->b_kll_off.ensure(4 * n) | ctx->b_kb_idx.ensure(4 * n);
    bad |= ctx->b_kl.ensure(4 * n) | ctx->b_kll.ensure(4 * n) | ctx->b_kb_bin.ensure(4 * n) | ctx->b_kb_lead_off.ensure(4 * n) | ctx->b_kb_lead_n.ensure(4 * n) | ctx->b_kb_long_off.ensure(4 * n) | ctx->b_kb_long_n.ensure(4 * n);
    bad |= ctx->b_kb_seed.ensure(4 * n) | ctx->b_kb_chain.ensure(4 * n) | ctx->b_kb_repeat.ensure(n) | ctx->b_seg_start.ensure(4 * n) | ctx->b_c_next.ensure(4 * n) | ctx->b_c_last.ensure(4 * n) | ctx->b_c_sd.ensure(8 * n) | ctx->b_c_mean.ensure(8 * n);
    bad |= ctx->b_c_rep.ensure(n) | ctx->b_seg_sd_last.ensure(8 * n) | ctx->b_seg_maxsd.ensure(8 * n) | ctx->b_cl_first.ensure(4 * n) | ctx->b_cl_last.ensure(4 * n) | ctx->b_cl_rep.ensure(n);
    bad |= ctx->b_s_hi.ensure(8 * n) | ctx->b_s_lo.ensure(8 * n) | ctx->b_s_a.ensure(4 * n) | ctx->b_s_b.ensure(4 * n) | ctx->b_s_c.ensure(4 * n) | ctx->b_s_d.ensure(4 * n) | ctx->b_s_e.ensure(4 * n) | ctx->b_ord.ensure(4 * n);
    bad |= ctx->b_ml_slot.ensure(4 * n) | ctx->b_ml_svlen.ensure(4 * n) | ctx->b_ml_seqlen.ensure(4 * n) | ctx->b_ml_plo.ensure(4 * n) | ctx->b_ml_pn.ensure(4 * n) | ctx->b_ml_has.ensure(n) | ctx->b_subl.ensure(4 * n);
    bad |= ctx->b_sub_cnt.ensure(4 * n) | ctx->b_sub_off.ensure(4 * n) | ctx->b_t_lo.ensure(4 * n) | ctx->b_t_n.ensure(4 * n) | ctx->b_t_bin.ensure(4 * n) | ctx->b_sub_cluster.ensure(4 * n) | ctx->b_sub_lo.ensure(4 * n) | ctx->b_sub_n.ensure(4 * n) | ctx->b_sub_bin.ensure(4 * n);
    bad |= ctx->b_cand_tmp.ensure(sizeof(snfb_cand) * n) | ctx->b_cand_valid.ensure(4 * n) | ctx->b_cand_id.ensure(4 * n) | ctx->b_cand_nlead.ensure(4 * n) | ctx->b_cand_lead_off.ensure(4 * n) | ctx->b_cand_nrn.ensure(4 * n) | ctx->b_cand_rn_off.ensure(4 * n);
    // outputs: candidates <= sub-clusters <= leads; their leads (incl. leads_long copies) and names can exceed the lead count only through
    // shared leads_long, so allow 2x and check on device
    ctx->cand_cap = nb + 8; ctx->cand_lead_cap = 2 * nb + 64; ctx->rn_cap = 2 * nb + 64;
    bad |= ctx->b_cand.ensure(sizeof(snfb_cand) * ctx->cand_cap) | ctx->b_cand_leads.ensure(sizeof(snfb_lead) * ctx->cand_lead_cap) | ctx->b_cand_lead_ml.ensure(4 * ctx->cand_lead_cap) | ctx->b_rnames.ensure(8 * ctx->rn_cap) | ctx->b_rn_off_out.ensure(4 * (ctx->cand_cap + 1));
    bad |= ctx->b_plan_best.ensure(4 * ctx->cand_cap) | ctx->b_plan_nother.ensure(4 * ctx->cand_cap) | ctx->b_plan_otot.ensure(4 * ctx->cand_cap) | ctx->b_alt_len.ensure(4 * ctx->cand_cap) | ctx->b_scr_len.ensure(4 * ctx->cand_cap) | ctx->b_alt_off.ensure(4 * ctx->cand_cap) | ctx->b_scr_off.ensure(4 * ctx->cand_cap);
    return bad;
}

static cluster::B make_b(snfb_ctx* ctx) {
    cluster::B b{};
    b.leads = ctx->b_leads.as<snfb_lead>(); b.rec = ctx->d_rec; b.task = ctx->b_task.as<snfb_task>(); b.contig = ctx->b_contig.as<snfb_contig>();
    b.tr = ctx->n_tr ? ctx->b_tr.as<int32_t>() : nullptr; b.tr_pmax = ctx->b_trp.as<int32_t>();
    b.rec_pos = ctx->b_rec_pos.as<int32_t>(); b.rec_end = ctx->b_rec_end.as<int32_t>(); b.rec_flags = ctx->b_rec_flags.as<uint8_t>(); b.rec_nm = ctx->b_rec_nm.as<double>();
    b.rec_nlead = ctx->b_rec_nlead.as<uint32_t>(); b.rec_lead_off = ctx->b_rec_lead_off.as<uint32_t>();
    b.task_first = ctx->b_task_first.as<uint32_t>(); b.task_last = ctx->b_task_last.as<uint32_t>(); b.task_maxspan = ctx->b_task_span.as<int32_t>();
    b.mask = ctx->n_mask ? ctx->b_mask.as<int32_t>() : nullptr; b.mask_task_off = ctx->n_mask ? ctx->b_mask_off.as<uint32_t>() : nullptr;
    b.n_task = ctx->n_task; b.n_bound = ctx->n_bound; b.ctr = ctx->b_ctr.as<DevCounters>(); b.cfg = ctx->cfg;
    b.key0 = ctx->b_key0.as<uint64_t>(); b.val0 = ctx->b_val0.as<uint32_t>(); b.key1 = ctx->b_key1.as<uint64_t>(); b.val1 = ctx->b_val1.as<uint32_t>();
    b.skey = ctx->sorted_in_first ? b.key0 : b.key1; b.sval = ctx->sorted_in_first ? b.val0 : b.val1;
    b.flag = ctx->b_flag.as<uint32_t>(); b.scan = ctx->b_scan.as<uint32_t>();
    b.bin_start = ctx->b_bin_start.as<uint32_t>(); b.bin_nl = ctx->b_bin_nl.as<uint32_t>(); b.bin_nlong = ctx->b_bin_nlong.as<uint32_t>(); b.bin_kept = ctx->b_bin_kept.as<uint32_t>(); b.bin_hap = ctx->b_bin_hap.as<uint32_t>();
    b.kl_off = ctx->b_kl_off.as<uint32_t>(); b.kll_off = ctx->b_kll_off.as<uint32_t>(); b.kb_idx = ctx->b_kb_idx.as<uint32_t>(); b.kl = ctx->b_kl.as<uint32_t>(); b.kll = ctx->b_kll.as<uint32_t>();
    b.kb_bin = ctx->b_kb_bin.as<uint32_t>(); b.kb_lead_off = ctx->b_kb_lead_off.as<uint32_t>(); b.kb_lead_n = ctx->b_kb_lead_n.as<uint32_t>(); b.kb_long_off = ctx->b_kb_long_off.as<uint32_t>(); b.kb_long_n = ctx->b_kb_long_n.as<uint32_t>();
    b.kb_seed = ctx->b_kb_seed.as<int32_t>(); b.kb_chain = ctx->b_kb_chain.as<uint32_t>(); b.kb_repeat = ctx->b_kb_repeat.as<uint8_t>();
    b.seg_start = ctx->b_seg_start.as<uint32_t>(); b.c_next = ctx->b_c_next.as<uint32_t>(); b.c_last = ctx->b_c_last.as<uint32_t>(); b.c_sd = ctx->b_c_sd.as<double>(); b.c_mean = ctx->b_c_mean.as<double>(); b.c_rep = ctx->b_c_rep.as<uint8_t>();
    b.seg_sd_last = ctx->b_seg_sd_last.as<double>(); b.seg_maxsd_first = ctx->b_seg_maxsd.as<double>(); b.cl_first = ctx->b_cl_first.as<uint32_t>(); b.cl_last = ctx->b_cl_last.as<uint32_t>(); b.cl_rep = ctx->b_cl_rep.as<uint8_t>();
    b.s_hi = ctx->b_s_hi.as<uint64_t>(); b.s_lo = ctx->b_s_lo.as<uint64_t>(); b.s_a = ctx->b_s_a.as<uint32_t>(); b.s_b = ctx->b_s_b.as<uint32_t>(); b.s_c = ctx->b_s_c.as<uint32_t>(); b.s_d = ctx->b_s_d.as<uint32_t>(); b.s_e = ctx->b_s_e.as<uint32_t>();
    b.ord = ctx->b_ord.as<uint32_t>(); b.ml_slot = ctx->b_ml_slot.as<uint32_t>(); b.ml_svlen = ctx->b_ml_svlen.as<int32_t>(); b.ml_seqlen = ctx->b_ml_seqlen.as<int32_t>(); b.ml_plo = ctx->b_ml_plo.as<uint32_t>(); b.ml_pn = ctx->b_ml_pn.as<uint32_t>();
    b.ml_has = ctx->b_ml_has.as<uint8_t>(); b.subl = ctx->b_subl.as<uint32_t>(); b.sub_cnt = ctx->b_sub_cnt.as<uint32_t>(); b.sub_off = ctx->b_sub_off.as<uint32_t>();
    b.t_lo = ctx->b_t_lo.as<uint32_t>(); b.t_n = ctx->b_t_n.as<uint32_t>(); b.t_bin = ctx->b_t_bin.as<int32_t>(); b.sub_cluster = ctx->b_sub_cluster.as<uint32_t>(); b.sub_lo = ctx->b_sub_lo.as<uint32_t>(); b.sub_n = ctx->b_sub_n.as<uint32_t>(); b.sub_bin = ctx->b_sub_bin.as<int32_t>();
    b.cand_tmp = ctx->b_cand_tmp.as<snfb_cand>(); b.cand_valid = ctx->b_cand_valid.as<uint32_t>(); b.cand_id = ctx->b_cand_id.as<uint32_t>(); b.cand_nlead = ctx->b_cand_nlead.as<uint32_t>(); b.cand_lead_off = ctx->b_cand_lead_off.as<uint32_t>();
    b.cand_nrn = ctx->b_cand_nrn.as<uint32_t>(); b.cand_rn_off = ctx->b_cand_rn_off.as<uint32_t>(); b.cand = ctx->b_cand.as<snfb_cand>(); b.cand_leads = ctx->b_cand_leads.as<snfb_lead>(); b.cand_lead_ml = ctx->b_cand_lead_ml.as<uint32_t>();
    b.rnames = ctx->b_rnames.as<uint64_t>(); b.rn_off_out = ctx->b_rn_off_out.as<uint32_t>();
    b.cand_cap = ctx->cand_cap; b.cand_lead_cap = ctx->cand_lead_cap; b.rn_cap = ctx->rn_cap; b.scan_tmp = ctx->b_scan_tmp.as<uint32_t>();
    return b;
}

static int bits_for(uint32_t n) { int b = 0; while ((1ull << b) < n) ++b; return b; }

// stage A plus the bin sort: everything LeadProvider.build_leadtab leaves behind
static int run_stage_a(snfb_ctx* ctx) {
    if (!ctx->loaded) return fail(ctx, "no records loaded");
    if (!ctx->have_cfg) return fail(ctx, "snfb_set_config was not called");
    cudaSetDevice(ctx->device);
    const uint64_t nrec = ctx->n_rec; const uint32_t nt = ctx->n_task;
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (ctx->lead_cap == 0) ctx->lead_cap = std::max<unsigned long long>(1ull << 16, nrec * 2) + (unsigned long long)extract::SA_BLOCKS * extract::SA_THREADS * extract::SLOT_CHUNK;
        int bad = ctx->b_ctr.ensure(sizeof(DevCounters)) | ctx->b_leads.ensure(sizeof(snfb_lead) * ctx->lead_cap) | ctx->b_rec_pos.ensure(4 * (nrec + 1)) | ctx->b_rec_end.ensure(4 * (nrec + 1)) | ctx->b_rec_flags.ensure(nrec + 1)
                | ctx->b_rec_nm.ensure(8 * (nrec + 1)) | ctx->b_rec_nlead.ensure(4 * (nrec + 1)) | ctx->b_rec_lead_off.ensure(4 * (nrec + 1)) | ctx->b_task_first.ensure(4 * nt) | ctx->b_task_last.ensure(4 * nt)
                | ctx->b_task_reads.ensure(4 * nt) | ctx->b_task_cov.ensure(8 * nt) | ctx->b_task_span.ensure(4 * nt) | ctx->b_task_nm.ensure(8 * nt) | ctx->b_scan_tmp.ensure(4 * (prims::scan_tmp_elems(nrec) + 16));
        if (bad) return fail(ctx, "out of device memory (stage A)");
        CUDA_TRY(cudaMemsetAsync(ctx->b_ctr.p, 0, sizeof(DevCounters), ctx->st));
        CUDA_TRY(cudaMemsetAsync(ctx->b_task_first.p, 0, 4 * nt, ctx->st)); CUDA_TRY(cudaMemsetAsync(ctx->b_task_last.p, 0, 4 * nt, ctx->st));
        CUDA_TRY(cudaMemsetAsync(ctx->b_task_reads.p, 0, 4 * nt, ctx->st)); CUDA_TRY(cudaMemsetAsync(ctx->b_task_cov.p, 0, 8 * nt, ctx->st)); CUDA_TRY(cudaMemsetAsync(ctx->b_task_span.p, 0, 4 * nt, ctx->st));
        const snfb_config& cf = ctx->cfg;
        if (ctx->b_ev.ensure(sizeof(extract::Event) * ctx->lead_cap) || ctx->b_sa_list.ensure(4 * (nrec + 1))
            || ctx->b_scan_tmp.ensure(4 * (prims::scan_tmp_elems(std::max<unsigned long long>(nrec, ctx->lead_cap)) + 16))) return fail(ctx, "out of device memory (stage A lists)");
        DevCounters* ctr = ctx->b_ctr.as<DevCounters>();
        if (ctx->b_scanrec.ensure(sizeof(extract::RecScan) * (nrec + 1)) || ctx->b_clip.ensure(sizeof(extract::RecClip) * (nrec + 1)) || ctx->b_rec_big.ensure(4 * (nrec + 1))) return fail(ctx, "out of device memory (record descriptors)");
        extract::ScanParams S{};
        S.scan = ctx->b_scanrec.as<extract::RecScan>(); S.cigar = ctx->d_cigar; S.task = ctx->b_task.as<snfb_task>(); S.n_rec = (uint32_t)nrec;
        S.rec_end = ctx->b_rec_end.as<int32_t>(); S.rec_nlead = ctx->b_rec_nlead.as<uint32_t>(); S.rec_big = ctx->b_rec_big.as<int32_t>();
        S.ev = ctx->b_ev.as<extract::Event>(); S.ev_cap = ctx->lead_cap; S.n_ev = &ctr->n_ev; S.sa_list = ctx->b_sa_list.as<uint32_t>(); S.n_sa = &ctr->n_sa; S.ctr = ctr;
        S.minsv = cf.minsvlen_screen;
        { const bool want_nm = cf.qc_nm_measure || cf.phase; int t = cf.minsvlen_screen < 1 ? 1 : cf.minsvlen_screen; if (want_nm && t > 11) t = 11; if (t > 0x1000) t = 0x1000;
          S.gt_add = (uint32_t)(0x1000 - t) * 0x00010001u; }
        uint32_t* task_first = ctx->b_task_first.as<uint32_t>(); uint32_t* task_last = ctx->b_task_last.as<uint32_t>();
        uint8_t* rec_flags = ctx->b_rec_flags.as<uint8_t>(); double* rec_nm = ctx->b_rec_nm.as<double>();
        mark(ctx, "k_rec_index");
        if (nrec) {
            extract::IndexParams I{};
            I.rec = ctx->d_rec; I.cigar = ctx->d_cigar; I.task = S.task; I.n_rec = (uint32_t)nrec; I.rec_pos = ctx->b_rec_pos.as<int32_t>(); I.task_first = task_first; I.task_last = task_last;
            I.scan = ctx->b_scanrec.as<extract::RecScan>(); I.clip = ctx->b_clip.as<extract::RecClip>(); I.rec_end = S.rec_end; I.rec_flags = rec_flags; I.rec_nm = rec_nm; I.rec_nlead = S.rec_nlead; I.ctr = ctr;
            I.mapq_min = cf.mapq; I.alen_min = cf.min_alignment_length; I.excl = cf.exclude_flags; I.want_nm = (cf.qc_nm_measure || cf.phase) ? 1 : 0;
            extract::k_rec_index<<<(unsigned)((nrec + 255) / 256), 256, 0, ctx->st>>>(I);
            // algorithmic bytes of the streaming kernel: scan descriptors + CIGAR16 words (+ its per-record outputs and event slices, added by bench.py)
            mark(ctx, "k_scan", sizeof(extract::RecScan) * nrec + 2 * ctx->n_cigar);
            unsigned long long blocks = (nrec + 7) / 8; const unsigned long long maxb = 148ull * 6 * 4;
            extract::k_scan<<<(int)std::min(blocks, maxb), 256, 0, ctx->st>>>(S);
            mark(ctx, "k_rec_post");
            extract::PostParams Q{};
            Q.scan = I.scan; Q.clip = I.clip; Q.task = S.task; Q.n_rec = (uint32_t)nrec; Q.rec_end = S.rec_end; Q.rec_big = S.rec_big; Q.rec_nm = rec_nm;
            Q.task_reads = ctx->b_task_reads.as<uint32_t>(); Q.task_cov_bp = ctx->b_task_cov.as<unsigned long long>(); Q.task_maxspan = ctx->b_task_span.as<int32_t>();
            extract::k_rec_post<<<(unsigned)((nrec + 255) / 256), 256, 0, ctx->st>>>(Q); LAUNCHED(ctx, 1);
            mark(ctx, "k_emit");
            extract::EmitParams E{};
            E.rec = ctx->d_rec; E.clip = ctx->b_clip.as<extract::RecClip>(); E.var = ctx->d_var; E.ev = S.ev; E.n_ev = S.n_ev; E.ev_cap = ctx->lead_cap;
            E.leads = ctx->b_leads.as<snfb_lead>(); E.ctr = ctr; E.maxlen = cf.dev_seq_cache_maxlen; E.detect_large_ins = cf.detect_large_ins;
            E.longinslen = (double)cf.long_ins_length / 2.0;
            extract::k_emit<<<148 * 16, 128, 0, ctx->st>>>(E); LAUNCHED(ctx, 3);      // thread per event; k_rec_index and k_scan are counted here too
            mark(ctx, "k_sa");
            extract::SaParams A{};
            A.rec = ctx->d_rec; A.clip = ctx->b_clip.as<extract::RecClip>(); A.var = ctx->d_var; A.task = S.task; A.contig = ctx->b_contig.as<snfb_contig>(); A.n_contig = ctx->n_contig;
            A.sa_list = S.sa_list; A.n_sa = S.n_sa; A.rec_end = S.rec_end; A.rec_nlead = S.rec_nlead; A.leads = E.leads; A.lead_cap = ctx->lead_cap; A.ctr = ctr; A.cfg = cf;
            if (ctx->b_sa_seg.ensure(sizeof(extract::Seg) * (size_t)extract::MAXSEG * extract::SA_THREADS * extract::SA_BLOCKS)) return fail(ctx, "out of device memory (split segments)");
            A.seg_scratch = ctx->b_sa_seg.as<extract::Seg>();
            extract::k_sa<<<extract::SA_BLOCKS, extract::SA_THREADS, 0, ctx->st>>>(A);
            mark(ctx, "k_task_nm");
            const int cpt = (int)((nrec + extract::NM_CHUNK - 1) / extract::NM_CHUNK);
            if (ctx->b_nm_part.ensure(8 * (size_t)cpt * nt + 8) || ctx->b_nm_cnt.ensure(4 * (size_t)cpt * nt + 8)) return fail(ctx, "out of device memory (nm partials)");
            extract::k_nm_partial<<<dim3(cpt, nt), 256, 0, ctx->st>>>(rec_flags, rec_nm, task_first, task_last, ctx->b_nm_part.as<double>(), ctx->b_nm_cnt.as<unsigned>());
            extract::k_task_nm<<<nt, 256, 0, ctx->st>>>(task_first, task_last, ctx->b_nm_part.as<double>(), ctx->b_nm_cnt.as<unsigned>(), cpt, ctx->b_task_nm.as<double>()); LAUNCHED(ctx, 5);
        }
        mark(ctx, "scan_rec_leads");
        LAUNCHED(ctx, prims::exclusive_scan(S.rec_nlead, ctx->b_rec_lead_off.as<uint32_t>(), ctx->b_scan_tmp.as<uint32_t>(), nullptr, nrec, &ctr->n_leads, ctx->st));
        mark(ctx, nullptr);
        CUDA_TRY(cudaGetLastError());
        if (fetch_counters(ctx)) return 1;
        if (ctx->h_ctr.unsorted) return fail(ctx, "records are not coordinate sorted inside a task");
        if (ctx->h_ctr.lead_overflow == 0 && ctx->h_ctr.n_slots <= ctx->lead_cap) break;
        ctx->lead_cap = ctx->h_ctr.n_slots + ctx->h_ctr.n_slots / 4 + 1024;       // retry with a buffer that fits
        ctx->n_ev = ctx->n_ev_load;
        if (attempt == 2) return fail(ctx, "lead buffer overflow");
    }
    ctx->n_bound = ctx->h_ctr.n_leads;
    if (ensure_stage_b(ctx, ctx->n_bound)) return fail(ctx, "out of device memory (stage B)");
    cluster::B b = make_b(ctx);
    const unsigned long long nb = ctx->n_bound; const int g = grid_for(nb, 256);
    if (nb) {
        mark(ctx, "sort_leads", nb * (sizeof(snfb_lead) + 24));
        cluster::k_scatter_keys<<<grid_for(ctx->h_ctr.n_slots, 256), 256, 0, ctx->st>>>(b, ctx->h_ctr.n_slots);
        prims::RadixTemp rt{ ctx->b_hist.as<uint32_t>(), ctx->b_scan_tmp.as<uint32_t>() };
        bool first = true;
        LAUNCHED(ctx, 1 + 3); LAUNCHED(ctx, prims::radix_sort(b.key0, b.val0, b.key1, b.val1, rt, &b.ctr->n_leads, nb, cluster::TASK_SHIFT + bits_for(ctx->n_task), &first, ctx->st));
        ctx->sorted_in_first = first; b = make_b(ctx);
        mark(ctx, "bins");
        cluster::k_bin_heads<<<g, 256, 0, ctx->st>>>(b);
        LAUNCHED(ctx, prims::exclusive_scan(b.flag, b.scan, b.scan_tmp, nullptr, nb, &b.ctr->n_bins, ctx->st));
        cluster::k_bin_build<<<g, 256, 0, ctx->st>>>(b);
        cluster::k_bin_stats<<<g, 256, 0, ctx->st>>>(b);
        mark(ctx, nullptr);
    } else CUDA_TRY(cudaMemsetAsync(&b.ctr->n_bins, 0, 8, ctx->st));
    CUDA_TRY(cudaGetLastError());
    ctx->stage_a_done = true; ctx->stage_b_done = false;
    return 0;
}

__global__ void k_gather_leads(const snfb_lead* __restrict__ leads, const uint32_t* __restrict__ sval, snfb_lead* __restrict__ out, unsigned long long n) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) out[i] = leads[sval[i]];
}
__global__ void k_copy_ml(const uint32_t* __restrict__ subl, const uint32_t* __restrict__ sub_lo, const uint32_t* __restrict__ cand_valid, const uint32_t* __restrict__ cand_lead_off,
                          const snfb_cand* __restrict__ cand_tmp, uint32_t* __restrict__ cand_lead_ml, const unsigned long long* n_sub, unsigned long long cap) {
    for (unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; s < *n_sub; s += (unsigned long long)gridDim.x * blockDim.x) {
        if (!cand_valid[s]) continue;
        const uint32_t lo = cand_lead_off[s]; const int n = cand_tmp[s].lead_n;
        for (int i = 0; i < n; ++i) if ((unsigned long long)lo + i < cap) cand_lead_ml[lo + i] = subl[sub_lo[s] + i];
    }
}

static int fill_lead_view(snfb_ctx* ctx, snfb_lead_view* out) {
    const unsigned long long nl = ctx->n_bound; const uint32_t nt = ctx->n_task;
    if (ctx->b_sorted_leads.ensure(sizeof(snfb_lead) * (nl + 1)) || ctx->h_leads.ensure(sizeof(snfb_lead) * (nl + 1)) || ctx->h_task_reads.ensure(4 * nt) || ctx->h_task_nm.ensure(8 * nt) || ctx->h_rec_nm.ensure(8 * (ctx->n_rec + 1)))
        return fail(ctx, "out of memory for the lead view");
    if (nl) {
        cluster::B b = make_b(ctx);
        k_gather_leads<<<grid_for(nl, 256), 256, 0, ctx->st>>>(b.leads, b.sval, ctx->b_sorted_leads.as<snfb_lead>(), nl);
        CUDA_TRY(cudaMemcpyAsync(ctx->h_leads.p, ctx->b_sorted_leads.p, sizeof(snfb_lead) * nl, cudaMemcpyDeviceToHost, ctx->st));
    }
    CUDA_TRY(cudaMemcpyAsync(ctx->h_task_reads.p, ctx->b_task_reads.p, 4 * nt, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaMemcpyAsync(ctx->h_task_nm.p, ctx->b_task_nm.p, 8 * nt, cudaMemcpyDeviceToHost, ctx->st));
    if (ctx->n_rec) CUDA_TRY(cudaMemcpyAsync(ctx->h_rec_nm.p, ctx->b_rec_nm.p, 8 * ctx->n_rec, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    out->n_leads = nl; out->leads = ctx->h_leads.as<snfb_lead>();
    out->task_read_count = ctx->h_task_reads.as<uint32_t>(); out->task_mean_nm = ctx->h_task_nm.as<double>(); out->rec_nm = ctx->h_rec_nm.as<double>();
    uint64_t np = 0; for (uint32_t t = 0; t < nt; ++t) np += out->task_read_count[t];
    out->n_pass = np; out->soft_errors = ctx->h_ctr.soft_errors;
    return 0;
}

static int run_stage_b(snfb_ctx* ctx) {
    if (!ctx->stage_a_done) return fail(ctx, "snfb_extract_leads must run first");
    cudaSetDevice(ctx->device);
    cluster::B b = make_b(ctx);
    const unsigned long long nb = ctx->n_bound; const int g = grid_for(nb, 128);
    DevCounters* ctr = b.ctr;
    if (nb) {
        mark(ctx, "kept_bins");
        LAUNCHED(ctx, prims::exclusive_scan(b.bin_nl, b.kl_off, b.scan_tmp, nullptr, nb, nullptr, ctx->st));
        LAUNCHED(ctx, prims::exclusive_scan(b.bin_nlong, b.kll_off, b.scan_tmp, nullptr, nb, nullptr, ctx->st));
        LAUNCHED(ctx, prims::exclusive_scan(b.bin_kept, b.kb_idx, b.scan_tmp, nullptr, nb, &ctr->n_kbins, ctx->st));
        cluster::k_kbin_build<<<g, 128, 0, ctx->st>>>(b);
        mark(ctx, "merge_chains");
        cluster::k_seg_heads<<<g, 128, 0, ctx->st>>>(b);
        LAUNCHED(ctx, prims::exclusive_scan(b.flag, b.scan, b.scan_tmp, nullptr, nb, &ctr->n_segs, ctx->st));
        cluster::k_seg_build<<<g, 128, 0, ctx->st>>>(b);
        cluster::k_merge<<<g, 128, 0, ctx->st>>>(b);
        cluster::k_verify_cuts<<<g, 128, 0, ctx->st>>>(b);
        LAUNCHED(ctx, prims::exclusive_scan(b.flag, b.scan, b.scan_tmp, nullptr, nb, &ctr->n_clusters, ctx->st));
        cluster::k_cluster_build<<<g, 128, 0, ctx->st>>>(b);
        mark(ctx, "cluster_post");
        cluster::k_cluster_post<<<g, 128, 0, ctx->st>>>(b);
        LAUNCHED(ctx, prims::exclusive_scan(b.sub_cnt, b.sub_off, b.scan_tmp, nullptr, nb, &ctr->n_sub, ctx->st));
        cluster::k_sub_build<<<g, 128, 0, ctx->st>>>(b);
        mark(ctx, "call_from");
        cluster::k_call<<<g, 128, 0, ctx->st>>>(b);
        LAUNCHED(ctx, prims::exclusive_scan(b.cand_valid, b.cand_id, b.scan_tmp, nullptr, nb, &ctr->n_cand, ctx->st));
        LAUNCHED(ctx, prims::exclusive_scan(b.cand_nlead, b.cand_lead_off, b.scan_tmp, nullptr, nb, &ctr->n_cand_leads, ctx->st));
        LAUNCHED(ctx, prims::exclusive_scan(b.cand_nrn, b.cand_rn_off, b.scan_tmp, nullptr, nb, &ctr->n_rnames, ctx->st));
        mark(ctx, "cand_finish");
        cluster::k_cand_finish<<<g, 128, 0, ctx->st>>>(b);
        k_copy_ml<<<g, 128, 0, ctx->st>>>(b.subl, b.sub_lo, b.cand_valid, b.cand_lead_off, b.cand_tmp, b.cand_lead_ml, &ctr->n_sub, ctx->cand_lead_cap);
        mark(ctx, "coverage");
        if (ctx->n_mask) { cluster::k_mask_bp<<<grid_for((unsigned long long)ctx->n_mask * 32, 128), 128, 0, ctx->st>>>(b, ctx->b_mask_task.as<uint32_t>(), ctx->n_mask, ctx->b_task_cov.as<unsigned long long>()); LAUNCHED(ctx, 1); }
        cluster::k_coverage<<<g, 128, 0, ctx->st>>>(b); LAUNCHED(ctx, 12);
        mark(ctx, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    ctx->stage_b_done = true;
    return 0;
}

// the bulk of the candidate view (lead table, read names) is final once stage B is done: in snfb_run it is copied on its own
// stream while the consensus kernels run
static int cand_bulk_copy(snfb_ctx* ctx, cudaStream_t stream) {
    const DevCounters& c = ctx->h_ctr;
    if (c.n_cand > ctx->cand_cap || c.n_cand_leads > ctx->cand_lead_cap || c.n_rnames > ctx->rn_cap) return fail(ctx, "candidate output buffers too small");
    if (ctx->h_cand.ensure(sizeof(snfb_cand) * (c.n_cand + 1)) || ctx->h_cand_leads.ensure(sizeof(snfb_lead) * (c.n_cand_leads + 1)) || ctx->h_rnames.ensure(8 * (c.n_rnames + 1)) || ctx->h_rn_off.ensure(4 * (c.n_cand + 2)) || ctx->h_task_cov.ensure(8 * ctx->n_task))
        return fail(ctx, "out of pinned memory for the candidate view");
    if (c.n_cand) CUDA_TRY(cudaMemcpyAsync(ctx->h_rn_off.p, ctx->b_rn_off_out.p, 4 * c.n_cand, cudaMemcpyDeviceToHost, stream));
    if (c.n_cand_leads) CUDA_TRY(cudaMemcpyAsync(ctx->h_cand_leads.p, ctx->b_cand_leads.p, sizeof(snfb_lead) * c.n_cand_leads, cudaMemcpyDeviceToHost, stream));
    if (c.n_rnames) CUDA_TRY(cudaMemcpyAsync(ctx->h_rnames.p, ctx->b_rnames.p, 8 * c.n_rnames, cudaMemcpyDeviceToHost, stream));
    return 0;
}

static int fill_cand_view(snfb_ctx* ctx, snfb_cand_view* out, bool cand_copied_later = false) {
    if (fetch_counters(ctx)) return 1;
    const DevCounters& c = ctx->h_ctr; const uint32_t nt = ctx->n_task;
    if (c.scratch_overflow) return fail(ctx, "candidate output buffers overflowed");
    if (ctx->cand_prefetched) { CUDA_TRY(cudaStreamSynchronize(ctx->st_copy)); ctx->cand_prefetched = false; }
    else if (cand_bulk_copy(ctx, ctx->st)) return 1;
    if (c.n_cand && !cand_copied_later) CUDA_TRY(cudaMemcpyAsync(ctx->h_cand.p, ctx->b_cand.p, sizeof(snfb_cand) * c.n_cand, cudaMemcpyDeviceToHost, ctx->st));
    std::vector<unsigned long long> cov(nt);
    CUDA_TRY(cudaMemcpyAsync(cov.data(), ctx->b_task_cov.p, 8 * nt, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    ctx->h_rn_off.as<uint32_t>()[c.n_cand] = (uint32_t)c.n_rnames;
    // coverage_average_total: integer base-pair sum / contig length, one rounding (postprocessing.py:130)
    double* cm = ctx->h_task_cov.as<double>();
    for (uint32_t t = 0; t < nt; ++t) cm[t] = ctx->tasks[t].contig_len > 0 ? (double)cov[t] / (double)ctx->tasks[t].contig_len : 0.0;
    out->n_cand = c.n_cand; out->cand = ctx->h_cand.as<snfb_cand>(); out->n_cand_leads = c.n_cand_leads; out->cand_leads = ctx->h_cand_leads.as<snfb_lead>();
    out->rnames = ctx->h_rnames.as<uint64_t>(); out->rnames_off = ctx->h_rn_off.as<uint32_t>(); out->task_coverage_mean = cm; out->unverified_breaks = c.unverified_breaks;
    if (c.unverified_breaks) return fail(ctx, "a chain cut could not be verified (cluster stdev too large); rerun with a larger cut distance");
    return 0;
}

static int run_stage_c(snfb_ctx* ctx) {
    if (!ctx->stage_b_done) return fail(ctx, "snfb_cluster_call must run first");
    cudaSetDevice(ctx->device);
    if (ctx->n_bound == 0) return 0;
    consensus::C c{};
    c.cand = ctx->b_cand.as<snfb_cand>(); c.cand_rw = ctx->b_cand.as<snfb_cand>(); c.cand_leads = ctx->b_cand_leads.as<snfb_lead>(); c.cand_lead_ml = ctx->b_cand_lead_ml.as<uint32_t>();
    c.ml_plo = ctx->b_ml_plo.as<uint32_t>(); c.ml_pn = ctx->b_ml_pn.as<uint32_t>(); c.ord = ctx->b_ord.as<uint32_t>(); c.leads = ctx->b_leads.as<snfb_lead>(); c.rec = ctx->d_rec; c.seq = ctx->d_seq;
    c.plan_best = ctx->b_plan_best.as<uint32_t>(); c.plan_nother = ctx->b_plan_nother.as<uint32_t>(); c.plan_otot = ctx->b_plan_otot.as<uint32_t>(); c.alt_len = ctx->b_alt_len.as<uint32_t>(); c.scr_len = ctx->b_scr_len.as<uint32_t>();
    c.alt_off = ctx->b_alt_off.as<uint32_t>(); c.scr_off = ctx->b_scr_off.as<uint32_t>(); c.cand_cap = ctx->cand_cap; c.ctr = ctx->b_ctr.as<DevCounters>(); c.cfg = ctx->cfg;
    c.item_cap = ctx->cand_lead_cap; c.tile_cap = ctx->cand_cap + ctx->cand_lead_cap / 8 + 1024;
    if (ctx->b_work_big.ensure(4 * ctx->cand_cap) || ctx->b_work_small.ensure(4 * ctx->cand_cap) || ctx->b_work_ctr.ensure(64) || ctx->b_items_big.ensure(16 * c.item_cap) || ctx->b_items_small.ensure(16 * c.item_cap)
        || ctx->b_tiles.ensure(8 * c.tile_cap)) return fail(ctx, "out of device memory (consensus queue)");
    c.work_big = ctx->b_work_big.as<uint32_t>(); c.work_small = ctx->b_work_small.as<uint32_t>(); c.work_ctr = ctx->b_work_ctr.as<uint32_t>();
    c.items_big = ctx->b_items_big.as<consensus::C::Item>(); c.items_small = ctx->b_items_small.as<consensus::C::Item>(); c.tiles = ctx->b_tiles.as<uint2>();
    CUDA_TRY(cudaMemsetAsync(c.work_ctr, 0, 64, ctx->st));
    mark(ctx, "consensus_plan");
    consensus::k_plan<<<grid_for(ctx->cand_cap, 128), 128, 0, ctx->st>>>(c); LAUNCHED(ctx, 1);
    LAUNCHED(ctx, prims::exclusive_scan(c.alt_len, c.alt_off, ctx->b_scan_tmp.as<uint32_t>(), nullptr, ctx->cand_cap, &c.ctr->n_alt_bytes, ctx->st));
    LAUNCHED(ctx, prims::exclusive_scan(c.scr_len, c.scr_off, ctx->b_scan_tmp.as<uint32_t>(), nullptr, ctx->cand_cap, &c.ctr->n_seq_bytes, ctx->st));
    mark(ctx, nullptr);
    if (fetch_counters(ctx)) return 1;
    if (ctx->want_cand_prefetch) {        // stage B's outputs are final: start their device -> host copy next to the consensus kernels
        CUDA_TRY(cudaEventRecord(ctx->ev_copy, ctx->st)); CUDA_TRY(cudaStreamWaitEvent(ctx->st_copy, ctx->ev_copy, 0));
        if (cand_bulk_copy(ctx, ctx->st_copy)) return 1;
        ctx->cand_prefetched = true;
    }
    if (ctx->h_ctr.n_seq_bytes > 0xfffffff0ull) return fail(ctx, "consensus scratch exceeds 64 GiB");
    if (ctx->b_alt.ensure(ctx->h_ctr.n_alt_bytes + 16) || ctx->b_scr.ensure(ctx->h_ctr.n_seq_bytes * 16 + 64)) return fail(ctx, "out of device memory (consensus)");
    c.alt = ctx->b_alt.as<uint8_t>(); c.scr = ctx->b_scr.as<uint8_t>(); c.alt_cap = ctx->h_ctr.n_alt_bytes; c.scr_cap16 = ctx->h_ctr.n_seq_bytes;
    c.arena_off = nullptr;
    if (ctx->seq_on_demand && ctx->h_ctr.n_cand) {
        // seq on demand: the device lists the base slices it will read, the host gathers exactly those bytes from its
        // arena into pinned staging, one H2D copy brings them in (PCIe bytes ~ algorithmic bytes instead of the whole arena)
        mark(ctx, "seq_requests");
        const unsigned long long req_cap = ctx->cand_lead_cap + 64;
        if (ctx->b_seq_req.ensure(sizeof(consensus::SeqReq) * req_cap) || ctx->b_arena_off.ensure(4 * (ctx->lead_cap + 8))) return fail(ctx, "out of device memory (seq requests)");
        DevCounters* ctr = ctx->b_ctr.as<DevCounters>();
        CUDA_TRY(cudaMemsetAsync(&ctr->n_ev, 0, 16, ctx->st));          // n_ev / n_sa are free again after stage A: reuse as request / unit counters
        consensus::k_seq_requests<<<grid_for(ctx->h_ctr.n_cand, 128), 128, 0, ctx->st>>>(c, ctx->b_seq_req.as<consensus::SeqReq>(), req_cap, ctx->b_arena_off.as<uint32_t>(), &ctr->n_ev, &ctr->n_sa); LAUNCHED(ctx, 1);
        mark(ctx, nullptr);
        if (fetch_counters(ctx)) return 1;
        const unsigned long long nreq = ctx->h_ctr.n_ev, nunits = ctx->h_ctr.n_sa;
        if (nreq > req_cap) return fail(ctx, "seq request list overflow");
        if (ctx->h_seq_req.ensure(sizeof(consensus::SeqReq) * (nreq + 1)) || ctx->h_seq_arena.ensure(nunits * 16 + 64) || ctx->b_seq_arena.ensure(nunits * 16 + 64)) return fail(ctx, "out of memory (seq arena)");
        if (nreq) {
            CUDA_TRY(cudaMemcpyAsync(ctx->h_seq_req.p, ctx->b_seq_req.p, sizeof(consensus::SeqReq) * nreq, cudaMemcpyDeviceToHost, ctx->st));
            CUDA_TRY(cudaStreamSynchronize(ctx->st));
            const consensus::SeqReq* rq = ctx->h_seq_req.as<consensus::SeqReq>(); uint8_t* dst = ctx->h_seq_arena.as<uint8_t>(); const uint8_t* src = ctx->h_seq; const uint64_t nseq = ctx->n_seq;
            #pragma omp parallel for schedule(static, 256)
            for (long long i = 0; i < (long long)nreq; ++i) {
                unsigned long long s0 = rq[i].src; unsigned long long nb = rq[i].nbytes; if (s0 > nseq) s0 = nseq; if (s0 + nb > nseq) nb = nseq - s0;
                memcpy(dst + (size_t)rq[i].dst16 * 16, src + s0, (size_t)nb);
            }
            mark(ctx, "h2d_seq_slices", nunits * 16);
            CUDA_TRY(cudaMemcpyAsync(ctx->b_seq_arena.p, ctx->h_seq_arena.p, nunits * 16, cudaMemcpyHostToDevice, ctx->st));
            mark(ctx, nullptr);
        }
        ctx->seq_h2d_bytes = nunits * 16;
        c.seq = ctx->b_seq_arena.as<uint8_t>(); c.arena_off = ctx->b_arena_off.as<uint32_t>();
    }
    if (ctx->h_ctr.n_cand) {
        mark(ctx, "consensus", ctx->h_ctr.n_seq_bytes * 16);
        consensus::k_prep<<<148 * 8, 128, 0, ctx->st>>>(c);
        mark(ctx, "consensus_align");
        consensus::k_align<<<148 * 8, consensus::ALIGN_WARPS * 32, 0, ctx->st>>>(c);
        mark(ctx, "consensus_vote");
        consensus::k_vote<<<148 * 16, consensus::VOTE_THREADS, 0, ctx->st>>>(c); LAUNCHED(ctx, 3);
        mark(ctx, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int fill_seq_view(snfb_ctx* ctx, snfb_seq_view* out, snfb_cand_view* cands) {
    const unsigned long long na = ctx->h_ctr.n_alt_bytes;
    if (ctx->h_alt.ensure(na + 16)) return fail(ctx, "out of pinned memory for the ALT arena");
    if (na) CUDA_TRY(cudaMemcpyAsync(ctx->h_alt.p, ctx->b_alt.p, na, cudaMemcpyDeviceToHost, ctx->st));
    (void)cands;
    if (ctx->h_ctr.n_cand && ctx->h_cand.p && ctx->h_cand.cap >= sizeof(snfb_cand) * ctx->h_ctr.n_cand) CUDA_TRY(cudaMemcpyAsync(ctx->h_cand.p, ctx->b_cand.p, sizeof(snfb_cand) * ctx->h_ctr.n_cand, cudaMemcpyDeviceToHost, ctx->st));   // alt_off/alt_len
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    if (fetch_counters(ctx)) return 1;
    if (ctx->h_ctr.scratch_overflow) return fail(ctx, "consensus buffers overflowed");
    if (out) { out->n_alt_bytes = na; out->alt = ctx->h_alt.as<uint8_t>(); }
    return 0;
}

int snfb_extract_leads(snfb_ctx* ctx, snfb_lead_view* out) {
    if (!ctx) return 1;
    ctx->n_ev = ctx->n_ev_load;          // keep the h2d interval of the last load
    if (run_stage_a(ctx)) return 1;
    if (out) { memset(out, 0, sizeof *out); if (fill_lead_view(ctx, out)) return 1; }
    return 0;
}
int snfb_cluster_call(snfb_ctx* ctx, snfb_cand_view* out) {
    if (!ctx) return 1;
    if (run_stage_b(ctx)) return 1;
    if (out) { memset(out, 0, sizeof *out); if (fill_cand_view(ctx, out)) return 1; }
    return 0;
}
int snfb_consensus(snfb_ctx* ctx, snfb_seq_view* out) {
    if (!ctx) return 1;
    if (run_stage_c(ctx)) return 1;
    if (out) { memset(out, 0, sizeof *out); }
    return fill_seq_view(ctx, out, nullptr);
}
int snfb_run(snfb_ctx* ctx, snfb_lead_view* leads, snfb_cand_view* cands, snfb_seq_view* seqs) {
    if (!ctx) return 1;
    ctx->n_ev = ctx->n_ev_load;
    ctx->want_cand_prefetch = cands != nullptr; ctx->cand_prefetched = false;
    const int rc = run_stage_a(ctx) || run_stage_b(ctx) || run_stage_c(ctx);
    ctx->want_cand_prefetch = false;
    if (rc) { if (ctx->cand_prefetched) { cudaStreamSynchronize(ctx->st_copy); ctx->cand_prefetched = false; } return 1; }
    if (leads) { memset(leads, 0, sizeof *leads); if (fill_lead_view(ctx, leads)) return 1; }
    if (cands) { memset(cands, 0, sizeof *cands); if (fill_cand_view(ctx, cands, true)) return 1; }
    if (seqs) memset(seqs, 0, sizeof *seqs);
    if (fill_seq_view(ctx, seqs, cands)) return 1;
    if (!cands) { if (ctx->h_ctr.unverified_breaks) return fail(ctx, "a chain cut could not be verified"); }
    return 0;
}

int snfb_last_timings(snfb_ctx* ctx, const char** names, float* ms, uint64_t* bytes, int cap) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    int n = 0;
    for (int i = 0; i + 1 < ctx->n_ev && n < cap; ++i) {
        if (!ctx->ev_name[i]) continue;
        float t = 0; if (cudaEventElapsedTime(&t, ctx->ev[i], ctx->ev[i + 1]) != cudaSuccess) t = -1.f;
        names[n] = ctx->ev_name[i]; ms[n] = t; if (bytes) bytes[n] = ctx->ev_bytes[i]; ++n;
    }
    if (ctx->n_ev - ctx->n_ev_load >= 2 && n < cap) {   // first stage mark .. last mark: device time of the run including its host round trips
        float t = 0; if (cudaEventElapsedTime(&t, ctx->ev[ctx->n_ev_load], ctx->ev[ctx->n_ev - 1]) != cudaSuccess) t = -1.f;
        names[n] = "total"; ms[n] = t; if (bytes) bytes[n] = 0; ++n;
    }
    return n;
}

int snfb_device_candidates(snfb_ctx* ctx, void** dptr, uint64_t* n_cand) {
    if (!ctx || !ctx->stage_b_done) return 1;
    if (dptr) *dptr = ctx->b_cand.p; if (n_cand) *n_cand = ctx->h_ctr.n_cand; return 0;
}

int snfb_device_alt(snfb_ctx* ctx, void** dptr, uint64_t* n_bytes) {
    if (!ctx || !ctx->stage_b_done) return 1;
    if (dptr) *dptr = ctx->b_alt.p; if (n_bytes) *n_bytes = ctx->h_ctr.n_alt_bytes; return 0;
}
uint64_t snfb_launch_count(snfb_ctx* ctx) { return ctx ? ctx->launches : 0; }
int snfb_pin_host(void* p, size_t bytes) { return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : 1; }
int snfb_unpin_host(void* p) { return cudaHostUnregister(p) == cudaSuccess ? 0 : 1; }

}  // extern "C"
