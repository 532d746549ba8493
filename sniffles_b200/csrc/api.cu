// api.cu — C ABI of libsnfb200.so (include/snfb.h): context, memory, stage drivers.
// The product path has no CPU implementation: every stage below launches CUDA kernels.
//
// Run model.  Every buffer of the pipeline has a capacity kept in the context.  A run enqueues stage A -> B -> C on the
// context's stream WITHOUT host synchronisation: sizes live in device counters, kernels are launched for the capacities
// and bound their loops and stores by the device-side counts.  The host looks at the counters twice: once on the copy
// stream after stage B (while the consensus kernels run) to size the device -> host copies, and once at the end.  If a
// capacity was too small (first run on a context, or a block unlike the previous one) the run is repeated with
// capacities that fit — the counters always report what was needed.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <dlfcn.h>
#include <omp.h>

#include "common.cuh"
#include "prims.cuh"
#include "extract.cuh"
#include "cluster.cuh"
#include "consensus.cuh"
#include "poa.cuh"
#include "combine.cuh"
#include "ingest.cuh"

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cudaGetLastError(); return 1; }
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
        if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; cudaGetLastError(); return 1; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};
// one allocation carved into 256-byte aligned arrays: a measuring pass, then an assigning pass over the same list
struct Carver {
    uint8_t* base = nullptr; size_t off = 0;
    template <class T> T* take(size_t n) { const size_t o = off; off += (n * sizeof(T) + 255) & ~(size_t)255; return base ? reinterpret_cast<T*>(base + o) : nullptr; }
};

constexpr int MAX_TIMINGS = 64;

// ---- NCCL, resolved at run time (the library links no collective library: a process that never gathers needs none) ----
struct Id128 { char b[128]; };      // ncclUniqueId is passed by value: 128 bytes
struct NcclApi {
    void* h = nullptr; bool tried = false;
    int (*GetUniqueId)(void*) = nullptr; int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr; int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr; const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static bool nccl_load(std::string* err) {
    if (g_nccl.tried) { if (!g_nccl.AllGather && err) *err = "NCCL is not available in this process"; return g_nccl.AllGather != nullptr; }
    g_nccl.tried = true;
    // a process that already carries NCCL (torch) must use that copy: two NCCL instances do not share bootstrap state
    void* h = dlopen(nullptr, RTLD_NOW | RTLD_GLOBAL);
    if (!h || !dlsym(h, "ncclAllGather")) { h = nullptr; const char* names[] = { "libnccl.so.2", "libnccl.so" }; for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; } }
    if (!h) { if (err) *err = "libnccl.so.2 not found"; return false; }
    g_nccl.h = h;
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) { g_nccl.AllGather = nullptr; if (err) *err = "NCCL symbols missing"; return false; }
    return true;
}

struct Caps { unsigned long long lead = 0, cand = 0, cand_lead = 0, rn = 0, alt = 0, scr16 = 0, item = 0, tile = 0, req = 0, req16 = 0; };

struct snfb_ctx {
    int device = 0; cudaStream_t st = nullptr, st_copy = nullptr, st_side = nullptr; cudaEvent_t ev_b = nullptr, ev_mid = nullptr, ev_fork = nullptr, ev_join = nullptr; std::string err;
    snfb_config cfg{}; bool have_cfg = false;
    // records
    bool loaded = false, on_device = false, seq_on_demand = false; const uint8_t* h_seq = nullptr; uint32_t evt_min = 0;     // E-bit threshold of the loaded CIGAR16 arena
    uint64_t n_rec = 0, n_cigar = 0, n_var = 0, n_seq = 0; uint32_t n_task = 0, n_contig = 0, n_tr = 0, n_mask = 0;
    const snfb_rec* d_rec = nullptr; const uint16_t* d_cigar = nullptr; const uint8_t* d_var = nullptr; const uint8_t* d_seq = nullptr;
    DevBuf b_rec, b_cigar, b_var, b_seq, b_task, b_contig, b_tr, b_trp, b_mask, b_mask_off, b_mask_task;
    HostBuf h_c16, h_rec16;        // BAM32 host input converted to CIGAR16 before the upload
    DevBuf b_comp, b_raw, b_ing; HostBuf h_ing; uint64_t ing_sizes[8] = {0, 0, 0, 0, 0, 0, 0, 0}; bool from_bam = false;      // device BAM ingest: BGZF bytes, inflated stream, work arrays
    std::vector<snfb_task> tasks;
    // capacities and the three arenas carved by them
    Caps cap; bool force_no_cuts = false;
    DevBuf b_ctr, arena_r, arena_l, arena_c;            // counters; per-record arrays; per-lead arrays (stages A + B); stage C
    uint64_t arena_r_for = 0, arena_r_cigar = 0; Caps arena_l_for, arena_c_for; uint32_t arena_r_tasks = 0;
    // per-record (arena_r)
    uint32_t* pass_chunks; uint32_t* choff; uint32_t* rec_foff; uint32_t* rec_nf; uint2* cdesc; uint2* csum; uint32_t* flist; uint2* fcnt; unsigned long long chunk_cap = 0;      // the chunk table of the CIGAR walk (extract.cuh)
    int32_t* rec_pos; int32_t* rec_end; uint8_t* rec_flags; double* rec_nm; uint32_t* rec_nlead; uint32_t* rec_lead_off; uint32_t* sa_list; extract::RecScan* scanrec; extract::RecClip* clip; int32_t* rec_big;
    uint32_t* task_first; uint32_t* task_last; uint32_t* task_reads; unsigned long long* task_cov; int32_t* task_span; double* task_nm; double* nm_part; unsigned* nm_cnt; extract::Seg* sa_seg; uint32_t* scan_tmp_r;
    // per-lead (arena_l): stage A leads + the whole of stage B
    snfb_lead* leads; extract::Event* ev_buf; snfb_lead* sorted_leads; uint32_t* radix_hist;
    cluster::B B{};
    // stage C (arena_c)
    consensus::C Cc{}; consensus::SeqReq* seq_req; uint32_t* arena_off; uint8_t* seq_arena;
    HostBuf h_seq_req, h_seq_arena; uint64_t seq_h2d_bytes = 0;
    bool stage_a_done = false, stage_b_done = false, stage_c_done = false;
    // host staging
    HostBuf h_ctr_buf; DevCounters* h_mid = nullptr; DevCounters* h_fin = nullptr; uint32_t* h_work = nullptr;
    HostBuf h_leads, h_task_reads, h_task_nm, h_rec_nm, h_cand, h_cand_leads, h_rnames, h_rn_off, h_task_cov, h_task_cov_raw, h_alt, h_cov_bins;
    // gather
    void* comm = nullptr; int rank = 0, nranks = 1; DevBuf b_gsend, b_grecv; HostBuf h_gather; unsigned long long gather_cap = 0;
    // timings
    cudaEvent_t ev[MAX_TIMINGS + 1]; const char* ev_name[MAX_TIMINGS + 1]; uint64_t ev_bytes[MAX_TIMINGS + 1]; int n_ev = 0; int n_ev_load = 0; uint64_t launches = 0; uint64_t reruns = 0;
};

// the shortest I / D / S the configured path looks at: SV signatures of minsvlen_screen, indels above 10 for the NM correction (leadprov.py:198-224)
static uint32_t evt_need(const snfb_ctx* ctx) {
    if (!ctx->have_cfg) return SNFB_CIGAR16_EVT_MIN;
    int t = ctx->cfg.minsvlen_screen < 1 ? 1 : ctx->cfg.minsvlen_screen;
    if ((ctx->cfg.qc_nm_measure || ctx->cfg.phase) && t > 11) t = 11;
    return (uint32_t)(t < (int)SNFB_CIGAR16_EVT_MIN ? t : (int)SNFB_CIGAR16_EVT_MIN);
}
static void ctx_fail(snfb_ctx* ctx, const char* what, const char* msg) { ctx->err = std::string(what) + ": " + msg; }
static int fail(snfb_ctx* ctx, const std::string& m) { ctx->err = m; return 1; }

// timing marks: every call records an event on the ctx stream; a named mark opens an interval that the
// next mark (named or not) closes, so host-side gaps between stages are never attributed to a kernel
static void mark(snfb_ctx* ctx, const char* name, uint64_t bytes = 0) {
    if (ctx->n_ev >= MAX_TIMINGS) return;
    cudaEventRecord(ctx->ev[ctx->n_ev], ctx->st);
    ctx->ev_name[ctx->n_ev] = name; ctx->ev_bytes[ctx->n_ev] = bytes; ++ctx->n_ev;
}
#define LAUNCHED(ctx, k) ((ctx)->launches += (k))
static int grid_for(unsigned long long n, int threads) { unsigned long long g = (n + threads - 1) / threads; if (g < 1) g = 1; if (g > 148ull * 32) g = 148ull * 32; return (int)g; }
static int bits_for(uint32_t n) { int b = 0; while ((1ull << b) < n) ++b; return b; }

__global__ void k_gather_leads(const snfb_lead* __restrict__ leads, const uint32_t* __restrict__ sval, snfb_lead* __restrict__ out, const unsigned long long* n_ptr, unsigned long long cap) {
    const unsigned long long n = *n_ptr < cap ? *n_ptr : cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) out[i] = leads[sval[i]];
}
// mean coverage per bin of `binsize` bases over the whole contig (snf.py:248-267): sum over the bin's positions of the per-base
// depth / binsize, from the per-record (start, end) arrays: one thread per record adds its overlap with every bin it touches
__global__ void k_cov_bins(const int32_t* __restrict__ rec_pos, const int32_t* __restrict__ rec_end, const uint8_t* __restrict__ rec_flags, uint32_t lo, uint32_t hi,
                           int binsize, long long contig_len, long long nbins, unsigned long long* __restrict__ acc) {
    for (unsigned long long i = lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < hi; i += (unsigned long long)gridDim.x * blockDim.x) {
        if (!(rec_flags[i] & extract::RF_PASS)) continue;
        long long a = rec_pos[i], e = rec_end[i]; if (a < 0) a = 0; if (e > contig_len) e = contig_len;      // numpy slice clipping (leadprov.py:510)
        if (e <= a) continue;
        for (long long bb = a / binsize; bb <= (e - 1) / binsize && bb < nbins; ++bb) {
            const long long s0 = bb * binsize, s1 = s0 + binsize;
            const long long o0 = a > s0 ? a : s0, o1 = e < s1 ? e : s1;
            if (o1 > o0) atomicAdd(&acc[bb], (unsigned long long)(o1 - o0));
        }
    }
}


// ------------------------------------------------------------------------------------------------ all-gather of the candidate buffers
// slot of one rank: [GatherHdr][cand][alt][rnames][rn_off][cand_leads], every section 16-byte aligned
struct GatherHdr { unsigned long long n_cand, n_alt, n_rn, n_leads, need_bytes, overflow, pad0, pad1; };
__host__ __device__ inline void gather_offsets(const GatherHdr& h, unsigned long long off[6]) {
    auto al = [](unsigned long long x) { return (x + 15ull) & ~15ull; };
    off[0] = sizeof(GatherHdr); off[1] = al(off[0] + h.n_cand * sizeof(snfb_cand)); off[2] = al(off[1] + h.n_alt); off[3] = al(off[2] + 8 * h.n_rn);
    off[4] = al(off[3] + 4 * h.n_cand); off[5] = al(off[4] + h.n_leads * sizeof(snfb_lead));
}
__device__ inline void copy16(uint8_t* dst, const uint8_t* src, unsigned long long nbytes) {      // both 16-byte aligned, whole grid cooperates
    const unsigned long long n16 = (nbytes + 15) >> 4;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}
__global__ void k_gather_pack(const DevCounters* ctr, const snfb_cand* cand, const uint8_t* alt, const uint64_t* rnames, const uint32_t* rn_off, const snfb_lead* leads, int with_leads,
                              uint8_t* slot, unsigned long long cap) {
    GatherHdr h; h.n_cand = ctr->n_cand; h.n_alt = ctr->n_alt_bytes; h.n_rn = ctr->n_rnames; h.n_leads = with_leads ? ctr->n_cand_leads : 0ull; h.pad0 = h.pad1 = 0;
    unsigned long long off[6]; gather_offsets(h, off);
    h.need_bytes = off[5]; h.overflow = off[5] > cap ? 1ull : 0ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<GatherHdr*>(slot) = h;
    if (h.overflow) return;
    copy16(slot + off[0], reinterpret_cast<const uint8_t*>(cand), h.n_cand * sizeof(snfb_cand));
    copy16(slot + off[1], alt, h.n_alt);
    copy16(slot + off[2], reinterpret_cast<const uint8_t*>(rnames), 8 * h.n_rn);
    copy16(slot + off[3], reinterpret_cast<const uint8_t*>(rn_off), 4 * h.n_cand);
    if (with_leads) copy16(slot + off[4], reinterpret_cast<const uint8_t*>(leads), h.n_leads * sizeof(snfb_lead));
}
// gathered slots -> merged arrays with the offsets of every rank rebased; merged layout: [nranks headers][cand][alt][rnames][rn_off (+1)][leads]
__global__ void k_gather_merge(const uint8_t* recv, unsigned long long cap, int nranks, int with_leads, uint8_t* out, unsigned long long out_cap, unsigned long long* out_layout /* [8] */) {
    GatherHdr tot{}; bool ovf = false;
    for (int r = 0; r < nranks; ++r) { const GatherHdr* h = reinterpret_cast<const GatherHdr*>(recv + (size_t)r * cap); tot.n_cand += h->n_cand; tot.n_alt += h->n_alt; tot.n_rn += h->n_rn; tot.n_leads += h->n_leads; ovf = ovf || h->overflow; }
    auto al = [](unsigned long long x) { return (x + 255ull) & ~255ull; };
    const unsigned long long o_hdr = 0, o_cand = al((unsigned long long)nranks * sizeof(GatherHdr)), o_alt = al(o_cand + tot.n_cand * sizeof(snfb_cand)), o_rn = al(o_alt + tot.n_alt), o_ro = al(o_rn + 8 * tot.n_rn),
                             o_leads = al(o_ro + 4 * (tot.n_cand + 1)), o_end = al(o_leads + tot.n_leads * sizeof(snfb_lead));
    const bool fits = !ovf && o_end <= out_cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) { out_layout[0] = o_cand; out_layout[1] = o_alt; out_layout[2] = o_rn; out_layout[3] = o_ro; out_layout[4] = o_leads; out_layout[5] = o_end; out_layout[6] = fits ? 0 : 1; out_layout[7] = tot.n_cand; }
    if (!fits) return;
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x, nthr = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long b_cand = 0, b_alt = 0, b_rn = 0, b_leads = 0;
    for (int r = 0; r < nranks; ++r) {
        const uint8_t* slot = recv + (size_t)r * cap; const GatherHdr h = *reinterpret_cast<const GatherHdr*>(slot);
        unsigned long long off[6]; gather_offsets(h, off);
        if (tid == 0) reinterpret_cast<GatherHdr*>(out + o_hdr)[r] = h;
        const snfb_cand* sc = reinterpret_cast<const snfb_cand*>(slot + off[0]); snfb_cand* dc = reinterpret_cast<snfb_cand*>(out + o_cand) + b_cand;
        for (unsigned long long i = tid; i < h.n_cand; i += nthr) { snfb_cand c = sc[i]; if (c.alt_off >= 0) c.alt_off += (int)b_alt; if (with_leads) { c.lead_off += (int)b_leads; c.long_off += (int)b_leads; } dc[i] = c; }
        for (unsigned long long i = tid; i < h.n_alt; i += nthr) (out + o_alt + b_alt)[i] = (slot + off[1])[i];
        const uint64_t* sr = reinterpret_cast<const uint64_t*>(slot + off[2]); uint64_t* dr = reinterpret_cast<uint64_t*>(out + o_rn) + b_rn;
        for (unsigned long long i = tid; i < h.n_rn; i += nthr) dr[i] = sr[i];
        const uint32_t* so = reinterpret_cast<const uint32_t*>(slot + off[3]); uint32_t* dof = reinterpret_cast<uint32_t*>(out + o_ro) + b_cand;
        for (unsigned long long i = tid; i < h.n_cand; i += nthr) dof[i] = so[i] + (uint32_t)b_rn;
        if (with_leads) { const uint4* sl = reinterpret_cast<const uint4*>(slot + off[4]); uint4* dl = reinterpret_cast<uint4*>(out + o_leads) + 4 * b_leads; for (unsigned long long i = tid; i < 4 * h.n_leads; i += nthr) dl[i] = sl[i]; }
        b_cand += h.n_cand; b_alt += h.n_alt; b_rn += h.n_rn; b_leads += h.n_leads;
    }
    if (tid == 0) reinterpret_cast<uint32_t*>(out + o_ro)[tot.n_cand] = (uint32_t)tot.n_rn;
}

// lanes per BGZF block of the DEFLATE kernel: SNFB_INFLATE_LANES = 32, 16 or 8 (measured default below)
template <int NL> static void launch_inflate_nl(snfb_ctx* ctx, const ingest::BgzfBlock* d_blk, unsigned nb, ingest::IngestCounters* d_ctr) {
    static bool attr = false;
    const size_t smem = ingest::inflate_smem_bytes<NL>();
    if (!attr) { cudaFuncSetAttribute(ingest::k_inflate<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    const unsigned per_block = ingest::INF_WARPS * (32 / NL);
    const unsigned resident = (unsigned)std::max<size_t>(1, std::min<size_t>(8, (220 * 1024) / (smem + 1024)));
    const unsigned grid = std::min<unsigned>((nb + per_block - 1) / per_block, 148u * resident);
    ingest::k_inflate<NL><<<grid, ingest::INF_WARPS * 32, smem, ctx->st>>>(ctx->b_comp.as<uint8_t>(), d_blk, nb, ctx->b_raw.as<uint8_t>(), d_ctr);
}
static void launch_inflate(snfb_ctx* ctx, const ingest::BgzfBlock* d_blk, unsigned nb, ingest::IngestCounters* d_ctr) {
    static int lanes = 0;
    if (!lanes) { const char* e = getenv("SNFB_INFLATE_LANES"); lanes = e ? atoi(e) : 16; if (lanes != 32 && lanes != 16 && lanes != 8) lanes = 16; }
    if (lanes == 32) launch_inflate_nl<32>(ctx, d_blk, nb, d_ctr); else if (lanes == 16) launch_inflate_nl<16>(ctx, d_blk, nb, d_ctr); else launch_inflate_nl<8>(ctx, d_blk, nb, d_ctr);
}

extern "C" {

int snfb_version(void) { return SNFB_ABI_VERSION; }
size_t snfb_sizeof(int which) {
    switch (which) { case 0: return sizeof(snfb_rec); case 1: return sizeof(snfb_task); case 2: return sizeof(snfb_contig); case 3: return sizeof(snfb_records);
                     case 4: return sizeof(snfb_config); case 5: return sizeof(snfb_lead); case 6: return sizeof(snfb_cand); case 7: return sizeof(snfb_gather_view); default: return 0; }
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
    if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->st_copy, cudaStreamNonBlocking) != cudaSuccess
        || cudaStreamCreateWithFlags(&ctx->st_side, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return 4; }
    cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_mid, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    for (int i = 0; i <= MAX_TIMINGS; ++i) cudaEventCreate(&ctx->ev[i]);
    if (ctx->h_ctr_buf.ensure(2 * sizeof(DevCounters) + 256) || ctx->b_ctr.ensure(sizeof(DevCounters) + 64)) { delete ctx; return 5; }
    ctx->h_mid = ctx->h_ctr_buf.as<DevCounters>(); ctx->h_fin = ctx->h_mid + 1; ctx->h_work = reinterpret_cast<uint32_t*>(ctx->h_fin + 1);
    memset(ctx->h_ctr_buf.p, 0, 2 * sizeof(DevCounters) + 256);
    cudaFuncSetAttribute(cluster::k_cluster_warp<cluster::SMALL_CAP, cluster::CWS_WARPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster::CwCfg<cluster::SMALL_CAP, cluster::CWS_WARPS>::smem);
    cudaFuncSetAttribute(cluster::k_cluster_warp<cluster::WARP_CAP, cluster::CWM_WARPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster::CwCfg<cluster::WARP_CAP, cluster::CWM_WARPS>::smem);
    cudaFuncSetAttribute(cluster::k_cluster_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cluster::CB_SMEM);
    *out = ctx; return 0;
}

void snfb_ctx_destroy(snfb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st); cudaStreamSynchronize(ctx->st_copy); cudaStreamSynchronize(ctx->st_side);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    DevBuf* bufs[] = { &ctx->b_rec, &ctx->b_cigar, &ctx->b_var, &ctx->b_seq, &ctx->b_task, &ctx->b_contig, &ctx->b_tr, &ctx->b_trp, &ctx->b_mask, &ctx->b_mask_off, &ctx->b_mask_task,
                       &ctx->b_ctr, &ctx->arena_r, &ctx->arena_l, &ctx->arena_c, &ctx->b_gsend, &ctx->b_grecv, &ctx->b_comp, &ctx->b_raw, &ctx->b_ing };
    for (DevBuf* b : bufs) b->release();
    HostBuf* hb[] = { &ctx->h_c16, &ctx->h_rec16, &ctx->h_seq_req, &ctx->h_seq_arena, &ctx->h_ctr_buf, &ctx->h_leads, &ctx->h_task_reads, &ctx->h_task_nm, &ctx->h_rec_nm, &ctx->h_cand, &ctx->h_cand_leads,
                      &ctx->h_rnames, &ctx->h_rn_off, &ctx->h_task_cov, &ctx->h_task_cov_raw, &ctx->h_alt, &ctx->h_cov_bins, &ctx->h_gather, &ctx->h_ing };
    for (HostBuf* b : hb) b->release();
    for (int i = 0; i <= MAX_TIMINGS; ++i) cudaEventDestroy(ctx->ev[i]);
    cudaEventDestroy(ctx->ev_b); cudaEventDestroy(ctx->ev_mid); cudaEventDestroy(ctx->ev_fork); cudaEventDestroy(ctx->ev_join);
    cudaStreamDestroy(ctx->st_side); cudaStreamDestroy(ctx->st_copy); cudaStreamDestroy(ctx->st);
    delete ctx;
}

const char* snfb_last_error(snfb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int snfb_set_config(snfb_ctx* ctx, const snfb_config* cfg) {
    if (!ctx || !cfg) return 1;
    ctx->force_no_cuts = false;      // a new configuration decides again where chains of bins may be cut
    if (cfg->consensus_kmer_len != 6) return fail(ctx, "consensus_kmer_len must be 6 (the reference fixes it, config.py:550)");
    if (cfg->cluster_binsize <= 0 || cfg->cluster_resplit_binsize <= 0 || cfg->coverage_binsize <= 0) return fail(ctx, "bin sizes must be positive");
    ctx->cfg = *cfg; ctx->have_cfg = true; ctx->stage_a_done = ctx->stage_b_done = ctx->stage_c_done = false; return 0;
}

// ---- BAM CIGAR words -> CIGAR16 (include/snfb.h).  Host code; the only place the 32-bit form is read. ----
static inline int c16_group_words(uint32_t len) { return len < (1u << 11) ? 1 : (len < (1u << 23) ? 2 : 3); }
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
static inline uint64_t c16_write(const uint32_t* cg, uint32_t n, uint16_t* out, uint32_t evt_min) {      // returns the words written (= c16_count)
    uint64_t k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t len = cg[i] >> 4; const int g = c16_group_words(len); const unsigned cls = C16_CLASS[cg[i] & 15u];
        if ((k & 7) + g > 8) { while (k & 7) out[k++] = 0; }
        const unsigned e = ((cls == 1 || cls == 2 || cls == 5) && len >= evt_min) ? 0x4000u : 0u;      // E: an I / D / S the streaming kernel has to look at
        out[k++] = (uint16_t)(e | (cls << 11) | (len & 0x7ffu));
        if (g >= 2) out[k++] = (uint16_t)(0x8000u | (1u << 12) | ((len >> 11) & 0xfffu));
        if (g >= 3) out[k++] = (uint16_t)(0x8000u | (2u << 12) | ((len >> 23) & 0xfffu));
    }
    return k;
}
// pass 1: off[i] = first word of record i in the 16-bit arena (every record starts a 16-byte group); false when an op code is unknown
static bool c16_offsets(const snfb_rec* rec_in, uint64_t n_rec, const uint32_t* cigar32, std::vector<uint64_t>& off) {
    off.assign(n_rec + 1, 0);
    bool bad = false;
    #pragma omp parallel for schedule(static) reduction(|| : bad)
    for (long long i = 0; i < (long long)n_rec; ++i) { bool b = false; const uint64_t w = c16_count(cigar32 + rec_in[i].cigar_off, rec_in[i].n_cigar, &b); bad = bad || b; off[i + 1] = (w + 7) & ~7ull; }
    if (bad) return false;
    for (uint64_t i = 0; i < n_rec; ++i) off[i + 1] += off[i];
    return true;
}
// pass 2: the words and the rewritten records
static void c16_fill(const snfb_rec* rec_in, uint64_t n_rec, const uint32_t* cigar32, const std::vector<uint64_t>& off, snfb_rec* rec_out, uint16_t* out16, uint32_t evt_min) {
    #pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n_rec; ++i) {
        uint16_t* dst = out16 + off[i]; const uint64_t span = off[i + 1] - off[i];
        const uint64_t k = c16_write(cigar32 + rec_in[i].cigar_off, rec_in[i].n_cigar, dst, evt_min);
        if (k < span) memset(dst + k, 0, 2 * (span - k));
        snfb_rec r = rec_in[i]; r.n_cigar = (uint32_t)k; r.cigar_off = off[i];
        rec_out[i] = r;
    }
    memset(out16 + off[n_rec], 0, 16);                            // one zero group of slack after the last record
}
uint64_t snfb_pack_cigar16(const snfb_rec* rec_in, uint64_t n_rec, const uint32_t* cigar32, snfb_rec* rec_out, uint16_t* out16, uint64_t out_cap, uint32_t evt_min) {
    if (n_rec && (!rec_in || !cigar32)) return UINT64_MAX;
    if (evt_min == 0) evt_min = SNFB_CIGAR16_EVT_MIN;
    std::vector<uint64_t> off;
    if (!c16_offsets(rec_in, n_rec, cigar32, off)) return UINT64_MAX;
    const uint64_t total = off[n_rec] + 8;
    if (!out16) return total;
    if (!rec_out || out_cap < total) return UINT64_MAX;
    c16_fill(rec_in, n_rec, cigar32, off, rec_out, out16, evt_min);
    return total;
}

// thread per record: every offset of the block must stay inside its arena / table (a malformed block fails instead of reading out of bounds)
__global__ void k_validate(const snfb_rec* __restrict__ rec, uint32_t n_rec, uint32_t n_task, uint64_t n_cigar, uint64_t n_var, uint64_t n_seq, int check_seq, DevCounters* ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n_rec) return;
    const snfb_rec r = rec[i];
    bool bad = r.task < 0 || (uint32_t)r.task >= n_task || (r.cigar_off & 7) || r.cigar_off + (uint64_t)r.n_cigar > n_cigar || r.var_off + (uint64_t)r.l_qname + r.sa_len > n_var || r.l_seq < 0;
    if (check_seq && r.l_seq >= 0 && r.seq_off + (uint64_t)((r.l_seq + 1) / 2) > n_seq) bad = true;
    if (bad) atomicAdd(&ctr->bad_records, 1ULL);
}

// host-side checks of the small tables of a block
static int check_tables(snfb_ctx* ctx, const snfb_records* R) {
    for (uint32_t t = 0; t < R->n_task; ++t) {
        const snfb_task& k = R->task[t];
        if (k.tr_n < 0 || k.tr_off < 0 || (uint64_t)k.tr_off + (uint64_t)k.tr_n > R->n_tr) return fail(ctx, "task table: tandem-repeat range outside tr[]");
        if (R->n_contig && (k.contig < 0 || (uint32_t)k.contig >= R->n_contig)) return fail(ctx, "task table: contig index out of range");
        if (k.contig_len < 0 || k.start > k.end) return fail(ctx, "task table: bad region");
        if ((long long)k.contig_len >= ((long long)1 << 26) * (long long)(ctx->have_cfg ? ctx->cfg.cluster_binsize : 100)) return fail(ctx, "contig too long for the bin field of the sort key (contig_len / cluster_binsize must stay below 2^26)");
    }
    if (R->n_mask && R->mask && R->mask_task_off) {
        for (uint32_t t = 0; t < R->n_task; ++t) if (R->mask_task_off[t] > R->mask_task_off[t + 1] || R->mask_task_off[t + 1] > R->n_mask) return fail(ctx, "mask_task_off must be non-decreasing and end at n_mask");
    }
    return 0;
}
// task / contig / tandem-repeat / N-mask tables -> device
static int upload_tables(snfb_ctx* ctx, const snfb_records* R) {
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
    return 0;
}

int snfb_load_records(snfb_ctx* ctx, const snfb_records* R) {
    if (!ctx || !R) return 1;
    cudaSetDevice(ctx->device);
    ctx->loaded = false; ctx->stage_a_done = ctx->stage_b_done = ctx->stage_c_done = false; ctx->force_no_cuts = false;
    if (R->n_task == 0 || R->n_task > 65535) return fail(ctx, "n_task must be in 1..65535");
    if (R->n_rec >= (1ull << 28)) return fail(ctx, "too many records in one block (2^28)");
    if (!R->task || (R->n_rec && (!R->rec || !R->cigar))) return fail(ctx, "null table in the record block");
    if (check_tables(ctx, R)) return 1;
    ctx->n_rec = R->n_rec; ctx->n_cigar = R->n_cigar; ctx->n_var = R->n_var; ctx->n_seq = R->n_seq;
    ctx->n_task = R->n_task; ctx->n_contig = R->n_contig; ctx->n_tr = R->n_tr; ctx->on_device = R->on_device == SNFB_MEM_DEVICE; ctx->seq_on_demand = R->on_device == SNFB_MEM_HOST_SEQ_ON_DEMAND; ctx->h_seq = ctx->seq_on_demand ? R->seq : nullptr;
    ctx->n_ev = 0;
    const snfb_rec* src_rec = R->rec; const uint16_t* src_cigar = reinterpret_cast<const uint16_t*>(R->cigar); uint64_t n_words = R->n_cigar;
    if (R->cigar_fmt == SNFB_CIGAR_BAM32) {
        if (R->on_device == SNFB_MEM_DEVICE) return fail(ctx, "device-resident records must carry CIGAR16 (convert with snfb_pack_cigar16)");
        for (uint64_t i = 0; i < R->n_rec; ++i) if (R->rec[i].cigar_off + (uint64_t)R->rec[i].n_cigar > R->n_cigar) return fail(ctx, "a record's CIGAR lies outside the cigar arena");
        // host conversion: the kernels only read CIGAR16 (two passes over the BAM words: offsets, then the words)
        std::vector<uint64_t> off;
        if (!c16_offsets(R->rec, R->n_rec, reinterpret_cast<const uint32_t*>(R->cigar), off)) return fail(ctx, "a CIGAR holds an operation the path does not know");
        const uint64_t need = off[R->n_rec] + 8;
        if (ctx->h_c16.ensure(2 * need + 16) || ctx->h_rec16.ensure(sizeof(snfb_rec) * (R->n_rec + 1))) return fail(ctx, "out of pinned memory for the CIGAR16 conversion");
        c16_fill(R->rec, R->n_rec, reinterpret_cast<const uint32_t*>(R->cigar), off, ctx->h_rec16.as<snfb_rec>(), ctx->h_c16.as<uint16_t>(), evt_need(ctx));
        src_rec = ctx->h_rec16.as<snfb_rec>(); src_cigar = ctx->h_c16.as<uint16_t>(); n_words = need;
        ctx->evt_min = evt_need(ctx);
    } else if (R->cigar_fmt != SNFB_CIGAR_16) return fail(ctx, "unknown cigar_fmt");
    else ctx->evt_min = R->cigar_evt_min ? R->cigar_evt_min : SNFB_CIGAR16_EVT_MIN;
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
    if (upload_tables(ctx, R)) return 1;
    mark(ctx, nullptr);
    ctx->n_ev_load = ctx->n_ev;
    ctx->loaded = true; return 0;
}


// ---- device BAM ingest ----
// BGZF block headers of a buffer of whole blocks (SAM spec §4.1): payload offset / length, inflated size, start of every block
static int walk_bgzf(snfb_ctx* ctx, const uint8_t* z, uint64_t n, std::vector<ingest::BgzfBlock>& blocks, std::vector<uint64_t>& cstart, uint64_t* raw_len) {
    uint64_t o = 0, uo = 0;
    while (o < n) {
        if (o + 18 > n || z[o] != 0x1f || z[o + 1] != 0x8b || z[o + 2] != 8 || !(z[o + 3] & 4)) return fail(ctx, "not a BGZF block (gzip member with an extra field expected)");
        const uint32_t xlen = z[o + 10] | (z[o + 11] << 8);
        if (o + 12 + xlen > n) return fail(ctx, "truncated BGZF header");
        uint32_t bsize = 0; uint64_t e = o + 12; const uint64_t xend = o + 12 + xlen;
        while (e + 4 <= xend) { const uint32_t slen = z[e + 2] | (z[e + 3] << 8); if (z[e] == 66 && z[e + 1] == 67 && slen == 2 && e + 6 <= xend) bsize = (z[e + 4] | (z[e + 5] << 8)) + 1u; e += 4 + slen; }
        if (!bsize || bsize < 12 + xlen + 8 || o + bsize > n) return fail(ctx, "BGZF block without a BC field or truncated");
        ingest::BgzfBlock b; b.in_off = o + 12 + xlen; b.in_len = bsize - 12 - xlen - 8;
        memcpy(&b.isize, z + o + bsize - 4, 4); b.out_off = uo;
        if (b.isize > 65536u) return fail(ctx, "BGZF block claims more than 64 KiB of data");
        blocks.push_back(b); cstart.push_back(o);
        uo += b.isize; o += bsize;
    }
    *raw_len = uo; return 0;
}
static int inflate_to_device(snfb_ctx* ctx, const uint8_t* z, uint64_t n, std::vector<uint64_t>& cstart, std::vector<ingest::BgzfBlock>& blocks, uint64_t* raw_len) {
    if (walk_bgzf(ctx, z, n, blocks, cstart, raw_len)) return 1;
    if (*raw_len >= (1ull << 36)) return fail(ctx, "more than 64 GiB of inflated BAM in one ingest call: split the task list");
    if (ctx->b_comp.ensure(n + 64) || ctx->b_raw.ensure(*raw_len + 64)) return fail(ctx, "out of device memory for the BGZF bytes / the inflated stream");
    mark(ctx, "h2d_bgzf", n);
    CUDA_TRY(cudaMemcpyAsync(ctx->b_comp.p, z, n, cudaMemcpyHostToDevice, ctx->st));
    CUDA_TRY(cudaMemsetAsync(ctx->b_comp.as<uint8_t>() + n, 0, 64, ctx->st));
    CUDA_TRY(cudaMemsetAsync(ctx->b_raw.as<uint8_t>() + *raw_len, 0, 64, ctx->st));
    return 0;
}

int snfb_inflate_bgzf(snfb_ctx* ctx, const uint8_t* bgzf, uint64_t n_bytes, uint8_t* out, uint64_t out_cap, uint64_t* out_len) {
    if (!ctx || !bgzf || !out_len) return 1;
    cudaSetDevice(ctx->device);
    std::vector<ingest::BgzfBlock> blocks; std::vector<uint64_t> cstart; uint64_t raw_len = 0;
    ctx->n_ev = 0;
    if (inflate_to_device(ctx, bgzf, n_bytes, cstart, blocks, &raw_len)) return 1;
    *out_len = raw_len;
    const size_t nb = blocks.size();
    if (ctx->b_ing.ensure(256 + sizeof(ingest::BgzfBlock) * (nb + 1))) return fail(ctx, "out of device memory (ingest tables)");
    ingest::IngestCounters* d_ctr = ctx->b_ing.as<ingest::IngestCounters>(); ingest::BgzfBlock* d_blk = reinterpret_cast<ingest::BgzfBlock*>(ctx->b_ing.as<uint8_t>() + 256);
    CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(ingest::IngestCounters), ctx->st));
    CUDA_TRY(cudaMemcpyAsync(d_blk, blocks.data(), sizeof(ingest::BgzfBlock) * nb, cudaMemcpyHostToDevice, ctx->st));
    mark(ctx, "inflate", n_bytes + raw_len);
    if (nb) { launch_inflate(ctx, d_blk, (unsigned)nb, d_ctr); LAUNCHED(ctx, 1); }
    mark(ctx, nullptr);
    ingest::IngestCounters hc;
    CUDA_TRY(cudaMemcpyAsync(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    if (hc.bad_blocks) return fail(ctx, "inflate: " + std::to_string(hc.bad_blocks) + " BGZF block(s) failed to decode (first: block " + std::to_string(hc.first_bad_block) + ", code " + std::to_string(hc.first_bad_code) + ")");
    if (out) {
        if (out_cap < raw_len) return fail(ctx, "snfb_inflate_bgzf: output buffer too small");
        CUDA_TRY(cudaMemcpy(out, ctx->b_raw.p, raw_len, cudaMemcpyDeviceToHost));
    }
    return 0;
}

int snfb_load_bam(snfb_ctx* ctx, const snfb_bam_input* in) {
    if (!ctx || !in) return 1;
    cudaSetDevice(ctx->device);
    ctx->loaded = false; ctx->from_bam = false; ctx->stage_a_done = ctx->stage_b_done = ctx->stage_c_done = false; ctx->force_no_cuts = false;
    if (in->n_task == 0 || in->n_task > 65535) return fail(ctx, "n_task must be in 1..65535");
    if (!in->task || (in->n_bytes && !in->bgzf) || (in->n_span && !in->span)) return fail(ctx, "null table in the BAM input");
    snfb_records T; memset(&T, 0, sizeof(T));
    T.n_task = in->n_task; T.n_contig = in->n_contig; T.n_tr = in->n_tr; T.n_mask = in->n_mask; T.task = in->task; T.contig = in->contig; T.tr = in->tr; T.mask = in->mask; T.mask_task_off = in->mask_task_off;
    if (check_tables(ctx, &T)) return 1;
    ctx->n_ev = 0;
    std::vector<ingest::BgzfBlock> blocks; std::vector<uint64_t> cstart; uint64_t raw_len = 0;
    if (inflate_to_device(ctx, in->bgzf, in->n_bytes, cstart, blocks, &raw_len)) return 1;
    const size_t nb = blocks.size(); const uint64_t ns = in->n_span;
    // spans: virtual offsets -> offsets in the inflated stream
    std::vector<ingest::Span> spans(ns);
    for (uint64_t i = 0; i < ns; ++i) {
        const snfb_bam_span& sp = in->span[i];
        if (sp.task >= in->n_task) return fail(ctx, "span: task index out of range");
        auto resolve = [&](uint64_t c, uint32_t u, uint64_t* out) -> bool {
            if (c == in->n_bytes) { *out = raw_len; return u == 0; }
            auto it = std::lower_bound(cstart.begin(), cstart.end(), c);
            if (it == cstart.end() || *it != c) return false;
            const ingest::BgzfBlock& b = blocks[(size_t)(it - cstart.begin())];
            if (u > b.isize) return false;
            *out = b.out_off + u; return true;
        };
        uint64_t ub = 0, ue = 0;
        if (!resolve(sp.cbeg, sp.ubeg, &ub) || !resolve(sp.cend, sp.uend, &ue) || ue < ub) return fail(ctx, "span: a virtual offset does not name a BGZF block of the buffer");
        if (i && spans[i - 1].task > sp.task) return fail(ctx, "spans must be listed task by task");
        if (i && spans[i - 1].task == sp.task && spans[i - 1].uend > ub) return fail(ctx, "spans of a task must be in file order and must not overlap");
        spans[i].ubeg = ub; spans[i].uend = ue; spans[i].task = sp.task; spans[i]._pad = 0;
    }
    if (upload_tables(ctx, &T)) return 1;
    // fixed part of the work area
    Carver m0; ingest::IngestCounters* d_ctr = nullptr; ingest::BgzfBlock* d_blk = nullptr; ingest::Span* d_span = nullptr; uint32_t* span_cnt = nullptr; uint32_t* span_base = nullptr; uint32_t* scan_tmp0 = nullptr;
    auto carve0 = [&](Carver& c) { d_ctr = c.take<ingest::IngestCounters>(1); d_blk = c.take<ingest::BgzfBlock>(nb + 1); d_span = c.take<ingest::Span>(ns + 1); span_cnt = c.take<uint32_t>(ns + 1); span_base = c.take<uint32_t>(ns + 1); scan_tmp0 = c.take<uint32_t>(prims::scan_tmp_elems(ns + 1) + 16); };
    carve0(m0);
    if (ctx->b_ing.ensure(m0.off + 256)) return fail(ctx, "out of device memory (ingest tables)");
    { Carver a; a.base = ctx->b_ing.as<uint8_t>(); carve0(a); }
    cudaStream_t st = ctx->st; const uint8_t* raw = ctx->b_raw.as<uint8_t>();
    CUDA_TRY(cudaMemsetAsync(d_ctr, 0, sizeof(ingest::IngestCounters), st));
    CUDA_TRY(cudaMemcpyAsync(d_blk, blocks.data(), sizeof(ingest::BgzfBlock) * nb, cudaMemcpyHostToDevice, st));
    if (ns) CUDA_TRY(cudaMemcpyAsync(d_span, spans.data(), sizeof(ingest::Span) * ns, cudaMemcpyHostToDevice, st));
    mark(ctx, "inflate", in->n_bytes + raw_len);
    if (nb) { launch_inflate(ctx, d_blk, (unsigned)nb, d_ctr); LAUNCHED(ctx, 1); }
    mark(ctx, "walk_records", 0);
    if (ns) {
        ingest::k_walk<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(raw, raw_len, d_span, (unsigned)ns, 0, span_cnt, nullptr, nullptr, 0, d_ctr); LAUNCHED(ctx, 1);
        LAUNCHED(ctx, prims::exclusive_scan(span_cnt, span_base, scan_tmp0, nullptr, ns, &d_ctr->n_raw, st));
    }
    if (ctx->h_ing.ensure(2 * sizeof(ingest::IngestCounters))) return fail(ctx, "out of pinned memory");
    ingest::IngestCounters* hc = ctx->h_ing.as<ingest::IngestCounters>();
    CUDA_TRY(cudaMemcpyAsync(hc, d_ctr, sizeof(*hc), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (hc->bad_blocks) return fail(ctx, "inflate: " + std::to_string(hc->bad_blocks) + " BGZF block(s) failed to decode (first: block " + std::to_string(hc->first_bad_block) + ", code " + std::to_string(hc->first_bad_code) + ")");
    if (hc->bad_chain) return fail(ctx, "BAM record chain broken in " + std::to_string(hc->bad_chain) + " span(s): a span does not start or end on a record boundary, or the data is truncated");
    const uint64_t n_raw = hc->n_raw;
    if (n_raw >= (1ull << 28)) return fail(ctx, "too many records in one block (2^28)");
    // per-raw-record work arrays live behind the fixed part
    ingest::RawRec* recs = nullptr; uint32_t *keep = nullptr, *idx = nullptr, *groups = nullptr, *grp_off = nullptr, *var16 = nullptr, *var_off = nullptr, *seq16 = nullptr, *seq_off = nullptr, *scan_tmp = nullptr;
    auto carve1 = [&](Carver& c) { carve0(c); recs = c.take<ingest::RawRec>(n_raw + 1); keep = c.take<uint32_t>(n_raw + 1); idx = c.take<uint32_t>(n_raw + 1); groups = c.take<uint32_t>(n_raw + 1); grp_off = c.take<uint32_t>(n_raw + 1);
                                   var16 = c.take<uint32_t>(n_raw + 1); var_off = c.take<uint32_t>(n_raw + 1); seq16 = c.take<uint32_t>(n_raw + 1); seq_off = c.take<uint32_t>(n_raw + 1); scan_tmp = c.take<uint32_t>(prims::scan_tmp_elems(n_raw + 1) + 16); };
    Carver m1; carve1(m1);
    if (m1.off + 256 > ctx->b_ing.cap) {
        // grow without losing the fixed part: a new buffer, the fixed part copied over
        DevBuf nbuf; if (nbuf.ensure(m1.off + 256)) return fail(ctx, "out of device memory (ingest work arrays)");
        CUDA_TRY(cudaMemcpyAsync(nbuf.p, ctx->b_ing.p, m0.off, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        ctx->b_ing.release(); ctx->b_ing = nbuf;
    }
    { Carver a; a.base = ctx->b_ing.as<uint8_t>(); carve1(a); }
    const uint32_t evt = evt_need(ctx);
    uint64_t n_rec = 0, n_groups = 0, n_var16 = 0, n_seq16 = 0;
    if (n_raw) {
        const unsigned warp_grid = (unsigned)std::min<uint64_t>((n_raw + 7) / 8, 148ull * 16);
        ingest::k_walk<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(raw, raw_len, d_span, (unsigned)ns, 1, span_cnt, span_base, recs, n_raw, d_ctr);
        mark(ctx, "parse_records", 0);
        ingest::k_parse<<<(unsigned)((n_raw + 127) / 128), 128, 0, st>>>(raw, recs, (unsigned)n_raw, ctx->b_task.as<snfb_task>(), d_ctr);
        mark(ctx, "record_sizes", 0);
        ingest::k_rec_sizes<<<warp_grid, 256, 0, st>>>(raw, recs, (unsigned)n_raw, ctx->b_task.as<snfb_task>(), evt, keep, groups, var16, seq16, d_ctr); LAUNCHED(ctx, 3);
        LAUNCHED(ctx, prims::exclusive_scan(keep, idx, scan_tmp, nullptr, n_raw, &d_ctr->n_keep, st));
        LAUNCHED(ctx, prims::exclusive_scan(groups, grp_off, scan_tmp, nullptr, n_raw, &d_ctr->n_groups, st));
        LAUNCHED(ctx, prims::exclusive_scan(var16, var_off, scan_tmp, nullptr, n_raw, &d_ctr->n_var, st));
        LAUNCHED(ctx, prims::exclusive_scan(seq16, seq_off, scan_tmp, nullptr, n_raw, &d_ctr->n_seq16, st));
        CUDA_TRY(cudaMemcpyAsync(hc, d_ctr, sizeof(*hc), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (hc->malformed || hc->bad_cigar) return fail(ctx, std::to_string(hc->malformed) + " malformed BAM record(s), " + std::to_string(hc->bad_cigar) + " with a CIGAR operation the path does not know");
        n_rec = hc->n_keep; n_groups = hc->n_groups; n_var16 = hc->n_var; n_seq16 = hc->n_seq16;
    }
    const uint64_t n_words = 8 * n_groups + 8;               // one zero group of slack after the last record (as snfb_pack_cigar16)
    if (ctx->b_rec.ensure(sizeof(snfb_rec) * (n_rec + 1)) || ctx->b_cigar.ensure(2 * (n_words + 16)) || ctx->b_var.ensure(16 * n_var16 + 16) || ctx->b_seq.ensure(16 * n_seq16 + 16))
        return fail(ctx, "out of device memory for the record block");
    mark(ctx, "pack_records", sizeof(snfb_rec) * n_rec + 2 * n_words + 16 * n_var16 + 16 * n_seq16);
    CUDA_TRY(cudaMemsetAsync(ctx->b_cigar.as<uint16_t>() + 8 * n_groups, 0, 2 * 24, st));
    if (n_raw) {
        const unsigned warp_grid = (unsigned)std::min<uint64_t>((n_raw + 7) / 8, 148ull * 16);
        ingest::k_pack<<<warp_grid, 256, 0, st>>>(raw, recs, (unsigned)n_raw, evt, keep, idx, grp_off, groups, var_off, seq_off, ctx->b_rec.as<snfb_rec>(), ctx->b_cigar.as<uint16_t>(), ctx->b_var.as<uint8_t>(), ctx->b_seq.as<uint8_t>()); LAUNCHED(ctx, 1);
    }
    mark(ctx, nullptr);
    ctx->n_ev_load = ctx->n_ev;
    ctx->n_rec = n_rec; ctx->n_cigar = n_words; ctx->n_var = 16 * n_var16; ctx->n_seq = 16 * n_seq16; ctx->n_task = in->n_task; ctx->n_contig = in->n_contig; ctx->n_tr = in->n_tr;
    ctx->on_device = false; ctx->seq_on_demand = false; ctx->h_seq = nullptr; ctx->evt_min = evt;
    ctx->d_rec = ctx->b_rec.as<snfb_rec>(); ctx->d_cigar = ctx->b_cigar.as<uint16_t>(); ctx->d_var = ctx->b_var.as<uint8_t>(); ctx->d_seq = ctx->b_seq.as<uint8_t>();
    ctx->ing_sizes[0] = n_rec; ctx->ing_sizes[1] = n_words; ctx->ing_sizes[2] = 16 * n_var16; ctx->ing_sizes[3] = 16 * n_seq16; ctx->ing_sizes[4] = n_raw; ctx->ing_sizes[5] = nb; ctx->ing_sizes[6] = raw_len; ctx->ing_sizes[7] = in->n_bytes;
    CUDA_TRY(cudaStreamSynchronize(st));
    ctx->loaded = true; ctx->from_bam = true; return 0;
}

int snfb_ingest_sizes(snfb_ctx* ctx, uint64_t out[8]) {
    if (!ctx || !out || !ctx->from_bam) return 1;
    for (int i = 0; i < 8; ++i) out[i] = ctx->ing_sizes[i];
    return 0;
}
int snfb_ingest_fetch(snfb_ctx* ctx, snfb_rec* rec, uint16_t* cigar16, uint8_t* var, uint8_t* seq) {
    if (!ctx || !ctx->from_bam || !ctx->loaded) return ctx ? fail(ctx, "snfb_ingest_fetch: no block built by snfb_load_bam") : 1;
    cudaSetDevice(ctx->device);
    if (rec && ctx->n_rec) CUDA_TRY(cudaMemcpy(rec, ctx->d_rec, sizeof(snfb_rec) * ctx->n_rec, cudaMemcpyDeviceToHost));
    if (cigar16 && ctx->n_cigar) CUDA_TRY(cudaMemcpy(cigar16, ctx->d_cigar, 2 * ctx->n_cigar, cudaMemcpyDeviceToHost));
    if (var && ctx->n_var) CUDA_TRY(cudaMemcpy(var, ctx->d_var, ctx->n_var, cudaMemcpyDeviceToHost));
    if (seq && ctx->n_seq) CUDA_TRY(cudaMemcpy(seq, ctx->d_seq, ctx->n_seq, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ arenas
static void carve_r(snfb_ctx* ctx, Carver& c) {
    const size_t n = ctx->n_rec + 1, nt = ctx->n_task;
    ctx->rec_pos = c.take<int32_t>(n); ctx->rec_end = c.take<int32_t>(n); ctx->rec_flags = c.take<uint8_t>(n); ctx->rec_nm = c.take<double>(n); ctx->rec_nlead = c.take<uint32_t>(n); ctx->rec_lead_off = c.take<uint32_t>(n);
    ctx->pass_chunks = c.take<uint32_t>(n); ctx->choff = c.take<uint32_t>(n); ctx->rec_foff = c.take<uint32_t>(n); ctx->rec_nf = c.take<uint32_t>(n);
    // a record of g groups has ceil(g / CH) chunks: at most g / CH + 1
    ctx->chunk_cap = ctx->n_cigar / (8ull * extract::CH) + ctx->n_rec + 1; { const size_t nc = (size_t)ctx->chunk_cap + 1; ctx->cdesc = c.take<uint2>(nc); ctx->csum = c.take<uint2>(nc); ctx->flist = c.take<uint32_t>(nc); ctx->fcnt = c.take<uint2>(nc); }
    ctx->sa_list = c.take<uint32_t>(n); ctx->scanrec = c.take<extract::RecScan>(n); ctx->clip = c.take<extract::RecClip>(n); ctx->rec_big = c.take<int32_t>(n);
    ctx->task_first = c.take<uint32_t>(nt); ctx->task_last = c.take<uint32_t>(nt); ctx->task_reads = c.take<uint32_t>(nt); ctx->task_cov = c.take<unsigned long long>(nt); ctx->task_span = c.take<int32_t>(nt); ctx->task_nm = c.take<double>(nt);
    const size_t cpt = (ctx->n_rec + extract::NM_CHUNK - 1) / extract::NM_CHUNK + 1;
    ctx->nm_part = c.take<double>(cpt * nt + 1); ctx->nm_cnt = c.take<unsigned>(cpt * nt + 1);
    ctx->sa_seg = c.take<extract::Seg>((size_t)extract::MAXSEG * extract::SA_THREADS * extract::SA_BLOCKS);
    ctx->scan_tmp_r = c.take<uint32_t>(prims::scan_tmp_elems(n) + 16);
}
static void carve_l(snfb_ctx* ctx, Carver& c) {
    const size_t n = (size_t)ctx->cap.lead + 8; cluster::B& b = ctx->B;
    ctx->leads = c.take<snfb_lead>(n); ctx->ev_buf = c.take<extract::Event>(n); ctx->sorted_leads = c.take<snfb_lead>(n);
    b.key0 = c.take<uint64_t>(n); b.val0 = c.take<uint32_t>(n); b.key1 = c.take<uint64_t>(n); b.val1 = c.take<uint32_t>(n); b.flag = c.take<uint32_t>(n); b.scan = c.take<uint32_t>(n);
    b.scan_tmp = c.take<uint32_t>(prims::scan_tmp_elems(std::max<unsigned long long>(n, prims::radix_hist_elems(n))) + 16);
    ctx->radix_hist = c.take<uint32_t>(prims::radix_hist_elems(n) + 16);
    b.bin_start = c.take<uint32_t>(n); b.bin_nl = c.take<uint32_t>(n); b.bin_nlong = c.take<uint32_t>(n); b.bin_kept = c.take<uint32_t>(n); b.bin_hap = c.take<uint32_t>(3 * n);
    b.kl_off = c.take<uint32_t>(n); b.kll_off = c.take<uint32_t>(n); b.kb_idx = c.take<uint32_t>(n); b.kl = c.take<uint32_t>(n); b.kll = c.take<uint32_t>(n); b.kleads = c.take<snfb_lead>(n); b.klleads = c.take<snfb_lead>(n);
    b.kb_bin = c.take<uint32_t>(n); b.kb_lead_off = c.take<uint32_t>(n); b.kb_lead_n = c.take<uint32_t>(n); b.kb_long_off = c.take<uint32_t>(n); b.kb_long_n = c.take<uint32_t>(n); b.kb_seed = c.take<int32_t>(n); b.kb_chain = c.take<uint32_t>(n); b.kb_repeat = c.take<uint8_t>(n);
    b.seg_start = c.take<uint32_t>(n); b.c_next = c.take<uint32_t>(n); b.c_last = c.take<uint32_t>(n); b.c_sd = c.take<double>(n); b.c_mean = c.take<double>(n); b.c_rep = c.take<uint8_t>(n);
    b.seg_sd_last = c.take<double>(n); b.seg_maxsd_first = c.take<double>(n); b.cl_first = c.take<uint32_t>(n); b.cl_last = c.take<uint32_t>(n); b.cl_rep = c.take<uint8_t>(n); b.big_list = c.take<uint32_t>(n); b.mid_list = c.take<uint32_t>(n);
    b.g_khi = c.take<uint64_t>(n); b.g_klo = c.take<uint64_t>(n); b.g_u32 = c.take<uint32_t>((size_t)cluster::coop::NU32 * n);
    b.ord = c.take<uint32_t>(n); b.st_leads = c.take<snfb_lead>(n); b.st_plo = c.take<uint32_t>(n); b.st_pn = c.take<uint32_t>(n); b.st_rn = c.take<uint64_t>(n); b.cand_tmp = c.take<snfb_cand>(n); b.sub_valid = c.take<uint8_t>(n);
    b.cl_nsub = c.take<uint32_t>(n); b.cl_nvalid = c.take<uint32_t>(n); b.cl_nlead = c.take<uint32_t>(n); b.cl_nrn = c.take<uint32_t>(n); b.cl_cand_base = c.take<uint32_t>(n); b.cl_lead_base = c.take<uint32_t>(n); b.cl_rn_base = c.take<uint32_t>(n);
    ctx->arena_off = c.take<uint32_t>(n);
}
static void carve_c(snfb_ctx* ctx, Carver& c) {
    const Caps& k = ctx->cap; cluster::B& b = ctx->B; consensus::C& cc = ctx->Cc;
    b.cand = c.take<snfb_cand>(k.cand + 1); b.cand_leads = c.take<snfb_lead>(k.cand_lead + 1); b.out_plo = c.take<uint32_t>(k.cand_lead + 1); b.out_pn = c.take<uint32_t>(k.cand_lead + 1);
    b.rnames = c.take<uint64_t>(k.rn + 1); b.rn_off_out = c.take<uint32_t>(k.cand + 2);
    cc.plan_best = c.take<uint32_t>(k.cand + 1); cc.plan_nother = c.take<uint32_t>(k.cand + 1); cc.plan_otot = c.take<uint32_t>(k.cand + 1); cc.alt_len = c.take<uint32_t>(k.cand + 1); cc.scr_len = c.take<uint32_t>(k.cand + 1);
    cc.alt_off = c.take<uint32_t>(k.cand + 1); cc.scr_off = c.take<uint32_t>(k.cand + 1); cc.work_big = c.take<uint32_t>(k.cand + 1); cc.work_small = c.take<uint32_t>(k.cand + 1); cc.work_ctr = c.take<uint32_t>(64);
    cc.items_big = c.take<consensus::C::Item>(k.item + 1); cc.items_small = c.take<consensus::C::Item>(k.item + 1); cc.tiles = c.take<uint2>(k.tile + 1);
    cc.alt = c.take<uint8_t>(k.alt + 64); cc.scr = c.take<uint8_t>(k.scr16 * 16 + 64);
    cc.dbg = getenv("SNFB_DEBUG") ? c.take<unsigned long long>(148 * 9 * consensus::ALIGN_WARPS * 8 + 8) : nullptr;
    ctx->seq_req = c.take<consensus::SeqReq>(k.req + 1); ctx->seq_arena = c.take<uint8_t>(k.req16 * 16 + 64);
}
static int ensure_arenas(snfb_ctx* ctx) {
    if (ctx->arena_r_for != ctx->n_rec + 1 || ctx->arena_r_tasks != ctx->n_task || ctx->arena_r_cigar != ctx->n_cigar || !ctx->arena_r.p) {
        Carver m; carve_r(ctx, m);
        if (ctx->arena_r.ensure(m.off + 256)) return fail(ctx, "out of device memory (per-record arrays)");
        Carver a; a.base = ctx->arena_r.as<uint8_t>(); carve_r(ctx, a);
        ctx->arena_r_for = ctx->n_rec + 1; ctx->arena_r_tasks = ctx->n_task; ctx->arena_r_cigar = ctx->n_cigar;
    }
    if (ctx->arena_l_for.lead != ctx->cap.lead || !ctx->arena_l.p) {
        Carver m; carve_l(ctx, m);
        if (ctx->arena_l.ensure(m.off + 256)) return fail(ctx, "out of device memory (per-lead arrays)");
        Carver a; a.base = ctx->arena_l.as<uint8_t>(); carve_l(ctx, a);
        ctx->arena_l_for = ctx->cap;
    }
    const Caps& k = ctx->cap; const Caps& f = ctx->arena_c_for;
    if (!ctx->arena_c.p || f.cand != k.cand || f.cand_lead != k.cand_lead || f.rn != k.rn || f.alt != k.alt || f.scr16 != k.scr16 || f.item != k.item || f.tile != k.tile || f.req != k.req || f.req16 != k.req16) {
        Carver m; carve_c(ctx, m);
        if (ctx->arena_c.ensure(m.off + 256)) return fail(ctx, "out of device memory (candidate / consensus arrays)");
        Carver a; a.base = ctx->arena_c.as<uint8_t>(); carve_c(ctx, a);
        ctx->arena_c_for = ctx->cap;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ stage drivers (no host synchronisation inside)
static void bind_inputs(snfb_ctx* ctx) {
    cluster::B& b = ctx->B;
    b.leads = ctx->leads; b.rec = ctx->d_rec; b.task = ctx->b_task.as<snfb_task>(); b.contig = ctx->b_contig.as<snfb_contig>();
    b.tr = ctx->n_tr ? ctx->b_tr.as<int32_t>() : nullptr; b.tr_pmax = ctx->b_trp.as<int32_t>();
    b.rec_pos = ctx->rec_pos; b.rec_end = ctx->rec_end; b.rec_flags = ctx->rec_flags; b.rec_nm = ctx->rec_nm; b.rec_nlead = ctx->rec_nlead; b.rec_lead_off = ctx->rec_lead_off;
    b.task_first = ctx->task_first; b.task_last = ctx->task_last; b.task_maxspan = ctx->task_span;
    b.mask = ctx->n_mask ? ctx->b_mask.as<int32_t>() : nullptr; b.mask_task_off = ctx->n_mask ? ctx->b_mask_off.as<uint32_t>() : nullptr;
    b.n_task = ctx->n_task; b.n_bound = ctx->cap.lead; b.ctr = ctx->b_ctr.as<DevCounters>(); b.cfg = ctx->cfg;
    b.cut_gap = ctx->force_no_cuts ? INT_MAX : cluster::break_gap(ctx->cfg);
    b.cand_cap = ctx->cap.cand; b.cand_lead_cap = ctx->cap.cand_lead; b.rn_cap = ctx->cap.rn;
    consensus::C& c = ctx->Cc;
    c.cand = b.cand; c.cand_rw = b.cand; c.cand_leads = b.cand_leads; c.out_plo = b.out_plo; c.out_pn = b.out_pn; c.ord = b.ord; c.kleads = b.kleads; c.rec = ctx->d_rec; c.seq = ctx->d_seq; c.arena_off = nullptr;
    c.cand_cap = ctx->cap.cand; c.ctr = b.ctr; c.cfg = ctx->cfg; c.item_cap = ctx->cap.item; c.tile_cap = ctx->cap.tile; c.alt_cap = ctx->cap.alt; c.scr_cap16 = ctx->cap.scr16;
}

// stage A plus the bin sort: everything LeadProvider.build_leadtab leaves behind
static int enqueue_stage_a(snfb_ctx* ctx) {
    const uint64_t nrec = ctx->n_rec; const uint32_t nt = ctx->n_task; const snfb_config& cf = ctx->cfg; cluster::B& b = ctx->B;
    DevCounters* ctr = b.ctr; cudaStream_t st = ctx->st;
    CUDA_TRY(cudaMemsetAsync(ctr, 0, sizeof(DevCounters), st));
    CUDA_TRY(cudaMemsetAsync(ctx->task_first, 0, 4 * nt, st)); CUDA_TRY(cudaMemsetAsync(ctx->task_last, 0, 4 * nt, st));
    CUDA_TRY(cudaMemsetAsync(ctx->task_reads, 0, 4 * nt, st)); CUDA_TRY(cudaMemsetAsync(ctx->task_cov, 0, 8 * nt, st)); CUDA_TRY(cudaMemsetAsync(ctx->task_span, 0, 4 * nt, st));
    CUDA_TRY(cudaMemsetAsync(ctx->task_nm, 0, 8 * nt, st));
    extract::ChunkParams S{};
    S.scan = ctx->scanrec; S.choff = ctx->choff; S.n_rec = (uint32_t)nrec; S.cigar = ctx->d_cigar; S.task = b.task; S.cdesc = ctx->cdesc; S.csum = ctx->csum; S.flist = ctx->flist; S.fcnt = ctx->fcnt;
    S.rec_foff = ctx->rec_foff; S.rec_nf = ctx->rec_nf; S.n_chunks = &ctr->n_chunks; S.n_flag = &ctr->n_flagged; S.chunk_cap = ctx->chunk_cap;
    S.rec_end = ctx->rec_end; S.rec_nlead = ctx->rec_nlead; S.rec_big = ctx->rec_big;
    S.ev = ctx->ev_buf; S.ev_cap = ctx->cap.lead; S.n_ev = &ctr->n_ev; S.sa_list = ctx->sa_list; S.n_sa = &ctr->n_sa; S.ctr = ctr; S.minsv = cf.minsvlen_screen;
    if (evt_need(ctx) < ctx->evt_min) {
        // the block's E bits were set for longer events than this configuration looks at: lower the threshold in place
        if (ctx->on_device) return fail(ctx, "the device-resident CIGAR16 arena was packed with a larger event length than the configuration needs: repack with snfb_pack_cigar16(evt_min)");
        extract::k_reflag<<<148 * 8, 256, 0, st>>>(const_cast<uint16_t*>(ctx->d_cigar), ctx->n_cigar, evt_need(ctx)); LAUNCHED(ctx, 1);
        ctx->evt_min = evt_need(ctx);
    }
    mark(ctx, "k_rec_index");
    if (nrec) {
        k_validate<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(ctx->d_rec, (uint32_t)nrec, nt, ctx->n_cigar, ctx->n_var, ctx->n_seq, ctx->seq_on_demand ? 0 : 1, ctr);
        extract::IndexParams I{};
        I.rec = ctx->d_rec; I.cigar = ctx->d_cigar; I.task = b.task; I.n_rec = (uint32_t)nrec; I.n_task = nt; I.rec_pos = ctx->rec_pos; I.task_first = ctx->task_first; I.task_last = ctx->task_last;
        I.scan = ctx->scanrec; I.clip = ctx->clip; I.rec_end = ctx->rec_end; I.rec_flags = ctx->rec_flags; I.rec_nm = ctx->rec_nm; I.rec_nlead = ctx->rec_nlead; I.ctr = ctr;
        I.mapq_min = cf.mapq; I.alen_min = cf.min_alignment_length; I.excl = cf.exclude_flags; I.want_nm = (cf.qc_nm_measure || cf.phase) ? 1 : 0; I.n_cigar = ctx->n_cigar;
        I.pass_chunks = ctx->pass_chunks;
        extract::k_rec_index<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(I);
        // the chunk table: every passing record's CIGAR16 groups in chunks of CH, one thread of the streaming kernel per chunk
        LAUNCHED(ctx, prims::exclusive_scan(ctx->pass_chunks, ctx->choff, ctx->scan_tmp_r, nullptr, nrec, &ctr->n_chunks, st));
        extract::k_cdesc<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(S); LAUNCHED(ctx, 2);
        // algorithmic bytes of the streaming kernel: chunk descriptors + CIGAR16 words + its per-chunk sums (bench.py adds the last two from the counters)
        mark(ctx, "k_scan", 2 * ctx->n_cigar);
        extract::k_chunk_sum<<<148 * 8, 256, 0, st>>>(S);
        mark(ctx, "k_scan_rare");
        extract::k_rec_base<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(S);
        extract::k_chunk_rare<<<148 * 8, 128, 0, st>>>(S);
        extract::k_rec_fin<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(S); LAUNCHED(ctx, 4);
        mark(ctx, "k_rec_post");
        extract::PostParams Q{};
        Q.scan = I.scan; Q.clip = I.clip; Q.task = b.task; Q.n_rec = (uint32_t)nrec; Q.rec_end = ctx->rec_end; Q.rec_big = ctx->rec_big; Q.rec_nm = ctx->rec_nm;
        Q.task_reads = ctx->task_reads; Q.task_cov_bp = ctx->task_cov; Q.task_maxspan = ctx->task_span;
        extract::k_rec_post<<<(unsigned)((nrec + 255) / 256), 256, 0, st>>>(Q);
        mark(ctx, "k_emit");
        extract::EmitParams E{};
        E.rec = ctx->d_rec; E.clip = ctx->clip; E.var = ctx->d_var; E.ev = S.ev; E.n_ev = S.n_ev; E.ev_cap = ctx->cap.lead; E.fcnt = ctx->fcnt;
        E.leads = ctx->leads; E.ctr = ctr; E.maxlen = cf.dev_seq_cache_maxlen; E.detect_large_ins = cf.detect_large_ins; E.longinslen = (double)cf.long_ins_length / 2.0;
        extract::k_emit<<<148 * 16, 128, 0, st>>>(E);
        mark(ctx, "k_sa");
        extract::SaParams A{};
        A.rec = ctx->d_rec; A.clip = ctx->clip; A.var = ctx->d_var; A.task = b.task; A.contig = b.contig; A.n_contig = ctx->n_contig;
        A.sa_list = S.sa_list; A.n_sa = S.n_sa; A.rec_end = ctx->rec_end; A.rec_nlead = ctx->rec_nlead; A.leads = E.leads; A.lead_cap = ctx->cap.lead; A.ctr = ctr; A.cfg = cf; A.seg_scratch = ctx->sa_seg;
        extract::k_sa<<<extract::SA_BLOCKS, extract::SA_THREADS, 0, st>>>(A);
        mark(ctx, "k_task_nm");
        const int cpt = (int)((nrec + extract::NM_CHUNK - 1) / extract::NM_CHUNK);
        extract::k_nm_partial<<<dim3(cpt, nt), 256, 0, st>>>(ctx->rec_flags, ctx->rec_nm, ctx->task_first, ctx->task_last, ctx->nm_part, ctx->nm_cnt);
        extract::k_task_nm<<<nt, 256, 0, st>>>(ctx->task_first, ctx->task_last, ctx->nm_part, ctx->nm_cnt, cpt, ctx->task_nm); LAUNCHED(ctx, 8);
    }
    mark(ctx, "scan_rec_leads");
    LAUNCHED(ctx, prims::exclusive_scan(ctx->rec_nlead, ctx->rec_lead_off, ctx->scan_tmp_r, nullptr, nrec, &ctr->n_leads, st));
    const unsigned long long nb = b.n_bound; const int g = grid_for(nb, 256);
    mark(ctx, "sort_leads");
    cluster::k_scatter_keys<<<g, 256, 0, st>>>(b);
    prims::RadixTemp rt{ ctx->radix_hist, b.scan_tmp };
    bool first = true;
    LAUNCHED(ctx, 1 + prims::radix_sort(b.key0, b.val0, b.key1, b.val1, rt, &ctr->n_leads, nb, cluster::TASK_SHIFT + bits_for(ctx->n_task), &first, st));
    b.skey = first ? b.key0 : b.key1; b.sval = first ? b.val0 : b.val1;
    mark(ctx, "bins");
    cluster::k_bin_heads<<<g, 256, 0, st>>>(b);
    LAUNCHED(ctx, prims::exclusive_scan(b.flag, b.scan, b.scan_tmp, &ctr->n_leads, nb, &ctr->n_bins, st));
    cluster::k_bin_build<<<g, 256, 0, st>>>(b);
    cluster::k_bin_stats<<<g, 256, 0, st>>>(b); LAUNCHED(ctx, 3);
    mark(ctx, nullptr);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int enqueue_stage_b(snfb_ctx* ctx) {
    cluster::B& b = ctx->B; DevCounters* ctr = b.ctr; cudaStream_t st = ctx->st;
    const unsigned long long nb = b.n_bound; const int g = grid_for(nb, 128);
    mark(ctx, "kept_bins");
    LAUNCHED(ctx, prims::exclusive_scan(b.bin_nl, b.kl_off, b.scan_tmp, &ctr->n_bins, nb, &ctr->n_kl, st));
    LAUNCHED(ctx, prims::exclusive_scan(b.bin_nlong, b.kll_off, b.scan_tmp, &ctr->n_bins, nb, &ctr->n_kll, st));
    LAUNCHED(ctx, prims::exclusive_scan(b.bin_kept, b.kb_idx, b.scan_tmp, &ctr->n_bins, nb, &ctr->n_kbins, st));
    cluster::k_kbin_build<<<g, 128, 0, st>>>(b);
    cluster::k_gather_kept<<<grid_for(nb, 256), 256, 0, st>>>(b);
    mark(ctx, "merge_chains");
    cluster::k_seg_heads<<<g, 128, 0, st>>>(b);
    LAUNCHED(ctx, prims::exclusive_scan(b.flag, b.scan, b.scan_tmp, &ctr->n_kbins, nb, &ctr->n_segs, st));
    cluster::k_seg_build<<<g, 128, 0, st>>>(b);
    cluster::k_merge<<<g, 128, 0, st>>>(b);
    cluster::k_verify_cuts<<<g, 128, 0, st>>>(b);
    LAUNCHED(ctx, prims::exclusive_scan(b.flag, b.scan, b.scan_tmp, &ctr->n_kbins, nb, &ctr->n_clusters, st));
    cluster::k_cluster_build<<<g, 128, 0, st>>>(b);
    mark(ctx, "cluster_call");
    // clusters too large for one warp's shared memory go to a block each, next to the warp-per-cluster kernel
    CUDA_TRY(cudaEventRecord(ctx->ev_fork, st)); CUDA_TRY(cudaStreamWaitEvent(ctx->st_side, ctx->ev_fork, 0));
    cluster::k_cluster_block<<<148, cluster::CB_THREADS, cluster::CB_SMEM, ctx->st_side>>>(b);
    cluster::k_cluster_warp<cluster::WARP_CAP, cluster::CWM_WARPS, true><<<148 * 2, cluster::CWM_WARPS * 32, cluster::CwCfg<cluster::WARP_CAP, cluster::CWM_WARPS>::smem, ctx->st_side>>>(b);
    CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->st_side));
    cluster::k_cluster_warp<cluster::SMALL_CAP, cluster::CWS_WARPS, false><<<148 * 4, cluster::CWS_WARPS * 32, cluster::CwCfg<cluster::SMALL_CAP, cluster::CWS_WARPS>::smem, st>>>(b);
    CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    mark(ctx, "emit_cands");
    LAUNCHED(ctx, prims::exclusive_scan(b.cl_nvalid, b.cl_cand_base, b.scan_tmp, &ctr->n_clusters, nb, &ctr->n_cand, st));
    LAUNCHED(ctx, prims::exclusive_scan(b.cl_nlead, b.cl_lead_base, b.scan_tmp, &ctr->n_clusters, nb, &ctr->n_cand_leads, st));
    LAUNCHED(ctx, prims::exclusive_scan(b.cl_nrn, b.cl_rn_base, b.scan_tmp, &ctr->n_clusters, nb, &ctr->n_rnames, st));
    cluster::k_emit_cands<<<148 * 8, 128, 0, st>>>(b);
    mark(ctx, "coverage");
    if (ctx->n_mask) { cluster::k_mask_bp<<<grid_for((unsigned long long)ctx->n_mask * 32, 128), 128, 0, st>>>(b, ctx->b_mask_task.as<uint32_t>(), ctx->n_mask, ctx->task_cov); LAUNCHED(ctx, 1); }
    cluster::k_coverage<<<148 * 32, 128, 0, st>>>(b); LAUNCHED(ctx, 13);
    // consensus plan: best read per INS candidate, sizes and offsets of the ALT bytes and of the scratch; the candidate records are final after this
    consensus::C& c = ctx->Cc;
    CUDA_TRY(cudaMemsetAsync(c.work_ctr, 0, 64, st));
    mark(ctx, "consensus_plan");
    consensus::k_plan<<<grid_for(ctx->cap.cand, 128), 128, 0, st>>>(c);
    LAUNCHED(ctx, prims::exclusive_scan(c.alt_len, c.alt_off, b.scan_tmp, &ctr->n_cand, ctx->cap.cand, &ctr->n_alt_bytes, st));
    LAUNCHED(ctx, prims::exclusive_scan(c.scr_len, c.scr_off, b.scan_tmp, &ctr->n_cand, ctx->cap.cand, &ctr->n_seq_bytes, st));
    consensus::k_plan_finish<<<grid_for(ctx->cap.cand, 128), 128, 0, st>>>(c); LAUNCHED(ctx, 2);
    mark(ctx, nullptr);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// the consensus kernels; with the seq arena on the host ("seq on demand") the base slices are requested, gathered and uploaded first
static int enqueue_stage_c(snfb_ctx* ctx) {
    cluster::B& b = ctx->B; consensus::C& c = ctx->Cc; DevCounters* ctr = b.ctr; cudaStream_t st = ctx->st;
    c.seq = ctx->d_seq; c.arena_off = nullptr; ctx->seq_h2d_bytes = 0;
    if (ctx->seq_on_demand) {
        // the device lists the base slices it will read, the host gathers exactly those bytes from its arena into pinned staging,
        // one H2D copy brings them in (PCIe bytes ~ algorithmic bytes instead of the whole arena).  This path needs the host in the loop.
        mark(ctx, "seq_requests");
        consensus::k_seq_requests<<<grid_for(ctx->cap.cand, 128), 128, 0, st>>>(c, ctx->seq_req, ctx->cap.req, ctx->arena_off, &ctr->n_req, &ctr->n_req_units); LAUNCHED(ctx, 1);
        mark(ctx, nullptr);
        CUDA_TRY(cudaMemcpyAsync(ctx->h_fin, ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const unsigned long long nreq = ctx->h_fin->n_req, nunits = ctx->h_fin->n_req_units;
        if (nreq <= ctx->cap.req && nunits <= ctx->cap.req16 && nreq) {
            if (ctx->h_seq_req.ensure(sizeof(consensus::SeqReq) * (nreq + 1)) || ctx->h_seq_arena.ensure(nunits * 16 + 64)) return fail(ctx, "out of pinned memory (seq arena)");
            CUDA_TRY(cudaMemcpyAsync(ctx->h_seq_req.p, ctx->seq_req, sizeof(consensus::SeqReq) * nreq, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            const consensus::SeqReq* rq = ctx->h_seq_req.as<consensus::SeqReq>(); uint8_t* dst = ctx->h_seq_arena.as<uint8_t>(); const uint8_t* src = ctx->h_seq; const uint64_t nseq = ctx->n_seq;
            #pragma omp parallel for schedule(static, 256)
            for (long long i = 0; i < (long long)nreq; ++i) {
                unsigned long long s0 = rq[i].src; unsigned long long nbytes = rq[i].nbytes; if (s0 > nseq) s0 = nseq; if (s0 + nbytes > nseq) nbytes = nseq - s0;
                memcpy(dst + (size_t)rq[i].dst16 * 16, src + s0, (size_t)nbytes);
            }
            mark(ctx, "h2d_seq_slices", nunits * 16);
            CUDA_TRY(cudaMemcpyAsync(ctx->seq_arena, ctx->h_seq_arena.p, nunits * 16, cudaMemcpyHostToDevice, st));
            mark(ctx, nullptr);
            ctx->seq_h2d_bytes = nunits * 16;
        }
        c.seq = ctx->seq_arena; c.arena_off = ctx->arena_off;
    }
    mark(ctx, "consensus");
    consensus::k_prep<<<148 * 8, 128, 0, st>>>(c);
    mark(ctx, "consensus_align");
    consensus::k_align<<<148 * 7, consensus::ALIGN_WARPS * 32, 0, st>>>(c);
    mark(ctx, "consensus_vote");
    consensus::k_vote<<<148 * 16, consensus::VOTE_THREADS, 0, st>>>(c); LAUNCHED(ctx, 3);
    mark(ctx, nullptr);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ capacities
static void initial_caps(snfb_ctx* ctx) {
    Caps& k = ctx->cap;
    const unsigned long long nrec = ctx->n_rec;
    const unsigned long long lead = std::max<unsigned long long>(1ull << 16, nrec) + (unsigned long long)extract::SA_BLOCKS * extract::SA_THREADS * extract::SLOT_CHUNK / 4;
    if (k.lead < lead) k.lead = lead;
    if (k.cand < k.lead / 8 + 1024) k.cand = k.lead / 8 + 1024;
    if (k.cand_lead < k.lead / 2 + 1024) k.cand_lead = k.lead / 2 + 1024;
    if (k.rn < k.lead / 2 + 1024) k.rn = k.lead / 2 + 1024;
    if (k.alt < (4ull << 20)) k.alt = 4ull << 20;
    if (k.scr16 < (2ull << 20)) k.scr16 = 2ull << 20;
    if (k.item < k.cand_lead / 2 + 1024) k.item = k.cand_lead / 2 + 1024;
    if (k.tile < k.cand + 4096) k.tile = k.cand + 4096;
    if (ctx->seq_on_demand) { if (k.req < k.cand_lead / 2 + 1024) k.req = k.cand_lead / 2 + 1024; if (k.req16 < (1ull << 20)) k.req16 = 1ull << 20; }
    else { if (k.req < 16) k.req = 16; if (k.req16 < 16) k.req16 = 16; }
}
static unsigned long long grown(unsigned long long need) { return need + need / 4 + 1024; }
// does everything the counters report fit the capacities the run used?  If not, raise them (what was needed + 25 %).
static bool caps_fit(snfb_ctx* ctx, const DevCounters& c, const uint32_t* work, int upto) {
    Caps& k = ctx->cap; bool ok = true;
    const unsigned long long need_lead = std::max(c.n_slots, c.n_leads);
    if (need_lead > k.lead || c.lead_overflow) { k.lead = std::max(grown(need_lead), k.lead + k.lead / 2); ok = false; }
    if (upto >= 2) {
        if (c.n_cand > k.cand) { k.cand = grown(c.n_cand); ok = false; }
        if (c.n_cand_leads > k.cand_lead) { k.cand_lead = grown(c.n_cand_leads); ok = false; }
        if (c.n_rnames > k.rn) { k.rn = grown(c.n_rnames); ok = false; }
        if (c.n_alt_bytes > k.alt) { k.alt = grown(c.n_alt_bytes); ok = false; }
        if (c.n_seq_bytes > k.scr16) { k.scr16 = grown(c.n_seq_bytes); ok = false; }
        const unsigned long long items = std::max(work[4], work[5]);
        if (items > k.item) { k.item = grown(items); ok = false; }
        if (work[8] > k.tile) { k.tile = grown(work[8]); ok = false; }
    }
    if (upto >= 3) {
        if (c.n_req > k.req) { k.req = grown(c.n_req); ok = false; }
        if (c.n_req_units > k.req16) { k.req16 = grown(c.n_req_units); ok = false; }
        if (ok && c.scratch_overflow) { ok = false; k.alt = grown(k.alt); k.scr16 = grown(k.scr16); }      // should not happen: every capacity above fit
    } else if (ok && c.scratch_overflow) { ok = false; k.cand_lead = grown(k.cand_lead); k.rn = grown(k.rn); }
    return ok;
}

// ------------------------------------------------------------------------------------------------ views
static int fill_lead_view(snfb_ctx* ctx, snfb_lead_view* out) {
    const DevCounters& c = *ctx->h_fin; const unsigned long long nl = c.n_leads; const uint32_t nt = ctx->n_task;
    if (ctx->h_leads.ensure(sizeof(snfb_lead) * (nl + 1)) || ctx->h_task_reads.ensure(4 * nt) || ctx->h_task_nm.ensure(8 * nt) || ctx->h_rec_nm.ensure(8 * (ctx->n_rec + 1)))
        return fail(ctx, "out of memory for the lead view");
    if (nl) {
        k_gather_leads<<<grid_for(nl, 256), 256, 0, ctx->st>>>(ctx->leads, ctx->B.sval, ctx->sorted_leads, &ctx->B.ctr->n_leads, ctx->cap.lead); LAUNCHED(ctx, 1);
        CUDA_TRY(cudaMemcpyAsync(ctx->h_leads.p, ctx->sorted_leads, sizeof(snfb_lead) * nl, cudaMemcpyDeviceToHost, ctx->st));
    }
    CUDA_TRY(cudaMemcpyAsync(ctx->h_task_reads.p, ctx->task_reads, 4 * nt, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaMemcpyAsync(ctx->h_task_nm.p, ctx->task_nm, 8 * nt, cudaMemcpyDeviceToHost, ctx->st));
    if (ctx->n_rec) CUDA_TRY(cudaMemcpyAsync(ctx->h_rec_nm.p, ctx->rec_nm, 8 * ctx->n_rec, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    out->n_leads = nl; out->leads = ctx->h_leads.as<snfb_lead>();
    out->task_read_count = ctx->h_task_reads.as<uint32_t>(); out->task_mean_nm = ctx->h_task_nm.as<double>(); out->rec_nm = ctx->h_rec_nm.as<double>();
    uint64_t np = 0; for (uint32_t t = 0; t < nt; ++t) np += out->task_read_count[t];
    out->n_pass = np; out->soft_errors = c.soft_errors;
    return 0;
}
// device -> host copies of the candidate view on `stream`, sized by the counters `c`
static int enqueue_cand_copies(snfb_ctx* ctx, const DevCounters& c, cudaStream_t stream) {
    cluster::B& b = ctx->B; const uint32_t nt = ctx->n_task;
    if (ctx->h_cand.ensure(sizeof(snfb_cand) * (c.n_cand + 1)) || ctx->h_cand_leads.ensure(sizeof(snfb_lead) * (c.n_cand_leads + 1)) || ctx->h_rnames.ensure(8 * (c.n_rnames + 1)) || ctx->h_rn_off.ensure(4 * (c.n_cand + 2))
        || ctx->h_task_cov.ensure(8 * nt) || ctx->h_task_cov_raw.ensure(8 * nt)) return fail(ctx, "out of pinned memory for the candidate view");
    if (c.n_cand) { CUDA_TRY(cudaMemcpyAsync(ctx->h_cand.p, b.cand, sizeof(snfb_cand) * c.n_cand, cudaMemcpyDeviceToHost, stream)); CUDA_TRY(cudaMemcpyAsync(ctx->h_rn_off.p, b.rn_off_out, 4 * c.n_cand, cudaMemcpyDeviceToHost, stream)); }
    if (c.n_cand_leads) CUDA_TRY(cudaMemcpyAsync(ctx->h_cand_leads.p, b.cand_leads, sizeof(snfb_lead) * c.n_cand_leads, cudaMemcpyDeviceToHost, stream));
    if (c.n_rnames) CUDA_TRY(cudaMemcpyAsync(ctx->h_rnames.p, b.rnames, 8 * c.n_rnames, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(ctx->h_task_cov_raw.p, ctx->task_cov, 8 * nt, cudaMemcpyDeviceToHost, stream));
    return 0;
}
static void finish_cand_view(snfb_ctx* ctx, const DevCounters& c, snfb_cand_view* out) {
    const uint32_t nt = ctx->n_task;
    ctx->h_rn_off.as<uint32_t>()[c.n_cand] = (uint32_t)c.n_rnames;
    // coverage_average_total: integer base-pair sum / contig length, one rounding (postprocessing.py:130)
    double* cm = ctx->h_task_cov.as<double>(); const unsigned long long* cov = ctx->h_task_cov_raw.as<unsigned long long>();
    for (uint32_t t = 0; t < nt; ++t) cm[t] = ctx->tasks[t].contig_len > 0 ? (double)cov[t] / (double)ctx->tasks[t].contig_len : 0.0;
    out->n_cand = c.n_cand; out->cand = ctx->h_cand.as<snfb_cand>(); out->n_cand_leads = c.n_cand_leads; out->cand_leads = ctx->h_cand_leads.as<snfb_lead>();
    out->rnames = ctx->h_rnames.as<uint64_t>(); out->rnames_off = ctx->h_rn_off.as<uint32_t>(); out->task_coverage_mean = cm; out->unverified_breaks = c.unverified_breaks;
}

// ------------------------------------------------------------------------------------------------ the run loop
// upto: 1 = stage A (+ sort and bins), 2 = + stage B and the consensus plan, 3 = + consensus.
static int run_pipeline(snfb_ctx* ctx, int upto, snfb_lead_view* leads, snfb_cand_view* cands, snfb_seq_view* seqs) {
    if (!ctx->loaded) return fail(ctx, "no records loaded");
    if (!ctx->have_cfg) return fail(ctx, "snfb_set_config was not called");
    cudaSetDevice(ctx->device);
    for (uint32_t t = 0; t < ctx->n_task; ++t) if ((long long)ctx->tasks[t].contig_len / ctx->cfg.cluster_binsize >= (1ll << 26)) return fail(ctx, "contig_len / cluster_binsize must stay below 2^26 (bin field of the sort key)");
    initial_caps(ctx);
    ctx->stage_a_done = ctx->stage_b_done = ctx->stage_c_done = false;
    for (int attempt = 0; ; ++attempt) {
        if (attempt == 6) return fail(ctx, "buffer capacities did not converge");
        if (ensure_arenas(ctx)) return 1;
        bind_inputs(ctx);
        ctx->n_ev = ctx->n_ev_load;
        cudaStream_t st = ctx->st; DevCounters* ctr = ctx->B.ctr;
        if (enqueue_stage_a(ctx)) return 1;
        bool mid = false, copies = false;
        if (upto >= 2) {
            if (enqueue_stage_b(ctx)) return 1;
            // the host reads the counters on the copy stream while the consensus kernels run
            CUDA_TRY(cudaEventRecord(ctx->ev_b, st)); CUDA_TRY(cudaStreamWaitEvent(ctx->st_copy, ctx->ev_b, 0));
            CUDA_TRY(cudaMemcpyAsync(ctx->h_mid, ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, ctx->st_copy));
            CUDA_TRY(cudaMemcpyAsync(ctx->h_work, ctx->Cc.work_ctr, 64, cudaMemcpyDeviceToHost, ctx->st_copy));
            CUDA_TRY(cudaEventRecord(ctx->ev_mid, ctx->st_copy));
            mid = true;
        }
        if (upto >= 3 && !ctx->seq_on_demand) { if (enqueue_stage_c(ctx)) return 1; }
        if (mid) {
            CUDA_TRY(cudaEventSynchronize(ctx->ev_mid));
            if (ctx->h_mid->bad_records) { cudaStreamSynchronize(st); return fail(ctx, "the record block is malformed: a record points outside its task table or arenas"); }
            if (ctx->h_mid->unsorted) { cudaStreamSynchronize(st); return fail(ctx, "records are not coordinate sorted inside a task"); }
            if (ctx->h_mid->ordinal_overflow) { cudaStreamSynchronize(st); return fail(ctx, "a read carries more than 65535 SV signatures (16-bit lead ordinal)"); }
            if (!caps_fit(ctx, *ctx->h_mid, ctx->h_work, 2)) { CUDA_TRY(cudaStreamSynchronize(st)); ++ctx->reruns; continue; }
            if (ctx->h_mid->unverified_breaks && !ctx->force_no_cuts) { CUDA_TRY(cudaStreamSynchronize(st)); ctx->force_no_cuts = true; ++ctx->reruns; continue; }   // a chain cut was wrong: redo with whole chains
            if (cands) { if (enqueue_cand_copies(ctx, *ctx->h_mid, ctx->st_copy)) return 1; copies = true; }
            if (upto >= 3 && ctx->seq_on_demand) { if (enqueue_stage_c(ctx)) return 1; }
        }
        if (upto >= 3 && seqs) {
            const unsigned long long na = ctx->h_mid->n_alt_bytes;
            if (ctx->h_alt.ensure(na + 16)) return fail(ctx, "out of pinned memory for the ALT arena");
            if (na) CUDA_TRY(cudaMemcpyAsync(ctx->h_alt.p, ctx->Cc.alt, na, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaMemcpyAsync(ctx->h_fin, ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(ctx->h_work + 16, ctx->Cc.work_ctr, 64, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (copies) CUDA_TRY(cudaStreamSynchronize(ctx->st_copy));
        CUDA_TRY(cudaGetLastError());
        if (ctx->h_fin->bad_records) return fail(ctx, "the record block is malformed: a record points outside its task table or arenas");
        if (ctx->h_fin->unsorted) return fail(ctx, "records are not coordinate sorted inside a task");
        if (ctx->h_fin->ordinal_overflow) return fail(ctx, "a read carries more than 65535 SV signatures (16-bit lead ordinal)");
        if (!caps_fit(ctx, *ctx->h_fin, upto >= 2 ? ctx->h_work + 16 : ctx->h_work, upto)) { ++ctx->reruns; continue; }
        break;
    }
    ctx->stage_a_done = true; ctx->stage_b_done = upto >= 2; ctx->stage_c_done = upto >= 3;
    if (leads) { memset(leads, 0, sizeof *leads); if (fill_lead_view(ctx, leads)) return 1; }
    if (cands) { memset(cands, 0, sizeof *cands); finish_cand_view(ctx, *ctx->h_fin, cands); }
    if (seqs) { memset(seqs, 0, sizeof *seqs); seqs->n_alt_bytes = ctx->h_fin->n_alt_bytes; seqs->alt = ctx->h_alt.as<uint8_t>(); }
    return 0;
}

extern "C" {

int snfb_extract_leads(snfb_ctx* ctx, snfb_lead_view* out) {
    if (!ctx) return 1;
    return run_pipeline(ctx, 1, out, nullptr, nullptr);
}
int snfb_cluster_call(snfb_ctx* ctx, snfb_cand_view* out) {
    if (!ctx) return 1;
    if (!ctx->stage_a_done) return fail(ctx, "snfb_extract_leads must run first");
    return run_pipeline(ctx, 2, nullptr, out, nullptr);
}
int snfb_consensus(snfb_ctx* ctx, snfb_seq_view* out) {
    if (!ctx) return 1;
    if (!ctx->stage_b_done) return fail(ctx, "snfb_cluster_call must run first");
    snfb_seq_view tmp; return run_pipeline(ctx, 3, nullptr, nullptr, out ? out : &tmp);
}
int snfb_run(snfb_ctx* ctx, snfb_lead_view* leads, snfb_cand_view* cands, snfb_seq_view* seqs) {
    if (!ctx) return 1;
    snfb_seq_view tmp; return run_pipeline(ctx, 3, leads, cands, seqs ? seqs : &tmp);
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
    if (ctx->n_ev - ctx->n_ev_load >= 2 && n < cap) {   // first stage mark .. last mark: device time of the run
        float t = 0; if (cudaEventElapsedTime(&t, ctx->ev[ctx->n_ev_load], ctx->ev[ctx->n_ev - 1]) != cudaSuccess) t = -1.f;
        names[n] = "total"; ms[n] = t; if (bytes) bytes[n] = 0; ++n;
    }
    return n;
}

int snfb_device_candidates(snfb_ctx* ctx, void** dptr, uint64_t* n_cand) {
    if (!ctx || !ctx->stage_b_done) return 1;
    if (dptr) *dptr = ctx->B.cand; if (n_cand) *n_cand = ctx->h_fin->n_cand; return 0;
}
int snfb_device_alt(snfb_ctx* ctx, void** dptr, uint64_t* n_bytes) {
    if (!ctx || !ctx->stage_c_done) return 1;
    if (dptr) *dptr = ctx->Cc.alt; if (n_bytes) *n_bytes = ctx->h_fin->n_alt_bytes; return 0;
}
uint64_t snfb_launch_count(snfb_ctx* ctx) { return ctx ? ctx->launches : 0; }
double snfb_selftest_sqrt_frac(uint64_t p_hi, uint64_t p_lo, uint64_t q, int slow) { const u128 P = ((u128)p_hi << 64) | p_lo; return slow ? sqrt_frac_rn_slow(P, q) : sqrt_frac_rn(P, q); }
uint64_t snfb_rerun_count(snfb_ctx* ctx) { return ctx ? ctx->reruns : 0; }
/* developer aid (SNFB_DEBUG set when the ctx ran): per-warp timing of the consensus alignment kernel, 8 words per warp */
int snfb_debug_dump(snfb_ctx* ctx, uint64_t* out, uint64_t n_words) {
    if (!ctx || !ctx->Cc.dbg || !out) return 1;
    const uint64_t have = 148ull * 8 * consensus::ALIGN_WARPS * 8; if (n_words > have) n_words = have;
    return cudaMemcpy(out, ctx->Cc.dbg, 8 * n_words, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}
int snfb_pin_host(void* p, size_t bytes) { return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? 0 : 1; }
int snfb_unpin_host(void* p) { return cudaHostUnregister(p) == cudaSuccess ? 0 : 1; }

int snfb_nccl_unique_id(void* out128) {
    std::string e; if (!out128 || !nccl_load(&e)) return 1;
    Id128 id; memset(&id, 0, sizeof id);
    if (g_nccl.GetUniqueId(&id) != 0) return 2;
    memcpy(out128, &id, 128); return 0;
}
int snfb_comm_init(snfb_ctx* ctx, const void* unique_id128, int rank, int nranks) {
    if (!ctx || !unique_id128 || nranks < 1 || rank < 0 || rank >= nranks) return 1;
    std::string e; if (!nccl_load(&e)) return fail(ctx, "NCCL: " + e);
    cudaSetDevice(ctx->device);
    if (ctx->comm) { g_nccl.CommDestroy(ctx->comm); ctx->comm = nullptr; }
    Id128 id; memcpy(&id, unique_id128, 128);
    const int rc = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (rc != 0) { ctx->comm = nullptr; return fail(ctx, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); }
    ctx->rank = rank; ctx->nranks = nranks; ctx->gather_cap = 0; return 0;
}
int snfb_allgather_candidates(snfb_ctx* ctx, uint32_t flags, snfb_gather_view* out) {
    if (!ctx || !out) return 1;
    if (!ctx->stage_c_done) return fail(ctx, "snfb_run must run first");
    if (ctx->nranks > 1 && !ctx->comm) return fail(ctx, "snfb_comm_init was not called");
    cudaSetDevice(ctx->device);
    const int nr = ctx->nranks, with_leads = (flags & SNFB_GATHER_LEADS) ? 1 : 0; cudaStream_t st = ctx->st; cluster::B& b = ctx->B;
    memset(out, 0, sizeof *out);
    if (ctx->h_gather.ensure(64 * 8 + (size_t)nr * (sizeof(GatherHdr) + 8) + 256)) return fail(ctx, "out of pinned memory (gather)");
    unsigned long long* h_layout = ctx->h_gather.as<unsigned long long>();                   // [8] layout, then the header table
    GatherHdr* h_hdr = reinterpret_cast<GatherHdr*>(h_layout + 8);
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (ctx->gather_cap == 0) {
            // first call: the slot size every rank uses is agreed through a small all-gather of what each rank needs
            const DevCounters& c = *ctx->h_fin; GatherHdr h{}; h.n_cand = c.n_cand; h.n_alt = c.n_alt_bytes; h.n_rn = c.n_rnames; h.n_leads = with_leads ? c.n_cand_leads : 0;
            unsigned long long off[6]; gather_offsets(h, off);
            unsigned long long mx = off[5];
            if (nr > 1) {
                if (ctx->b_gsend.ensure(256) || ctx->b_grecv.ensure(8 * (size_t)nr + 256)) return fail(ctx, "out of device memory (gather)");
                CUDA_TRY(cudaMemcpyAsync(ctx->b_gsend.p, &off[5], 8, cudaMemcpyHostToDevice, st));
                if (g_nccl.AllGather(ctx->b_gsend.p, ctx->b_grecv.p, 8, 0 /* ncclChar */, ctx->comm, st) != 0) return fail(ctx, "ncclAllGather (sizes) failed");
                std::vector<unsigned long long> all(nr);
                CUDA_TRY(cudaMemcpyAsync(all.data(), ctx->b_grecv.p, 8 * (size_t)nr, cudaMemcpyDeviceToHost, st)); CUDA_TRY(cudaStreamSynchronize(st));
                for (int r = 0; r < nr; ++r) mx = std::max(mx, all[r]);
            }
            ctx->gather_cap = (grown(mx) + 255ull) & ~255ull;
        }
        const unsigned long long cap = ctx->gather_cap, out_cap = (unsigned long long)nr * cap + 4096ull * 8;
        if (ctx->b_gsend.ensure(cap + 256) || ctx->b_grecv.ensure((size_t)nr * cap + out_cap + 1024)) return fail(ctx, "out of device memory (gather)");
        uint8_t* recv = ctx->b_grecv.as<uint8_t>(); uint8_t* merged = recv + (((size_t)nr * cap + 255) & ~(size_t)255); unsigned long long* d_layout = reinterpret_cast<unsigned long long*>(merged + out_cap);     // layout words live behind the merged arrays
        mark(ctx, "allgather");
        k_gather_pack<<<148 * 4, 256, 0, st>>>(b.ctr, b.cand, ctx->Cc.alt, b.rnames, b.rn_off_out, b.cand_leads, with_leads, nr > 1 ? ctx->b_gsend.as<uint8_t>() : recv, cap); LAUNCHED(ctx, 1);
        if (nr > 1 && g_nccl.AllGather(ctx->b_gsend.p, recv, cap, 0 /* ncclChar */, ctx->comm, st) != 0) return fail(ctx, "ncclAllGather failed");
        k_gather_merge<<<148 * 4, 256, 0, st>>>(recv, cap, nr, with_leads, merged, out_cap, d_layout); LAUNCHED(ctx, 1);
        mark(ctx, nullptr);
        CUDA_TRY(cudaMemcpyAsync(h_layout, d_layout, 64, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_hdr, recv, sizeof(GatherHdr), cudaMemcpyDeviceToHost, st));     // slot 0's header; the rest below
        for (int r = 1; r < nr; ++r) CUDA_TRY(cudaMemcpyAsync(h_hdr + r, recv + (size_t)r * cap, sizeof(GatherHdr), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        unsigned long long mx = 0; for (int r = 0; r < nr; ++r) mx = std::max(mx, h_hdr[r].need_bytes);
        if (h_layout[6]) { ctx->gather_cap = (grown(mx) + 255ull) & ~255ull; ++ctx->reruns; continue; }          // every rank sees the same table: the same new slot size everywhere
        // next call: a slot size every rank derives from the same table
        if (grown(mx) > cap) ctx->gather_cap = (grown(mx) + 255ull) & ~255ull;
        GatherHdr tot{}; for (int r = 0; r < nr; ++r) { tot.n_cand += h_hdr[r].n_cand; tot.n_alt += h_hdr[r].n_alt; tot.n_rn += h_hdr[r].n_rn; tot.n_leads += h_hdr[r].n_leads; }
        uint64_t* rank_n = reinterpret_cast<uint64_t*>(h_hdr + nr); for (int r = 0; r < nr; ++r) rank_n[r] = h_hdr[r].n_cand;
        out->n_cand = tot.n_cand; out->n_alt_bytes = tot.n_alt; out->n_rnames = tot.n_rn; out->n_cand_leads = tot.n_leads; out->rank_n_cand = rank_n;
        out->dev_buffer = merged; out->dev_bytes_per_rank = cap;
        if (!(flags & SNFB_GATHER_DEVICE_ONLY)) {
            const unsigned long long total = h_layout[5];
            if (ctx->h_gather.cap < 64 * 8 + (size_t)nr * (sizeof(GatherHdr) + 8) + 256 + total + 512) {
                // grow while keeping the table: copy it aside
                std::vector<uint8_t> keep((size_t)64 * 8 + (size_t)nr * (sizeof(GatherHdr) + 8)); memcpy(keep.data(), ctx->h_gather.p, keep.size());
                if (ctx->h_gather.ensure(64 * 8 + (size_t)nr * (sizeof(GatherHdr) + 8) + 256 + total + 512)) return fail(ctx, "out of pinned memory (gather result)");
                memcpy(ctx->h_gather.p, keep.data(), keep.size());
                h_layout = ctx->h_gather.as<unsigned long long>(); h_hdr = reinterpret_cast<GatherHdr*>(h_layout + 8); rank_n = reinterpret_cast<uint64_t*>(h_hdr + nr); out->rank_n_cand = rank_n;
            }
            uint8_t* hm = ctx->h_gather.as<uint8_t>() + (((size_t)64 * 8 + (size_t)nr * (sizeof(GatherHdr) + 8) + 255) & ~(size_t)255);
            CUDA_TRY(cudaMemcpyAsync(hm, merged, total, cudaMemcpyDeviceToHost, st)); CUDA_TRY(cudaStreamSynchronize(st));
            out->cand = reinterpret_cast<const snfb_cand*>(hm + h_layout[0]); out->alt = hm + h_layout[1]; out->rnames = reinterpret_cast<const uint64_t*>(hm + h_layout[2]);
            out->rnames_off = reinterpret_cast<const uint32_t*>(hm + h_layout[3]); out->cand_leads = with_leads ? reinterpret_cast<const snfb_lead*>(hm + h_layout[4]) : nullptr;
        }
        return 0;
    }
    return fail(ctx, "gather buffer sizes did not converge");
}

// partial-order-alignment jobs (LocalAsm, local_asm.py:254-304): host buffers in, host buffers out; a block per job
int snfb_poa(snfb_ctx* ctx, const snfb_poa_job* jobs, uint32_t n_jobs, const uint8_t* seqs, uint64_t n_seq_bytes, const int32_t* offs, uint64_t n_offs, uint8_t* out, uint64_t out_bytes, int32_t* out_len) {
    if (!ctx || (n_jobs && (!jobs || !seqs || !offs || !out || !out_len))) return ctx ? fail(ctx, "snfb_poa: null argument") : 1;
    if (n_jobs == 0) return 0;
    cudaSetDevice(ctx->device);
    // scratch: the largest job decides the per-block size; as many blocks as a third of the free memory allows
    size_t smax = 0;
    for (uint32_t k = 0; k < n_jobs; ++k) {
        const snfb_poa_job& j = jobs[k];
        if ((uint64_t)j.offs_off + j.n_seq + 1 > n_offs) return fail(ctx, "snfb_poa: offsets outside offs[]");
        const int32_t* o = offs + j.offs_off; int total = 0, maxl = 0;
        for (uint32_t i = 0; i < j.n_seq; ++i) { const int l = o[i + 1] - o[i]; if (l < 0) return fail(ctx, "snfb_poa: decreasing offsets"); total += l; if (l > maxl) maxl = l; }
        if (j.seq_off + (uint64_t)(j.n_seq ? o[j.n_seq] : 0) > n_seq_bytes) return fail(ctx, "snfb_poa: sequences outside seqs[]");
        if (j.out_off + (uint64_t)j.out_cap * (j.mode == 1 ? 2 : 1) > out_bytes) return fail(ctx, "snfb_poa: output outside out[]");
        if (j.mode == 1 && j.n_seq != 2) return fail(ctx, "snfb_poa: the MSA mode takes exactly two sequences");
        const int bw = 2 * j.band + 1 < maxl ? 2 * j.band + 1 : (maxl > 0 ? maxl : 1);
        smax = std::max(smax, poa::scratch_bytes(total, maxl, bw));
    }
    size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
    size_t nblk = std::min<size_t>(n_jobs, 148 * 2);
    while (nblk > 1 && nblk * smax > free_b / 3) --nblk;
    if (smax > free_b / 2) return fail(ctx, "snfb_poa: a job needs more scratch than the device has free");
    DevBuf d_jobs, d_seqs, d_offs, d_out, d_len, d_scr, d_ctr;
    int rc = d_jobs.ensure(sizeof(snfb_poa_job) * n_jobs) | d_seqs.ensure(n_seq_bytes + 16) | d_offs.ensure(4 * n_offs + 16) | d_out.ensure(out_bytes + 16) | d_len.ensure(4 * (size_t)n_jobs) | d_scr.ensure(nblk * smax) | d_ctr.ensure(64);
    auto done = [&](int r) { d_jobs.release(); d_seqs.release(); d_offs.release(); d_out.release(); d_len.release(); d_scr.release(); d_ctr.release(); return r; };
    if (rc) return done(fail(ctx, "snfb_poa: out of device memory"));
    cudaStream_t st = ctx->st;
    cudaMemcpyAsync(d_jobs.p, jobs, sizeof(snfb_poa_job) * n_jobs, cudaMemcpyHostToDevice, st); cudaMemcpyAsync(d_seqs.p, seqs, n_seq_bytes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_offs.p, offs, 4 * n_offs, cudaMemcpyHostToDevice, st); cudaMemsetAsync(d_ctr.p, 0, 64, st); cudaMemsetAsync(d_out.p, 0, out_bytes, st);
    poa::Params P{}; P.jobs = d_jobs.as<poa::Job>(); P.n_jobs = n_jobs; P.seqs = d_seqs.as<uint8_t>(); P.offs = d_offs.as<int>(); P.out = d_out.as<uint8_t>(); P.out_len = d_len.as<int>();
    P.scratch = d_scr.as<uint8_t>(); P.scratch_per_block = smax; P.next_job = d_ctr.as<unsigned>();
    mark(ctx, "poa");
    poa::k_poa<<<(unsigned)nblk, poa::THREADS, 0, st>>>(P); LAUNCHED(ctx, 1);
    mark(ctx, nullptr);
    cudaMemcpyAsync(out, d_out.p, out_bytes, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(out_len, d_len.p, 4 * (size_t)n_jobs, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(fail(ctx, std::string("snfb_poa: ") + cudaGetErrorString(e)));
    return done(0);
}

// multi-sample combine: every (task, svtype) chain of the plan by one warp (combine.cuh); host buffers in, host buffers out
int snfb_combine_groups(snfb_ctx* ctx, const snfb_combine_in* in, snfb_combine_out* out) {
    if (!ctx || !in || !out) return ctx ? fail(ctx, "snfb_combine_groups: null argument") : 1;
    if (in->n_cand == 0 || in->n_chain == 0) return 0;
    if (!in->chains || !in->chunks || !in->pos || !in->svlen || !in->sample || !out->cand_group || !out->emit_chunk || !out->emit_ord || !out->cov_non)
        return fail(ctx, "snfb_combine_groups: null array");
    if (in->n_samples == 0 || in->bins_per_block <= 0 || in->cov_binsize <= 0) return fail(ctx, "snfb_combine_groups: bad sample count / coverage geometry");
    // the plan must tile [0, n_cand) and [0, n_chunk): chains in order, chunks of a chain consecutive
    { uint64_t c = 0, k = 0;
      for (uint32_t i = 0; i < in->n_chain; ++i) { const snfb_combine_chain& ch = in->chains[i];
          if (ch.cand_off != c || ch.chunk_off != k || ch.n_chunk == 0) return fail(ctx, "snfb_combine_groups: chains do not tile the candidates / chunks");
          uint64_t cc = c;
          for (uint32_t j = 0; j < ch.n_chunk; ++j) { if (k + j >= in->n_chunk) return fail(ctx, "snfb_combine_groups: chunk index out of range"); const snfb_combine_chunk& ck = in->chunks[k + j];
              if ((uint64_t)ck.cand_off != cc || ck.n_cand <= 0 || ck.cov_block < -1 || ck.cov_block >= (int64_t)in->n_cov_block) return fail(ctx, "snfb_combine_groups: bad chunk"); cc += ck.n_cand; }
          if (cc != c + ch.n_cand) return fail(ctx, "snfb_combine_groups: chunk sizes do not add up to the chain");
          if (ch.is_bnd && (!in->mate_contig || !in->mate_pos)) return fail(ctx, "snfb_combine_groups: BND chain without mate arrays");
          c += ch.n_cand; k += ch.n_chunk; }
      if (c != in->n_cand || k != in->n_chunk) return fail(ctx, "snfb_combine_groups: plan does not cover n_cand / n_chunk");
      for (uint32_t i = 0; i < in->n_cand; ++i) if (in->sample[i] >= in->n_samples) return fail(ctx, "snfb_combine_groups: sample index out of range"); }
    cudaSetDevice(ctx->device);
    const size_t n = in->n_cand, S = in->n_samples, W = (S + 31) / 32, ncov = (size_t)in->n_cov_block * S * in->bins_per_block;
    const bool use_alt = in->combine_pctseq != 0.0 && in->alt && in->alt_off && in->alt_len;
    uint32_t max_alt = 16;
    if (use_alt) for (uint32_t i = 0; i < in->n_cand; ++i) { if (in->alt_off[i] + in->alt_len[i] > in->n_alt_bytes) return fail(ctx, "snfb_combine_groups: ALT outside alt[]"); max_alt = std::max(max_alt, in->alt_len[i]); }
    max_alt = (max_alt + 15u) & ~15u;
    const unsigned blocks = (unsigned)std::min<size_t>((in->n_chain + 3) / 4, 148 * 4);
    DevBuf b_in, b_state, b_out;
    // inputs in one buffer, group state in one, outputs in one
    Carver ci, cs, co;
    auto lay = [&](Carver& c, combine::P& P, bool in_, bool st_, bool out_) {
        if (in_) { P.chains = c.take<snfb_combine_chain>(in->n_chain); P.chunks = c.take<snfb_combine_chunk>(in->n_chunk); P.pos = c.take<int32_t>(n); P.svlen = c.take<int32_t>(n); P.sample = c.take<uint32_t>(n);
                   P.mate_contig = c.take<int32_t>(n); P.mate_pos = c.take<int32_t>(n); P.block_start = c.take<long long>(in->n_cov_block + 1); P.cov = c.take<int32_t>(ncov + 1);
                   if (use_alt) { P.alt = c.take<uint8_t>(in->n_alt_bytes + 16); P.alt_off = c.take<unsigned long long>(n); P.alt_len = c.take<uint32_t>(n); } }
        if (st_) { P.g_pos = c.take<double>(n); P.g_len = c.take<double>(n); P.g_mate = c.take<double>(n); P.g_n = c.take<uint32_t>(n); P.g_mc = c.take<int32_t>(n); P.g_incl = c.take<uint32_t>(n * W); P.act = c.take<uint32_t>(n);
                   P.next_chain = c.take<unsigned int>(4); P.g_first = c.take<uint32_t>(n); P.ex_stamp = c.take<uint32_t>(n); if (use_alt) P.hs = c.take<int8_t>((size_t)blocks * 4 * max_alt + 16); }
        if (out_) { P.cand_group = c.take<uint32_t>(n); P.emit_chunk = c.take<int32_t>(n); P.emit_ord = c.take<uint32_t>(n); P.cov_non = c.take<int32_t>(n * S); }
    };
    combine::P P{};
    lay(ci, P, true, false, false); lay(cs, P, false, true, false); lay(co, P, false, false, true);
    auto done = [&](int r) { b_in.release(); b_state.release(); b_out.release(); return r; };
    if (b_in.ensure(ci.off + 256) | b_state.ensure(cs.off + 256) | b_out.ensure(co.off + 256)) return done(fail(ctx, "snfb_combine_groups: out of device memory"));
    ci = Carver{ b_in.as<uint8_t>(), 0 }; cs = Carver{ b_state.as<uint8_t>(), 0 }; co = Carver{ b_out.as<uint8_t>(), 0 };
    lay(ci, P, true, false, false); lay(cs, P, false, true, false); lay(co, P, false, false, true);
    P.n_chain = in->n_chain; P.n_chunk = in->n_chunk; P.n_cand = in->n_cand; P.n_samples = in->n_samples; P.words = (uint32_t)W;
    P.bins_per_block = in->bins_per_block; P.cov_binsize = in->cov_binsize;
    P.combine_match = in->combine_match; P.combine_match_max = in->combine_match_max; P.cluster_merge_bnd = in->cluster_merge_bnd; P.separate_intra = in->combine_separate_intra; P.overlap_abs = in->combine_overlap_abs;
    cudaStream_t st = ctx->st;
    auto up = [&](const void* dst, const void* src, size_t bytes) { if (bytes && src) cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, st); };
    up(P.chains, in->chains, sizeof(snfb_combine_chain) * in->n_chain); up(P.chunks, in->chunks, sizeof(snfb_combine_chunk) * in->n_chunk);
    up(P.pos, in->pos, 4 * n); up(P.svlen, in->svlen, 4 * n); up(P.sample, in->sample, 4 * n); up(P.mate_contig, in->mate_contig, 4 * n); up(P.mate_pos, in->mate_pos, 4 * n);
    up(P.block_start, in->block_start, 8 * (size_t)in->n_cov_block); up(P.cov, in->cov, 4 * ncov);
    P.pctseq = use_alt ? in->combine_pctseq : 0.0; P.max_alt = max_alt;
    if (use_alt) { up(P.alt, in->alt, in->n_alt_bytes); up(P.alt_off, in->alt_off, 8 * n); up(P.alt_len, in->alt_len, 4 * n); }
    cudaMemsetAsync(P.next_chain, 0, 16, st);
    mark(ctx, "combine_groups");
    combine::k_combine<<<blocks, 128, 0, st>>>(P); LAUNCHED(ctx, 1);
    mark(ctx, nullptr);
    cudaMemcpyAsync(out->cand_group, P.cand_group, 4 * n, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(out->emit_chunk, P.emit_chunk, 4 * n, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(out->emit_ord, P.emit_ord, 4 * n, cudaMemcpyDeviceToHost, st); cudaMemcpyAsync(out->cov_non, P.cov_non, 4 * n * S, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(fail(ctx, std::string("snfb_combine_groups: ") + cudaGetErrorString(e)));
    return done(0);
}

int snfb_selftest_edit_distance(snfb_ctx* ctx, const uint8_t* bytes, uint64_t n_bytes, const uint64_t* a_off, const uint32_t* a_len, const uint64_t* b_off, const uint32_t* b_len, uint32_t n_pairs, int32_t* out) {
    if (!ctx || !bytes || !a_off || !a_len || !b_off || !b_len || !out) return ctx ? fail(ctx, "snfb_selftest_edit_distance: null argument") : 1;
    if (n_pairs == 0) return 0;
    uint32_t max_len = 16;
    for (uint32_t i = 0; i < n_pairs; ++i) { if (a_off[i] + a_len[i] > n_bytes || b_off[i] + b_len[i] > n_bytes) return fail(ctx, "snfb_selftest_edit_distance: string outside bytes[]"); max_len = std::max(max_len, std::max(a_len[i], b_len[i])); }
    max_len = (max_len + 15u) & ~15u;
    cudaSetDevice(ctx->device);
    const unsigned blocks = (unsigned)std::min<uint32_t>((n_pairs + 3) / 4, 148 * 4);
    DevBuf d_b, d_o, d_hs, d_out;
    auto done = [&](int r) { d_b.release(); d_o.release(); d_hs.release(); d_out.release(); return r; };
    if (d_b.ensure(n_bytes + 16) | d_o.ensure((size_t)n_pairs * 24 + 64) | d_hs.ensure((size_t)blocks * 4 * max_len + 16) | d_out.ensure((size_t)n_pairs * 4)) return done(fail(ctx, "snfb_selftest_edit_distance: out of device memory"));
    unsigned long long* ao = d_o.as<unsigned long long>(); unsigned long long* bo = ao + n_pairs; uint32_t* al = reinterpret_cast<uint32_t*>(bo + n_pairs); uint32_t* bl = al + n_pairs;
    cudaStream_t st = ctx->st;
    cudaMemcpyAsync(d_b.p, bytes, n_bytes, cudaMemcpyHostToDevice, st); cudaMemcpyAsync(ao, a_off, 8 * (size_t)n_pairs, cudaMemcpyHostToDevice, st); cudaMemcpyAsync(bo, b_off, 8 * (size_t)n_pairs, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(al, a_len, 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, st); cudaMemcpyAsync(bl, b_len, 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, st);
    combine::k_edit_selftest<<<blocks, 128, 0, st>>>(d_b.as<uint8_t>(), ao, al, bo, bl, n_pairs, d_hs.as<int8_t>(), max_len, d_out.as<int>()); LAUNCHED(ctx, 1);
    cudaMemcpyAsync(out, d_out.p, 4 * (size_t)n_pairs, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(fail(ctx, std::string("snfb_selftest_edit_distance: ") + cudaGetErrorString(e)));
    return done(0);
}

// mean coverage of `binsize`-base bins over one task's region (snf.py:248-267: the 500-bp means the SNF writer stores)
int snfb_coverage_bins(snfb_ctx* ctx, uint32_t task, int binsize, const double** out, uint64_t* n_bins) {
    if (!ctx || !ctx->stage_a_done) return ctx ? fail(ctx, "snfb_extract_leads must run first") : 1;
    if (task >= ctx->n_task || binsize <= 0 || !out || !n_bins) return fail(ctx, "bad arguments");
    cudaSetDevice(ctx->device);
    const snfb_task& tk = ctx->tasks[task];
    // the reference pads the contig-long coverage vector with zeros to a multiple of the bin size and takes row means (snf.py:255-256)
    const long long L = tk.contig_len; const long long nb = L > 0 ? (L + binsize - 1) / binsize : 0;
    if (ctx->h_cov_bins.ensure(16 * (size_t)(nb + 1))) return fail(ctx, "out of pinned memory (coverage bins)");
    DevBuf acc; if (acc.ensure(8 * (size_t)(nb + 1))) return fail(ctx, "out of device memory (coverage bins)");
    CUDA_TRY(cudaMemsetAsync(acc.p, 0, 8 * (size_t)(nb + 1), ctx->st));
    uint32_t lohi[2];
    CUDA_TRY(cudaMemcpyAsync(&lohi[0], ctx->task_first + task, 4, cudaMemcpyDeviceToHost, ctx->st)); CUDA_TRY(cudaMemcpyAsync(&lohi[1], ctx->task_last + task, 4, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    if (nb && lohi[1] > lohi[0]) { k_cov_bins<<<grid_for(lohi[1] - lohi[0], 256), 256, 0, ctx->st>>>(ctx->rec_pos, ctx->rec_end, ctx->rec_flags, lohi[0], lohi[1], binsize, L, nb, acc.as<unsigned long long>()); LAUNCHED(ctx, 1); }
    unsigned long long* raw = reinterpret_cast<unsigned long long*>(ctx->h_cov_bins.as<uint8_t>() + 8 * (size_t)(nb + 1));
    CUDA_TRY(cudaMemcpyAsync(raw, acc.p, 8 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(cudaStreamSynchronize(ctx->st));
    acc.release();
    double* o = ctx->h_cov_bins.as<double>();
    for (long long i = 0; i < nb; ++i) o[i] = (double)raw[i] / (double)binsize;
    *out = o; *n_bins = (uint64_t)nb; return 0;
}

}  // extern "C"
