// poa.cuh — partial-order-alignment consensus for the reference's LocalAsm (local_asm.py:254-304, gate parallel.py:186-196),
// the step the reference hands to pyspoa.  One block per job; a job is either "consensus of n read windows" (poa(read_seq, local,
// min_coverage=round(n/2)), local_asm.py:287) or "two-row MSA of (consensus, reference window)" (local_asm.py:289-291).
//
// The algorithm is the one restated in oracle/poa_oracle.c (parity with pyspoa itself is unpinned — the library is not in this image;
// see that file's header): sequences are fused one by one into a partial order graph by a banded local alignment with a two-piece
// affine gap; heaviest-path consensus trimmed to min_coverage; MSA columns in topological order.
//
// Parallelisation: the dynamic programme is the hot part — O(nodes x band) cells per read.  A row (graph node) is computed by the whole
// block: every thread takes a run of consecutive columns, the vertical / diagonal terms come from the predecessor rows (global memory),
// and the horizontal gap states are a prefix maximum over the row (E_j = max_k<j Hn_k + g + (j - 1 - k) e), i.e. one block-wide max-scan
// per row instead of a serial sweep.  Graph bookkeeping (topological order, traceback, fusing the read, heaviest path) is O(nodes) per
// read and runs on one thread of the block.
#pragma once
#include "common.cuh"

namespace poa {

constexpr int MAXIN = 8;
constexpr int NEGV = -(1 << 29);
constexpr int THREADS = 256;

struct Job { unsigned long long seq_off; uint32_t offs_off, n_seq; int min_cov, m, n, g, e, q, c, band; uint32_t mode, out_cap; unsigned long long out_off; };
static_assert(sizeof(Job) == sizeof(snfb_poa_job), "poa::Job mirrors snfb_poa_job");

struct Graph { int* n_; int* ovf_; int cap; uint8_t* base; int* nin; int* in_node; int* in_w; int* col; int* cov; int* aligned; };      // n_ / ovf_: shared-memory words (node count, overflow flag)
#define GN(G) (*(G).n_)
struct Work {                    // per-block scratch, carved by carve()
    Graph G; int* order; int* rank; int* stack; int* edge; uint8_t* st; int* pn; int* pp; int* path; int* lo;
    int* H; int* F; int* O; int* E; int* Q; int* N; long long* score; int* prev; int* cons; int* colof; int* first; int* cpos; int* mem;
    int rmax, bw_max;
};
// bytes of scratch a job of `total` bases (longest sequence maxl, band bw) needs
__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
__host__ __device__ inline int job_rmax(int total, int maxl) { const long long r = 3ll * maxl + 1024; return (int)(r < total + 16 ? r : total + 16); }
__host__ __device__ inline size_t scratch_bytes(int total, int maxl, int bw) {
    const size_t cap = (size_t)total + 16, rmax = (size_t)job_rmax(total, maxl);
    size_t b = 0;
    b += align256(cap) + 7 * align256(4 * cap) + 2 * align256(4 * cap * MAXIN);        // base st | nin col cov aligned order rank stack(edge shares below) | in_node in_w
    b += 3 * align256(4 * cap);                                                           // edge, colof/first, cpos/mem (reused)
    b += 2 * align256(8 * cap) + align256(4 * ((size_t)maxl + 16)) + align256(4 * rmax);  // pn pp (2 cap ints each) | path | lo
    b += 6 * align256(4 * rmax * (size_t)bw);                                             // H F O E Q N
    b += align256(8 * cap) + 2 * align256(4 * cap);                                       // score | prev cons
    return b + 4096;
}
__device__ inline void carve(Work& w, uint8_t* p, int total, int maxl, int bw) {
    const size_t cap = (size_t)total + 16, rmax = (size_t)job_rmax(total, maxl);
    auto take = [&](size_t bytes) { uint8_t* r = p; p += align256(bytes); return r; };
    w.G.cap = (int)cap;
    w.G.base = take(cap); w.st = take(cap);
    w.G.nin = (int*)take(4 * cap); w.G.col = (int*)take(4 * cap); w.G.cov = (int*)take(4 * cap); w.G.aligned = (int*)take(4 * cap); w.order = (int*)take(4 * cap); w.rank = (int*)take(4 * cap); w.stack = (int*)take(4 * cap);
    w.G.in_node = (int*)take(4 * cap * MAXIN); w.G.in_w = (int*)take(4 * cap * MAXIN);
    w.edge = (int*)take(4 * cap); w.colof = (int*)take(4 * cap); w.cpos = (int*)take(4 * cap); w.first = w.edge; w.mem = w.cpos;
    w.pn = (int*)take(8 * cap); w.pp = (int*)take(8 * cap); w.path = (int*)take(4 * ((size_t)maxl + 16)); w.lo = (int*)take(4 * rmax);
    const size_t cells = 4 * rmax * (size_t)bw;
    w.H = (int*)take(cells); w.F = (int*)take(cells); w.O = (int*)take(cells); w.E = (int*)take(cells); w.Q = (int*)take(cells); w.N = (int*)take(cells);
    w.score = (long long*)take(8 * cap); w.prev = (int*)take(4 * cap); w.cons = (int*)take(4 * cap);
    w.rmax = (int)rmax; w.bw_max = bw;
}

// ---- graph bookkeeping (one thread) ----
__device__ inline int g_node(Graph& G, uint8_t b, int col) {
    if (GN(G) >= G.cap) { *G.ovf_ = 1; return GN(G) - 1; }
    const int v = GN(G)++; G.base[v] = b; G.nin[v] = 0; G.col[v] = col; G.cov[v] = 0; G.aligned[v] = v; return v;
}
__device__ inline void g_edge(Graph& G, int from, int to) {
    for (int k = 0; k < G.nin[to]; ++k) if (G.in_node[to * MAXIN + k] == from) { G.in_w[to * MAXIN + k] += 1; return; }
    if (G.nin[to] >= MAXIN) { *G.ovf_ = 1; return; }
    G.in_node[to * MAXIN + G.nin[to]] = from; G.in_w[to * MAXIN + G.nin[to]] = 1; G.nin[to]++;
}
__device__ inline void g_topo(const Graph& G, Work& w) {
    const int n = GN(G); int cnt = 0;
    for (int i = 0; i < n; ++i) w.st[i] = 0;
    for (int s = 0; s < n; ++s) {
        if (w.st[s]) continue;
        int sp = 0; w.stack[0] = s; w.edge[0] = 0; w.st[s] = 1;
        while (sp >= 0) {
            const int v = w.stack[sp];
            if (w.edge[sp] < G.nin[v]) { const int p = G.in_node[v * MAXIN + w.edge[sp]++]; if (!w.st[p]) { w.st[p] = 1; ++sp; w.stack[sp] = p; w.edge[sp] = 0; } }
            else { w.rank[v] = cnt; w.order[cnt++] = v; --sp; }
        }
    }
}
__device__ inline void add_sequence(Graph& G, const uint8_t* seq, int L, const int* pn, const int* pp, int np, int* path) {
    for (int i = 0; i < L; ++i) path[i] = -1;
    for (int k = 0; k < np; ++k) {
        if (pp[k] < 0) continue;
        const int i = pp[k]; const int v = pn[k];
        if (v < 0) continue;
        if (G.base[v] == seq[i]) { path[i] = v; continue; }
        int found = -1; for (int a = G.aligned[v]; a != v; a = G.aligned[a]) if (G.base[a] == seq[i]) { found = a; break; }
        if (found < 0) { found = g_node(G, seq[i], G.col[v]); G.aligned[found] = G.aligned[v]; G.aligned[v] = found; }
        path[i] = found;
    }
    for (int i = 0; i < L; ++i) if (path[i] < 0) path[i] = g_node(G, seq[i], i);
    for (int i = 0; i < L; ++i) { G.cov[path[i]] += 1; if (i > 0) g_edge(G, path[i - 1], path[i]); }
}

__device__ __forceinline__ int cellv(const int* M, const int* lo, int r, int j, int bw, int dflt) {
    const int c = j - lo[r]; if (c < 0 || c >= bw) return dflt; return M[(size_t)r * bw + c];
}

// ---- the alignment: rows in topological order, the block computes one row at a time ----
struct Scores { int m, n, g, e, q, c; };
__device__ inline int align_block(Work& w, const uint8_t* seq, int L, Scores S, int W, unsigned long long* s_red /* shared, >= 2 * THREADS + 8 */) {
    const Graph& G = w.G; const int R = GN(G); const int bw = 2 * W + 1 < L ? 2 * W + 1 : L;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int K = (bw + nthr - 1) / nthr;                 // consecutive columns per thread
    unsigned long long best_key = 0;                      // (H << 42) | ((2^21 - 1 - r) << 21) | (2^21 - 1 - j): largest H, then smallest r, then smallest j
    int* s_pm = reinterpret_cast<int*>(s_red);            // [2][nthr] chunk maxima for the two prefix-max scans
    for (int r = 0; r < R; ++r) {
        const int v = w.order[r];
        int lo = G.col[v] + 1 - W; if (lo + bw - 1 > L) lo = L - bw + 1; if (lo < 1) lo = 1;
        if (tid == 0) w.lo[r] = lo;
        __syncthreads();                                   // lo[r] visible; the previous row's cells too
        int* H = w.H + (size_t)r * bw; int* F = w.F + (size_t)r * bw; int* O = w.O + (size_t)r * bw; int* E = w.E + (size_t)r * bw; int* Q = w.Q + (size_t)r * bw; int* N = w.N + (size_t)r * bw;
        const int c0 = tid * K, c1 = min(c0 + K, bw);
        const int nin = G.nin[v]; const uint8_t bv = G.base[v];
        // pass 1: Hn = max(0, D, F, O) per cell; chunk maxima of T_c = Hn_c - e c (and with the second piece's extension)
        int mxE = NEGV, mxQ = NEGV;
        for (int c = c0; c < c1; ++c) {
            const int j = lo + c; const int sc = seq[j - 1] == bv ? S.m : S.n;
            int d = NEGV, f = NEGV, o = NEGV;
            if (nin == 0) { d = sc; f = S.g; o = S.q; }
            for (int k = 0; k < nin; ++k) {
                const int pr = w.rank[G.in_node[v * MAXIN + k]];
                const int hd = j - 1 >= 1 ? cellv(w.H, w.lo, pr, j - 1, bw, 0) : 0; if (hd + sc > d) d = hd + sc;
                const int hv = cellv(w.H, w.lo, pr, j, bw, 0), fv = cellv(w.F, w.lo, pr, j, bw, NEGV), ov = cellv(w.O, w.lo, pr, j, bw, NEGV);
                int t = hv + S.g > fv + S.e ? hv + S.g : fv + S.e; if (t > f) f = t;
                t = hv + S.q > ov + S.c ? hv + S.q : ov + S.c; if (t > o) o = t;
            }
            int hn = 0; if (d > hn) hn = d; if (f > hn) hn = f; if (o > hn) hn = o;
            F[c] = f; O[c] = o; N[c] = hn;
            const int tE = hn - S.e * c, tQ = hn - S.c * c; if (tE > mxE) mxE = tE; if (tQ > mxQ) mxQ = tQ;
        }
        s_pm[tid] = mxE; s_pm[nthr + tid] = mxQ;
        __syncthreads();
        // exclusive prefix maximum over the chunks (the cell left of the band is a fresh start: Hn = 0 at c = -1)
        int pE = lo > 1 ? 0 - S.e * (-1) : NEGV, pQ = lo > 1 ? 0 - S.c * (-1) : NEGV;
        for (int t = 0; t < tid; ++t) { if (s_pm[t] > pE) pE = s_pm[t]; if (s_pm[nthr + t] > pQ) pQ = s_pm[nthr + t]; }
        // pass 2: E_c = g + e (c - 1) + max_{k < c} T_k, H = max(Hn, E, Q)
        for (int c = c0; c < c1; ++c) {
            const int j = lo + c;
            const int e = pE > NEGV / 2 ? pE + S.g + S.e * (c - 1) : NEGV, q = pQ > NEGV / 2 ? pQ + S.q + S.c * (c - 1) : NEGV;
            const int hn = N[c]; int h = hn; if (e > h) h = e; if (q > h) h = q;
            H[c] = h; E[c] = e; Q[c] = q;
            if (h > 0) { const unsigned long long key = ((unsigned long long)h << 42) | ((unsigned long long)(0x1fffff - r) << 21) | (unsigned long long)(0x1fffff - j); if (key > best_key) best_key = key; }
            const int tE = hn - S.e * c, tQ = hn - S.c * c; if (tE > pE) pE = tE; if (tQ > pQ) pQ = tQ;
        }
        __syncthreads();
    }
    // block maximum of the best cell
    s_red[tid] = best_key;
    __syncthreads();
    for (int o = nthr >> 1; o; o >>= 1) { if (tid < o && s_red[tid + o] > s_red[tid]) s_red[tid] = s_red[tid + o]; __syncthreads(); }
    const unsigned long long bk = s_red[0];
    __syncthreads();
    int np = 0;
    if (tid == 0 && bk) {
        const int best = (int)(bk >> 42); int r = 0x1fffff - (int)((bk >> 21) & 0x1fffff), j = 0x1fffff - (int)(bk & 0x1fffff);
        int state = 0, want = best; int* pair_node = w.pn; int* pair_pos = w.pp;
        while (r >= 0 && j >= 1) {
            const int v = w.order[r];
            if (state == 0) {
                const int h = want;
                if (h <= 0) break;
                const int sc = seq[j - 1] == G.base[v] ? S.m : S.n;
                int moved = 0;
                if (G.nin[v] == 0) { if (h == sc) { pair_node[np] = v; pair_pos[np] = j - 1; ++np; break; } }
                for (int k = 0; k < G.nin[v] && !moved; ++k) { const int pr = w.rank[G.in_node[v * MAXIN + k]]; const int hd = j - 1 >= 1 ? cellv(w.H, w.lo, pr, j - 1, bw, 0) : 0;
                    if (hd + sc == h) { pair_node[np] = v; pair_pos[np] = j - 1; ++np; r = pr; j = j - 1; want = hd; moved = 1; } }
                if (moved) continue;
                if (cellv(w.F, w.lo, r, j, bw, NEGV) == h) { state = 1; continue; }
                if (cellv(w.O, w.lo, r, j, bw, NEGV) == h) { state = 2; continue; }
                if (cellv(w.E, w.lo, r, j, bw, NEGV) == h) { state = 3; continue; }
                if (cellv(w.Q, w.lo, r, j, bw, NEGV) == h) { state = 4; continue; }
                break;
            } else if (state == 1 || state == 2) {
                const int* M = state == 1 ? w.F : w.O; const int open = state == 1 ? S.g : S.q, ext = state == 1 ? S.e : S.c;
                const int cur = cellv(M, w.lo, r, j, bw, NEGV);
                pair_node[np] = v; pair_pos[np] = -1; ++np;
                int moved = 0;
                if (G.nin[v] == 0) break;
                for (int k = 0; k < G.nin[v] && !moved; ++k) { const int pr = w.rank[G.in_node[v * MAXIN + k]];
                    if (cellv(w.H, w.lo, pr, j, bw, 0) + open == cur) { r = pr; state = 0; want = cur - open; moved = 1; } }
                for (int k = 0; k < G.nin[v] && !moved; ++k) { const int pr = w.rank[G.in_node[v * MAXIN + k]];
                    if (cellv(M, w.lo, pr, j, bw, NEGV) + ext == cur) { r = pr; moved = 1; } }
                if (!moved) break;
            } else {
                const int* M = state == 3 ? w.E : w.Q; const int open = state == 3 ? S.g : S.q, ext = state == 3 ? S.e : S.c;
                const int cur = cellv(M, w.lo, r, j, bw, NEGV);
                pair_node[np] = -1; pair_pos[np] = j - 1; ++np;
                const int nl = j - 1 >= 1 ? cellv(w.N, w.lo, r, j - 1, bw, 0) : 0;
                if (nl + open == cur) { state = 0; j = j - 1; want = nl; }
                else if (cellv(M, w.lo, r, j - 1, bw, NEGV) + ext == cur) { j = j - 1; }
                else break;
            }
        }
        for (int a = 0, b = np - 1; a < b; ++a, --b) { int t = pair_node[a]; pair_node[a] = pair_node[b]; pair_node[b] = t; t = pair_pos[a]; pair_pos[a] = pair_pos[b]; pair_pos[b] = t; }
    }
    return np;        // valid on thread 0
}

__device__ inline int consensus_path(Work& w, int min_cov) {
    const Graph& G = w.G; const int n = GN(G); int bestv = -1; long long bests = -1;
    for (int r = 0; r < n; ++r) {
        const int v = w.order[r]; w.prev[v] = -1; long long s = 0; int bwt = -1;
        for (int k = 0; k < G.nin[v]; ++k) { const int p = G.in_node[v * MAXIN + k], wt = G.in_w[v * MAXIN + k];
            if (wt > bwt || (wt == bwt && w.score[p] > w.score[w.prev[v]])) { bwt = wt; w.prev[v] = p; } }
        if (w.prev[v] >= 0) s = w.score[w.prev[v]] + bwt;
        w.score[v] = s;
        if (s > bests) { bests = s; bestv = v; }
    }
    int len = 0; for (int v = bestv; v >= 0; v = w.prev[v]) w.cons[len++] = v;
    for (int a = 0, b = len - 1; a < b; ++a, --b) { int t = w.cons[a]; w.cons[a] = w.cons[b]; w.cons[b] = t; }
    int a = 0, b = len; while (a < b && G.cov[w.cons[a]] < min_cov) ++a; while (b > a && G.cov[w.cons[b - 1]] < min_cov) --b;
    for (int i = a; i < b; ++i) w.cons[i - a] = w.cons[i];
    return b - a;
}

struct Params { const Job* jobs; uint32_t n_jobs; const uint8_t* seqs; const int* offs; uint8_t* out; int* out_len; uint8_t* scratch; size_t scratch_per_block; unsigned* next_job; };

__global__ void __launch_bounds__(THREADS) k_poa(const Params P) {
    __shared__ unsigned long long s_red[2 * THREADS + 8];
    __shared__ unsigned s_job; __shared__ int s_np, s_fail, s_gn, s_ovf;
    Work w;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_job = atomicAdd(P.next_job, 1u);
        __syncthreads();
        const unsigned ji = s_job; if (ji >= P.n_jobs) break;
        const Job jb = P.jobs[ji];
        const uint8_t* base = P.seqs + jb.seq_off; const int* offs = P.offs + jb.offs_off; const int n = (int)jb.n_seq;
        int total = 0, maxl = 0; for (int i = 0; i < n; ++i) { const int l = offs[i + 1] - offs[i]; total += l; if (l > maxl) maxl = l; }
        const int W = jb.band; const int bw = 2 * W + 1 < maxl ? 2 * W + 1 : (maxl > 0 ? maxl : 1);
        if (scratch_bytes(total, maxl, bw) > P.scratch_per_block) { if (threadIdx.x == 0) P.out_len[ji] = -2; continue; }
        carve(w, P.scratch + (size_t)blockIdx.x * P.scratch_per_block, total, maxl, bw);
        w.G.n_ = &s_gn; w.G.ovf_ = &s_ovf;
        const Scores S = { jb.m, jb.n, jb.g, jb.e, jb.q, jb.c };
        if (threadIdx.x == 0) { s_fail = 0; s_gn = 0; s_ovf = 0; }
        int* pa = w.pn + w.G.cap; int* pb = w.pp + w.G.cap;          // MSA mode: the two sequences' paths (upper halves of the pair buffers)
        for (int i = 0; i < n; ++i) {
            const uint8_t* s = base + offs[i]; const int L = offs[i + 1] - offs[i];
            __syncthreads();
            if (threadIdx.x == 0) { if (s_gn > 0 && L > 0) g_topo(w.G, w); if (s_gn > w.rmax) s_fail = 1; }
            __syncthreads();
            int np = 0;
            if (!s_fail && s_gn > 0 && L > 0) np = align_block(w, s, L, S, W, s_red);
            if (threadIdx.x == 0) {
                int* path = jb.mode == 1 ? (i == 0 ? pa : pb) : w.path;
                add_sequence(w.G, s, L, w.pn, w.pp, s_fail ? 0 : np, path);
                if (s_ovf) s_fail = 1;
            }
            __syncthreads();
        }
        (void)s_np;
        if (threadIdx.x == 0) {
            int ret = -1;
            if (!s_fail && !s_ovf) {
                g_topo(w.G, w);
                if (jb.mode == 0) {
                    const int len = s_gn ? consensus_path(w, jb.min_cov) : 0;
                    if (len <= (int)jb.out_cap) { for (int i = 0; i < len; ++i) P.out[jb.out_off + i] = w.G.base[w.cons[i]]; ret = len; }
                } else {
                    // columns = fused node sets, ordered topologically as a graph of their own
                    const int nn = s_gn; int ncol = 0;
                    for (int v = 0; v < nn; ++v) w.colof[v] = -1;
                    for (int v = 0; v < nn; ++v) { if (w.colof[v] >= 0) continue; w.colof[v] = ncol; for (int x = w.G.aligned[v]; x != v; x = w.G.aligned[x]) w.colof[x] = ncol; ++ncol; }
                    if (ncol <= (int)jb.out_cap) {
                        int* first = w.order; int* cpos = w.rank;           // the node order is no longer needed
                        for (int k = 0; k < ncol; ++k) first[k] = -1;
                        for (int v = nn - 1; v >= 0; --v) first[w.colof[v]] = v;
                        for (int k = 0; k < ncol; ++k) w.st[k] = 0;
                        int cnt = 0; int* stack = w.stack; int* mem = w.cpos; int* edge = w.edge;
                        for (int s0 = 0; s0 < ncol; ++s0) {
                            if (w.st[s0]) continue;
                            int sp = 0; stack[0] = s0; mem[0] = first[s0]; edge[0] = 0; w.st[s0] = 1;
                            while (sp >= 0) {
                                const int v = mem[sp];
                                if (edge[sp] < w.G.nin[v]) { const int pc = w.colof[w.G.in_node[v * MAXIN + edge[sp]++]]; if (!w.st[pc]) { w.st[pc] = 1; ++sp; stack[sp] = pc; mem[sp] = first[pc]; edge[sp] = 0; } }
                                else if (w.G.aligned[v] != first[stack[sp]]) { mem[sp] = w.G.aligned[v]; edge[sp] = 0; }
                                else { cpos[stack[sp]] = cnt++; --sp; }
                            }
                        }
                        uint8_t* ra = P.out + jb.out_off; uint8_t* rb = ra + jb.out_cap;
                        for (int k = 0; k < ncol; ++k) { ra[k] = 255; rb[k] = 255; }
                        const int la = offs[1] - offs[0], lb = n > 1 ? offs[2] - offs[1] : 0;
                        for (int i = 0; i < la; ++i) ra[cpos[w.colof[pa[i]]]] = base[offs[0] + i];
                        for (int i = 0; i < lb; ++i) rb[cpos[w.colof[pb[i]]]] = base[offs[1] + i];
                        ret = ncol;
                    }
                }
            }
            P.out_len[ji] = ret;
        }
    }
}

}  // namespace poa
