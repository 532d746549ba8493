/*
 * poa_oracle.c — CPU restatement of the partial-order-alignment consensus behind the reference's LocalAsm
 * (/root/reference/src/sniffles/local_asm.py:254-304).  TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).
 *
 * PARITY UNPINNED: the reference delegates this step to pyspoa (`from spoa import poa`, requirement pyspoa>=0.2.1,
 * setup.cfg:27; wraps rvaser/spoa), which is not in this image and which no reference test exercises.  What is restated
 * here is spoa's published algorithm at the level its documentation and the call sites fix:
 *   - sequences are added one by one to a partial order graph by a LOCAL (Smith-Waterman) alignment of the sequence
 *     against the graph (algorithm=0 at local_asm.py:287,289), scores m / n and a gap of length k costing
 *     max(g + (k-1) e, q + (k-1) c) (two affine pieces = spoa's convex mode; pyspoa defaults 5,-4,-8,-6,-10,-4 for the
 *     read pile-up, the size-class table of local_asm.py:26-73 for consensus-vs-reference);
 *   - aligned bases are fused into the graph (same base: same node, other base: a node aligned to it), unaligned ones
 *     become new nodes; every step of a sequence adds weight 1 to the edge it walks;
 *   - the consensus is the heaviest path (per node the heaviest incoming edge, ties to the predecessor with the larger
 *     path score), cut at both ends while fewer than min_coverage sequences pass through (min_coverage = round(n / 2),
 *     local_asm.py:285);
 *   - genmsa: one row per sequence over the graph's nodes in topological order, '-' where a sequence does not pass.
 * Tie-breaking inside spoa's SIMD engine is NOT reproduced (it cannot be pinned here); the stated tolerance against the real
 * library is <= 2 % normalised edit distance of the consensus and identical solve_ins / solve_del decisions.
 * The CUDA kernel (sniffles_b200/csrc/poa.cuh) implements exactly THIS algorithm, cell for cell, and is tested for equality.
 *
 * Alignment recurrences (node v in topological order, read position j, preds(v) = sources of v's incoming edges,
 * the virtual start has H = 0):
 *   D  = max_p H[p][j-1] + s(base[v], read[j])
 *   F  = max_p max(H[p][j] + g, F[p][j] + e)          O = max_p max(H[p][j] + q, O[p][j] + c)      (graph base, no read base)
 *   Hn = max(0, D, F, O)                              (the cell reached without a horizontal gap)
 *   E  = max(Hn[v][j-1] + g, E[v][j-1] + e)           Q = max(Hn[v][j-1] + q, Q[v][j-1] + c)       (read base, no graph base)
 *   H  = max(Hn, E, Q)
 * (a horizontal gap opens from Hn, never from another horizontal gap: switching between the two affine pieces inside one gap never
 * scores better than the better piece alone, and this way a row's E / Q are plain prefix maxima — what the CUDA kernel computes)
 * restricted to a band |j - col(v)| <= W around the column a node was created at (cells outside the band are fresh starts:
 * H = 0, gaps closed).  Best cell: largest H, then smallest topological index, then smallest j.  The traceback re-derives the
 * predecessor of a cell from the stored matrices in the fixed order D, F, O, E, Q (first predecessor in edge order that attains it).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXIN 8
#define NEG (-(1 << 29))

typedef struct {
    int n, cap;
    uint8_t* base; int* nin; int* in_node; int* in_w; int* col; int* cov; int* aligned;   /* aligned: ring of nodes fused into one column (next index, self when alone) */
    int overflow;
} graph_t;

typedef struct { int m, n, g, e, q, c; } scores_t;

static void g_init(graph_t* G, int cap) {
    memset(G, 0, sizeof *G); G->cap = cap;
    G->base = (uint8_t*)malloc(cap); G->nin = (int*)calloc(cap, 4); G->in_node = (int*)malloc((size_t)cap * MAXIN * 4); G->in_w = (int*)malloc((size_t)cap * MAXIN * 4);
    G->col = (int*)malloc((size_t)cap * 4); G->cov = (int*)calloc(cap, 4); G->aligned = (int*)malloc((size_t)cap * 4);
}
static void g_free(graph_t* G) { free(G->base); free(G->nin); free(G->in_node); free(G->in_w); free(G->col); free(G->cov); free(G->aligned); }
static int g_node(graph_t* G, uint8_t b, int col) {
    if (G->n >= G->cap) { G->overflow = 1; return G->n - 1; }
    const int v = G->n++; G->base[v] = b; G->nin[v] = 0; G->col[v] = col; G->cov[v] = 0; G->aligned[v] = v; return v;
}
static void g_edge(graph_t* G, int from, int to) {
    for (int k = 0; k < G->nin[to]; ++k) if (G->in_node[to * MAXIN + k] == from) { G->in_w[to * MAXIN + k] += 1; return; }
    if (G->nin[to] >= MAXIN) { G->overflow = 1; return; }
    G->in_node[to * MAXIN + G->nin[to]] = from; G->in_w[to * MAXIN + G->nin[to]] = 1; G->nin[to]++;
}
/* topological order from the in-edge lists alone: post-order of a depth-first walk against the edges, nodes tried in id order */
static void g_topo(const graph_t* G, int* order, int* rank) {
    const int n = G->n; uint8_t* st = (uint8_t*)calloc(n, 1); int* stack = (int*)malloc((size_t)n * 4); int* edge = (int*)malloc((size_t)n * 4); int cnt = 0;
    for (int s = 0; s < n; ++s) {
        if (st[s]) continue;
        int sp = 0; stack[0] = s; edge[0] = 0; st[s] = 1;
        while (sp >= 0) {
            const int v = stack[sp];
            if (edge[sp] < G->nin[v]) { const int p = G->in_node[v * MAXIN + edge[sp]++]; if (!st[p]) { st[p] = 1; ++sp; stack[sp] = p; edge[sp] = 0; } }
            else { rank[v] = cnt; order[cnt++] = v; --sp; }
        }
    }
    free(st); free(stack); free(edge);
}

typedef struct { int* H; int* F; int* O; int* E; int* Q; int* N; int* lo; long* off; } dp_t;     /* row r covers columns lo[r] .. lo[r] + bw - 1 (read positions 1-based) */

static inline int cell(const int* M, const dp_t* D, int r, int j, int bw, int dflt) {
    const int c = j - D->lo[r]; if (c < 0 || c >= bw) return dflt; return M[D->off[r] + c];
}

/* local alignment of seq[0..L) against the graph; returns pairs (node or -1, read index or -1) in forward order */
static int align(const graph_t* G, const int* order, const int* rank, const uint8_t* seq, int L, scores_t S, int W, int* pair_node, int* pair_pos) {
    const int R = G->n; const int bw = 2 * W + 1 < L ? 2 * W + 1 : L;
    dp_t D; D.lo = (int*)malloc((size_t)R * 4); D.off = (long*)malloc((size_t)R * 8);
    const size_t cells = (size_t)R * bw;
    D.H = (int*)malloc(cells * 4); D.F = (int*)malloc(cells * 4); D.O = (int*)malloc(cells * 4); D.E = (int*)malloc(cells * 4); D.Q = (int*)malloc(cells * 4); D.N = (int*)malloc(cells * 4);
    int best = 0, br = -1, bj = -1;
    for (int r = 0; r < R; ++r) {
        const int v = order[r];
        int lo = G->col[v] + 1 - W; if (lo + bw - 1 > L) lo = L - bw + 1; if (lo < 1) lo = 1;
        D.lo[r] = lo; D.off[r] = (long)r * bw;
        int* H = D.H + D.off[r]; int* F = D.F + D.off[r]; int* O = D.O + D.off[r]; int* E = D.E + D.off[r]; int* Q = D.Q + D.off[r]; int* N = D.N + D.off[r];
        for (int c = 0; c < bw; ++c) {
            const int j = lo + c; const int sc = seq[j - 1] == G->base[v] ? S.m : S.n;
            int d = NEG, f = NEG, o = NEG;
            if (G->nin[v] == 0) { d = sc; f = S.g; o = S.q; }
            for (int k = 0; k < G->nin[v]; ++k) {
                const int pr = rank[G->in_node[v * MAXIN + k]];
                const int hd = j - 1 >= 1 ? cell(D.H, &D, pr, j - 1, bw, 0) : 0; if (hd + sc > d) d = hd + sc;
                const int hv = cell(D.H, &D, pr, j, bw, 0), fv = cell(D.F, &D, pr, j, bw, NEG), ov = cell(D.O, &D, pr, j, bw, NEG);
                int t = hv + S.g > fv + S.e ? hv + S.g : fv + S.e; if (t > f) f = t;
                t = hv + S.q > ov + S.c ? hv + S.q : ov + S.c; if (t > o) o = t;
            }
            int e = NEG, q = NEG;
            if (c > 0) { e = N[c - 1] + S.g > E[c - 1] + S.e ? N[c - 1] + S.g : E[c - 1] + S.e; q = N[c - 1] + S.q > Q[c - 1] + S.c ? N[c - 1] + S.q : Q[c - 1] + S.c; }
            else if (j - 1 >= 1) { e = S.g; q = S.q; }                      /* the cell to the left is outside the band: a fresh start (Hn = 0) */
            int hn = 0; if (d > hn) hn = d; if (f > hn) hn = f; if (o > hn) hn = o;
            int h = hn; if (e > h) h = e; if (q > h) h = q;
            H[c] = h; F[c] = f; O[c] = o; E[c] = e; Q[c] = q; N[c] = hn;
            if (h > best) { best = h; br = r; bj = j; }
        }
    }
    /* traceback: `want` is the value the current H-state cell was entered with (a horizontal gap opens from Hn, which can be below H) */
    int np = 0;
    if (br >= 0) {
        int r = br, j = bj, state = 0, want = best;                         /* 0 H, 1 F, 2 O, 3 E, 4 Q */
        while (r >= 0 && j >= 1) {
            const int v = order[r];
            if (state == 0) {
                const int h = want;
                if (h <= 0) break;
                const int sc = seq[j - 1] == G->base[v] ? S.m : S.n;
                int moved = 0;
                if (G->nin[v] == 0) { if (h == sc) { pair_node[np] = v; pair_pos[np] = j - 1; ++np; break; } }
                for (int k = 0; k < G->nin[v] && !moved; ++k) { const int pr = rank[G->in_node[v * MAXIN + k]]; const int hd = j - 1 >= 1 ? cell(D.H, &D, pr, j - 1, bw, 0) : 0;
                    if (hd + sc == h) { pair_node[np] = v; pair_pos[np] = j - 1; ++np; r = pr; j = j - 1; want = hd; moved = 1; } }
                if (moved) continue;
                if (cell(D.F, &D, r, j, bw, NEG) == h) { state = 1; continue; }
                if (cell(D.O, &D, r, j, bw, NEG) == h) { state = 2; continue; }
                if (cell(D.E, &D, r, j, bw, NEG) == h) { state = 3; continue; }
                if (cell(D.Q, &D, r, j, bw, NEG) == h) { state = 4; continue; }
                break;
            } else if (state == 1 || state == 2) {                        /* graph base against a gap in the read */
                const int* M = state == 1 ? D.F : D.O; const int open = state == 1 ? S.g : S.q, ext = state == 1 ? S.e : S.c;
                const int cur = cell(M, &D, r, j, bw, NEG);
                pair_node[np] = v; pair_pos[np] = -1; ++np;
                int moved = 0;
                if (G->nin[v] == 0) break;
                for (int k = 0; k < G->nin[v] && !moved; ++k) { const int pr = rank[G->in_node[v * MAXIN + k]];
                    if (cell(D.H, &D, pr, j, bw, 0) + open == cur) { r = pr; state = 0; want = cur - open; moved = 1; } }
                for (int k = 0; k < G->nin[v] && !moved; ++k) { const int pr = rank[G->in_node[v * MAXIN + k]];
                    if (cell(M, &D, pr, j, bw, NEG) + ext == cur) { r = pr; moved = 1; } }
                if (!moved) break;
            } else {                                                       /* read base against a gap in the graph */
                const int* M = state == 3 ? D.E : D.Q; const int open = state == 3 ? S.g : S.q, ext = state == 3 ? S.e : S.c;
                const int cur = cell(M, &D, r, j, bw, NEG);
                pair_node[np] = -1; pair_pos[np] = j - 1; ++np;
                const int nl = j - 1 >= 1 ? cell(D.N, &D, r, j - 1, bw, 0) : 0;
                if (nl + open == cur) { state = 0; j = j - 1; want = nl; }
                else if (cell(M, &D, r, j - 1, bw, NEG) + ext == cur) { j = j - 1; }
                else break;
            }
        }
    }
    /* reverse */
    for (int a = 0, b = np - 1; a < b; ++a, --b) { int t = pair_node[a]; pair_node[a] = pair_node[b]; pair_node[b] = t; t = pair_pos[a]; pair_pos[a] = pair_pos[b]; pair_pos[b] = t; }
    free(D.lo); free(D.off); free(D.H); free(D.F); free(D.O); free(D.E); free(D.Q); free(D.N);
    return np;
}

/* fuse an aligned sequence into the graph; path[i] = node of read base i */
static void add_sequence(graph_t* G, const uint8_t* seq, int L, const int* pn, const int* pp, int np, int* path) {
    for (int i = 0; i < L; ++i) path[i] = -1;
    for (int k = 0; k < np; ++k) {
        if (pp[k] < 0) continue;
        const int i = pp[k]; const int v = pn[k];
        if (v < 0) continue;
        if (G->base[v] == seq[i]) { path[i] = v; continue; }
        int found = -1; for (int a = G->aligned[v]; a != v; a = G->aligned[a]) if (G->base[a] == seq[i]) { found = a; break; }
        if (found < 0) { found = g_node(G, seq[i], G->col[v]); G->aligned[found] = G->aligned[v]; G->aligned[v] = found; }
        path[i] = found;
    }
    for (int i = 0; i < L; ++i) if (path[i] < 0) path[i] = g_node(G, seq[i], i);
    for (int i = 0; i < L; ++i) { G->cov[path[i]] += 1; if (i > 0) g_edge(G, path[i - 1], path[i]); }
}

static int consensus_path(const graph_t* G, const int* order, int min_cov, int* out) {
    const int n = G->n; long* score = (long*)calloc(n, 8); int* prev = (int*)malloc((size_t)n * 4); int bestv = -1; long bests = -1;
    for (int r = 0; r < n; ++r) {
        const int v = order[r]; prev[v] = -1; long s = 0; int bw = -1;
        for (int k = 0; k < G->nin[v]; ++k) { const int p = G->in_node[v * MAXIN + k], w = G->in_w[v * MAXIN + k];
            if (w > bw || (w == bw && score[p] > score[prev[v]])) { bw = w; prev[v] = p; } }
        if (prev[v] >= 0) s = score[prev[v]] + bw;
        score[v] = s;
        if (s > bests) { bests = s; bestv = v; }
    }
    int len = 0; for (int v = bestv; v >= 0; v = prev[v]) out[len++] = v;
    for (int a = 0, b = len - 1; a < b; ++a, --b) { int t = out[a]; out[a] = out[b]; out[b] = t; }
    int a = 0, b = len; while (a < b && G->cov[out[a]] < min_cov) ++a; while (b > a && G->cov[out[b - 1]] < min_cov) --b;
    for (int i = a; i < b; ++i) out[i - a] = out[i];
    free(score); free(prev);
    return b - a;
}

/* ---- exported ---- */
/* consensus of n sequences (codes 0..255, concatenated, offs[n+1]); returns its length (<= out_cap) or -1 on overflow */
int po_consensus(const uint8_t* seqs, const int* offs, int n, int min_cov, int m, int nn, int g, int e, int q, int c, int W, uint8_t* out, int out_cap) {
    scores_t S = { m, nn, g, e, q, c };
    int total = 0, maxl = 0; for (int i = 0; i < n; ++i) { const int l = offs[i + 1] - offs[i]; total += l; if (l > maxl) maxl = l; }
    graph_t G; g_init(&G, total + 16);
    int* order = (int*)malloc((size_t)(total + 16) * 4); int* rank = (int*)malloc((size_t)(total + 16) * 4);
    int* pn = (int*)malloc((size_t)(2 * (total + 16)) * 4); int* pp = (int*)malloc((size_t)(2 * (total + 16)) * 4); int* path = (int*)malloc((size_t)(maxl + 1) * 4);
    for (int i = 0; i < n; ++i) {
        const uint8_t* s = seqs + offs[i]; const int L = offs[i + 1] - offs[i];
        int np = 0;
        if (G.n > 0 && L > 0) { g_topo(&G, order, rank); np = align(&G, order, rank, s, L, S, W, pn, pp); }
        add_sequence(&G, s, L, pn, pp, np, path);
    }
    g_topo(&G, order, rank);
    int* cons = (int*)malloc((size_t)(G.n + 1) * 4);
    const int len = G.n ? consensus_path(&G, order, min_cov, cons) : 0;
    int ret = len <= out_cap && !G.overflow ? len : -1;
    if (ret >= 0) for (int i = 0; i < len; ++i) out[i] = G.base[cons[i]];
    free(cons); free(order); free(rank); free(pn); free(pp); free(path); g_free(&G);
    return ret;
}
/* two-row MSA of (a, b): a is the backbone, b is aligned locally against it (the genmsa of poa([consensus, ref]) at local_asm.py:289-291).
 * Columns: nodes in topological order; rows hold the base code or 255 ('-').  Returns the number of columns or -1. */
int po_pair_msa(const uint8_t* a, int la, const uint8_t* b, int lb, int m, int nn, int g, int e, int q, int c, int W, uint8_t* row_a, uint8_t* row_b, int cap) {
    scores_t S = { m, nn, g, e, q, c };
    graph_t G; g_init(&G, la + lb + 16);
    int* order = (int*)malloc((size_t)(la + lb + 16) * 4); int* rank = (int*)malloc((size_t)(la + lb + 16) * 4);
    int* pn = (int*)calloc((size_t)(2 * (la + lb + 16)), 4); int* pp = (int*)calloc((size_t)(2 * (la + lb + 16)), 4);
    int* pa = (int*)malloc((size_t)(la + 1) * 4); int* pb = (int*)malloc((size_t)(lb + 1) * 4);
    add_sequence(&G, a, la, pn, pp, 0, pa);
    int np = 0;
    if (la > 0 && lb > 0) { g_topo(&G, order, rank); np = align(&G, order, rank, b, lb, S, W, pn, pp); }
    add_sequence(&G, b, lb, pn, pp, np, pb);
    g_topo(&G, order, rank);
    /* columns: fused nodes share a column; the columns are ordered topologically as a graph of their own (an edge between two nodes is an
     * edge between their columns), so that every sequence reads left to right */
    const int n = G.n; int* colof = (int*)malloc((size_t)n * 4); int ncol = 0;
    for (int v = 0; v < n; ++v) colof[v] = -1;
    for (int v = 0; v < n; ++v) { if (colof[v] >= 0) continue; colof[v] = ncol; for (int x = G.aligned[v]; x != v; x = G.aligned[x]) colof[x] = ncol; ++ncol; }
    int ret = 0;
    if (ncol > cap || G.overflow) ret = -1;
    else {
        /* members of each column, then a depth-first post-order against the edges (columns tried in the order of their first node id) */
        int* first = (int*)malloc((size_t)ncol * 4); int* cpos = (int*)malloc((size_t)ncol * 4);
        for (int k = 0; k < ncol; ++k) first[k] = -1;
        for (int v = n - 1; v >= 0; --v) first[colof[v]] = v;
        uint8_t* st = (uint8_t*)calloc(ncol, 1); int* stack = (int*)malloc((size_t)ncol * 4); int* mem = (int*)malloc((size_t)ncol * 4); int* edge = (int*)malloc((size_t)ncol * 4); int cnt = 0;
        for (int s0 = 0; s0 < ncol; ++s0) {
            if (st[s0]) continue;
            int sp = 0; stack[0] = s0; mem[0] = first[s0]; edge[0] = 0; st[s0] = 1;
            while (sp >= 0) {
                const int v = mem[sp];                                    /* current member of the column on top of the stack */
                if (edge[sp] < G.nin[v]) { const int pc = colof[G.in_node[v * MAXIN + edge[sp]++]]; if (!st[pc]) { st[pc] = 1; ++sp; stack[sp] = pc; mem[sp] = first[pc]; edge[sp] = 0; } }
                else if (G.aligned[v] != first[stack[sp]]) { mem[sp] = G.aligned[v]; edge[sp] = 0; }       /* next member of the ring */
                else { cpos[stack[sp]] = cnt++; --sp; }
            }
        }
        memset(row_a, 255, ncol); memset(row_b, 255, ncol);
        for (int i = 0; i < la; ++i) row_a[cpos[colof[pa[i]]]] = a[i];
        for (int i = 0; i < lb; ++i) row_b[cpos[colof[pb[i]]]] = b[i];
        ret = ncol;
        free(first); free(cpos); free(st); free(stack); free(mem); free(edge);
    }
    free(colof); free(order); free(rank); free(pn); free(pp); free(pa); free(pb); g_free(&G);
    return ret;
}
