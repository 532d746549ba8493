// Developer aid (not run by the test suite): memory-safety fuzz of the one-lane host build of ingest_core.h under AddressSanitizer.
//   g++ -O1 -g -fsanitize=address,undefined -o /tmp/fuzz_inflate tests/native/fuzz_inflate.cpp -lz && /tmp/fuzz_inflate 20000
// Valid streams must decode like zlib; corrupted, truncated and random inputs must return an error or some output without touching a
// byte outside in[0 .. n + 16) and out[0 .. cap).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <zlib.h>
#include "../../sniffles_b200/csrc/ingest_core.h"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& d, int level, int strategy) {
    z_stream z; memset(&z, 0, sizeof z);
    deflateInit2(&z, level, Z_DEFLATED, -15, 8, strategy);
    std::vector<uint8_t> out(deflateBound(&z, d.size()) + 64);
    z.next_in = (Bytef*)d.data(); z.avail_in = (uInt)d.size(); z.next_out = out.data(); z.avail_out = (uInt)out.size();
    deflate(&z, Z_FINISH); out.resize(z.total_out); deflateEnd(&z); return out;
}
static int run(const std::vector<uint8_t>& comp, size_t lead, uint32_t cap, std::vector<uint8_t>& out, uint32_t* n) {
    // exact-size heap buffers: ASan catches any access outside [0, lead + comp + 16) and [0, cap)
    uint8_t* in = (uint8_t*)malloc(((lead + comp.size() + 16 + 3) & ~(size_t)3) + 4);
    memset(in, 0xa5, lead); memcpy(in + lead, comp.data(), comp.size()); memset(in + lead + comp.size(), 0, 16);
    uint8_t* o = (uint8_t*)malloc(cap ? cap : 1);
    ingest::WarpTables T;
    const int rc = ingest::inflate_stream<1>(in, lead, lead + comp.size(), o, cap, &T, 0, 1u, n);
    out.assign(o, o + (*n <= cap ? *n : cap));
    free(o); free(in); return rc;
}
int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 5000;
    long ok = 0, rejected = 0, garbage = 0;
    for (int it = 0; it < iters; ++it) {
        const size_t n = 1 + rnd() % 66000; const unsigned alpha = 1 + rnd() % 255;
        std::vector<uint8_t> d(n);
        const int kind = rnd() % 4;
        for (size_t i = 0; i < n; ++i) d[i] = kind == 0 ? (uint8_t)(rnd() % alpha) : kind == 1 ? "ACGT"[rnd() & 3] : kind == 2 ? (uint8_t)(i / 7) : (uint8_t)((rnd() % 40) + (i % 3 == 0 ? 100 : 0));
        static const int strategies[4] = { Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE };
        std::vector<uint8_t> comp = deflate_raw(d, (int)(rnd() % 10), strategies[rnd() & 3]);
        std::vector<uint8_t> out; uint32_t m = 0;
        int rc = run(comp, rnd() & 3, (uint32_t)n, out, &m);
        if (rc != 0 || out != d) { printf("MISMATCH on a valid stream (it %d, rc %d)\n", it, rc); return 1; }
        ++ok;
        // corruptions: bit flips, truncation, smaller output cap, random bytes
        std::vector<uint8_t> c2 = comp;
        const int mode = rnd() % 4;
        if (mode == 0) for (int k = 0; k < 1 + (int)(rnd() % 4); ++k) c2[rnd() % c2.size()] ^= (uint8_t)(1u << (rnd() & 7));
        else if (mode == 1) c2.resize(rnd() % c2.size());
        else if (mode == 2) for (auto& b : c2) b = (uint8_t)rnd();
        uint32_t cap = mode == 3 ? (uint32_t)(rnd() % n) : (uint32_t)n;
        rc = run(c2, rnd() & 3, cap, out, &m);
        if (rc != 0) ++rejected; else ++garbage;
    }
    printf("valid %ld ok; corrupted: %ld rejected, %ld decoded to something\n", ok, rejected, garbage);
    return 0;
}
