// Test infrastructure (not part of the product): compiles sniffles_b200/csrc/ingest_core.h with g++ as a one-lane "warp" so the
// DEFLATE decoder, the BAM record decoder and the CIGAR16 converter that the CUDA ingest kernels instantiate with 32 lanes can be
// checked against zlib / bamio on a machine without a GPU (tests/test_ingest_core.py).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../sniffles_b200/csrc/ingest_core.h"

extern "C" {

// raw DEFLATE stream in[ipos .. iend) -> out; returns INF_* (0 = ok); *out_len = bytes produced.  `in` must have 16 readable bytes behind iend.
int ingest_host_inflate(const uint8_t* in, uint64_t ipos, uint64_t iend, uint8_t* out, uint32_t out_cap, uint32_t* out_len) {
    ingest::WarpTables* T = (ingest::WarpTables*)malloc(sizeof(ingest::WarpTables));
    const int rc = ingest::inflate_stream<1>(in, ipos, iend, out, out_cap, T, 0, 1u, out_len);
    free(T);
    return rc;
}

uint64_t ingest_host_sizeof_rawrec(void) { return sizeof(ingest::RawRec); }

// walk the record chain of raw[ubeg .. uend): returns the number of records, fills recs[] (up to cap) with the parsed records
int64_t ingest_host_parse(const uint8_t* raw, uint64_t raw_len, uint64_t ubeg, uint64_t uend, ingest::RawRec* recs, uint64_t cap) {
    uint64_t off = ubeg; int64_t n = 0;
    while (off + 4 <= uend && off + 4 <= raw_len) {
        const uint32_t bs = ingest::ld32u(raw, off);
        if (bs < 32 || off + 4 + bs > raw_len) return -1;
        if ((uint64_t)n < cap) ingest::parse_record(raw, off + 4, bs, &recs[n]);
        ++n; off += 4 + (uint64_t)bs;
    }
    return n;
}

// BAM CIGAR words at raw[src ..] -> CIGAR16 words (out may be NULL to count); returns words written, -1 on an unknown op
int64_t ingest_host_c16(const uint8_t* raw, uint64_t src, uint32_t n, uint16_t* out, uint32_t evt_min, int64_t* reflen) {
    long long ref = 0; int bad = 0;
    const uint32_t k = ingest::c16_convert<1>(raw, src, n, out, evt_min, 0, &ref, &bad);
    *reflen = ref;
    return bad ? -1 : (int64_t)k;
}

}
