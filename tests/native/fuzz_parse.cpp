// Developer aid (not run by the test suite): AddressSanitizer fuzz of the BAM record decoder and the CIGAR16 converter (one-lane host build of ingest_core.h).
//   g++ -O1 -g -fsanitize=address,undefined -o /tmp/fuzz_parse tests/native/fuzz_parse.cpp && /tmp/fuzz_parse
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../sniffles_b200/csrc/ingest_core.h"
static uint64_t s = 1234567ull; static uint32_t rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); }
int main() {
    long okc = 0, bad = 0;
    for (int it = 0; it < 300000; ++it) {
        // a plausible record, then mutations
        std::vector<uint8_t> b;
        auto put32 = [&](uint32_t v) { for (int k = 0; k < 4; ++k) b.push_back((uint8_t)(v >> (8 * k))); };
        auto put16 = [&](uint32_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); };
        const uint32_t l_rn = 1 + rnd() % 20, n_cig = rnd() % 6, l_seq = rnd() % 40;
        put32(0); put32(1000); b.push_back((uint8_t)l_rn); b.push_back(60); put16(4681); put16(n_cig); put16(0); put32(l_seq); put32(0xffffffffu); put32(0xffffffffu); put32(0);
        for (uint32_t k = 0; k + 1 < l_rn; ++k) b.push_back('a'); b.push_back(0);
        for (uint32_t k = 0; k < n_cig; ++k) put32(((1 + rnd() % 3000) << 4) | (rnd() % 9));
        for (uint32_t k = 0; k < (l_seq + 1) / 2 + l_seq; ++k) b.push_back((uint8_t)rnd());
        const char* tags[] = { "NMi", "HPC", "PSi", "SAZ", "XAA", "XfF", "MLB", "CGB", "XHH" };
        for (int t = 0; t < (int)(rnd() % 6); ++t) {
            const char* tg = tags[rnd() % 9]; b.push_back(tg[0]); b.push_back(tg[1]);
            char ty = tg[2] == 'F' ? 'f' : tg[2]; b.push_back((uint8_t)ty);
            if (ty == 'i' || ty == 'f') put32(rnd()); else if (ty == 'C' || ty == 'A') b.push_back((uint8_t)rnd());
            else if (ty == 'Z' || ty == 'H') { for (int k = 0; k < (int)(rnd() % 30); ++k) b.push_back('0' + rnd() % 10); b.push_back(0); }
            else if (ty == 'B') { const char sub[] = "cCsSiIf"; char sb = sub[rnd() % 7]; b.push_back((uint8_t)sb); uint32_t cnt = rnd() % 12; put32(cnt); int esz = (sb == 'c' || sb == 'C') ? 1 : (sb == 's' || sb == 'S') ? 2 : 4; for (uint32_t k = 0; k < cnt * esz; ++k) b.push_back((uint8_t)rnd()); }
        }
        const int muts = rnd() % 4;
        for (int m = 0; m < muts; ++m) b[rnd() % b.size()] = (uint8_t)rnd();
        uint32_t bs = (uint32_t)b.size(); if (rnd() % 5 == 0) bs = rnd() % (bs + 1);          // truncated body
        uint8_t* raw = (uint8_t*)malloc(bs + 64); memcpy(raw, b.data(), bs); memset(raw + bs, 0, 64);     // the device buffer has 64 bytes of slack behind the stream
        ingest::RawRec r; ingest::parse_record(raw, 0, bs, &r);
        if (r.status == ingest::ST_OK) {
            ++okc;
            if (r.cig_src + 4ull * r.n_cig > bs || r.seq_src + (uint64_t)((r.l_seq + 1) / 2) > bs || 32ull + r.l_qname > bs || (r.sa_len && r.sa_src + r.sa_len > bs)) { printf("accepted record points outside its body (it %d)\n", it); return 1; }
            long long ref = 0; int badop = 0;
            ingest::c16_convert<1>(raw, r.cig_src, r.n_cig, nullptr, 11, 0, &ref, &badop);
        } else ++bad;
        free(raw);
    }
    printf("parse fuzz: %ld accepted, %ld rejected, no out-of-bounds access\n", okc, bad);
}
