// DEFLATE decoder for ONE gzip member by ONE thread (RFC 1951 / 1952), written for the device (gz_inflate.cu: one
// thread per BGZF member, tens of thousands of members in flight) and compiled for the host as well so that the CPU
// tests exercise the same code.  Canonical-Huffman decoding by code length (counts per length + symbols in code
// order, as in zlib's contrib/puff): two tiny tables per code instead of multi-kilobyte lookup tables, so that a
// thread's tables fit in ~700 bytes of shared memory.  Table element e of a thread lives at tab[e * STRIDE]
// (device: the threads of a CTA interleave their tables, STRIDE = threads per CTA, so that the same element of
// neighbouring threads falls into neighbouring banks; host: STRIDE = 1).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GZI_HD __host__ __device__ __forceinline__
#else
#define GZI_HD inline
#endif

namespace gzi {

constexpr int MAXBITS = 15, MAXLCODES = 286, MAXDCODES = 30, FIXLCODES = 288;
constexpr int TAB_LCOUNT = 0, TAB_LSYM = 16, TAB_DCOUNT = 16 + 288, TAB_DSYM = 16 + 288 + 16, TAB_ELEMS = 16 + 288 + 16 + 32;

enum Status { OK = 0, ERR_HEADER = 1, ERR_BLOCK = 2, ERR_CODE = 3, ERR_DIST = 4, ERR_OVERRUN = 5, ERR_TRAILER = 6 };

struct BitIn {
    const uint8_t* p;    // next input byte
    const uint8_t* end;  // one past the member
    uint64_t buf;
    int cnt;
};

GZI_HD void refill(BitIn& b) {  // at least 32 bits afterwards (zeros behind the end of the member)
    while (b.cnt <= 32) {
        // four bytes at a time once the pointer is aligned; reading up to 3 bytes behind `end` is allowed (padding)
        if ((((uintptr_t)b.p) & 3u) == 0) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(b.p);
            b.buf |= (uint64_t)w << b.cnt;
            b.cnt += 32;
            b.p += 4;
        } else {
            b.buf |= (uint64_t)(*b.p) << b.cnt;
            b.cnt += 8;
            b.p += 1;
        }
    }
}
GZI_HD uint32_t take(BitIn& b, int n) {  // n <= 16, after refill
    const uint32_t v = (uint32_t)b.buf & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

template <int STRIDE>
GZI_HD int decode_sym(BitIn& b, const uint16_t* count, const uint16_t* symbol) {
    int code = 0, first = 0, index = 0;
    uint32_t bits = (uint32_t)b.buf;
    for (int len = 1; len <= MAXBITS; len++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[len * STRIDE];
        if (code - c < first) {
            b.buf >>= len;
            b.cnt -= len;
            return symbol[(index + (code - first)) * STRIDE];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// counts per length and symbols in code order from the code lengths; returns < 0 for an over-subscribed set,
// > 0 for an incomplete one (only acceptable for a single code of one bit, checked by the caller), 0 when complete
template <int STRIDE>
GZI_HD int construct(uint16_t* count, uint16_t* symbol, const uint8_t* length, int n) {
    for (int len = 0; len <= MAXBITS; len++) count[len * STRIDE] = 0;
    for (int s = 0; s < n; s++) count[length[s] * STRIDE]++;
    if (count[0] == n) return 0;  // no codes at all
    int left = 1;
    for (int len = 1; len <= MAXBITS; len++) {
        left <<= 1;
        left -= count[len * STRIDE];
        if (left < 0) return left;
    }
    uint16_t offs[MAXBITS + 1];
    offs[1] = 0;
    for (int len = 1; len < MAXBITS; len++) offs[len + 1] = (uint16_t)(offs[len] + count[len * STRIDE]);
    for (int s = 0; s < n; s++)
        if (length[s] != 0) symbol[(offs[length[s]]++) * STRIDE] = (uint16_t)s;
    return left;
}

// One gzip member src[0 .. n) -> dst[0 .. cap); tab: TAB_ELEMS elements of this thread at stride STRIDE.
// *produced = bytes written, *lines = '\n' among them.  Checks ISIZE (and that the member ends where it should); the
// CRC-32 of the member is not verified here (the host checks BGZF members it inflates itself; a corrupt member that
// still decodes to the right length would go through).
template <int STRIDE>
GZI_HD int inflate_member(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap, uint16_t* tab, uint32_t* produced, uint32_t* lines) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    *produced = 0;
    *lines = 0;
    // gzip header (RFC 1952): magic, CM = 8, FLG, MTIME(4), XFL, OS, then the optional fields
    if (n < 18 || src[0] != 0x1f || src[1] != 0x8b || src[2] != 8) return ERR_HEADER;
    const uint32_t flg = src[3];
    uint32_t pos = 10;
    if (flg & 4u) {
        const uint32_t xlen = src[10] | ((uint32_t)src[11] << 8);
        pos = 12 + xlen;
    }
    if (flg & 8u) {
        while (pos < n && src[pos]) pos++;
        pos++;
    }
    if (flg & 16u) {
        while (pos < n && src[pos]) pos++;
        pos++;
    }
    if (flg & 2u) pos += 2;
    if (pos + 8 > n) return ERR_HEADER;
    BitIn b = {src + pos, src + n, 0, 0};
    uint16_t* const lcount = tab + TAB_LCOUNT * STRIDE;
    uint16_t* const lsym = tab + TAB_LSYM * STRIDE;
    uint16_t* const dcount = tab + TAB_DCOUNT * STRIDE;
    uint16_t* const dsym = tab + TAB_DSYM * STRIDE;
    uint32_t out = 0, nl = 0;
    int last;
    do {
        refill(b);
        last = (int)take(b, 1);
        const uint32_t type = take(b, 2);
        if (type == 0) {  // stored
            take(b, b.cnt & 7);  // to the byte boundary
            refill(b);
            const uint32_t len = take(b, 16);
            refill(b);
            const uint32_t nlen = take(b, 16);
            if ((len ^ 0xFFFFu) != nlen) return ERR_BLOCK;
            // give the whole bytes in the bit buffer back
            const uint8_t* q = b.p - (b.cnt >> 3);
            b.buf = 0;
            b.cnt = 0;
            if (q + len > b.end || out + len > cap) return ERR_OVERRUN;
            for (uint32_t i = 0; i < len; i++) {
                const uint8_t c = q[i];
                dst[out + i] = c;
                nl += c == '\n';
            }
            out += len;
            b.p = q + len;
            continue;
        }
        if (type == 3) return ERR_BLOCK;
        uint8_t lengths[MAXLCODES + MAXDCODES + 2];
        if (type == 1) {  // fixed codes
            int s = 0;
            for (; s < 144; s++) lengths[s] = 8;
            for (; s < 256; s++) lengths[s] = 9;
            for (; s < 280; s++) lengths[s] = 7;
            for (; s < FIXLCODES; s++) lengths[s] = 8;
            construct<STRIDE>(lcount, lsym, lengths, FIXLCODES);
            for (s = 0; s < MAXDCODES; s++) lengths[s] = 5;
            construct<STRIDE>(dcount, dsym, lengths, MAXDCODES);
        } else {  // dynamic codes
            refill(b);
            const int nlen = (int)take(b, 5) + 257, ndist = (int)take(b, 5) + 1, ncode = (int)take(b, 4) + 4;
            if (nlen > MAXLCODES || ndist > MAXDCODES) return ERR_CODE;
            int idx = 0;
            for (; idx < ncode; idx++) {
                refill(b);
                lengths[order[idx]] = (uint8_t)take(b, 3);
            }
            for (; idx < 19; idx++) lengths[order[idx]] = 0;
            if (construct<STRIDE>(lcount, lsym, lengths, 19) != 0) return ERR_CODE;  // the code-length code must be complete
            idx = 0;
            while (idx < nlen + ndist) {
                refill(b);
                int sym = decode_sym<STRIDE>(b, lcount, lsym);
                if (sym < 0) return ERR_CODE;
                if (sym < 16) {
                    lengths[idx++] = (uint8_t)sym;
                } else {
                    int prev = 0, rep;
                    if (sym == 16) {
                        if (idx == 0) return ERR_CODE;
                        prev = lengths[idx - 1];
                        rep = 3 + (int)take(b, 2);
                    } else if (sym == 17) {
                        rep = 3 + (int)take(b, 3);
                    } else {
                        rep = 11 + (int)take(b, 7);
                    }
                    if (idx + rep > nlen + ndist) return ERR_CODE;
                    while (rep--) lengths[idx++] = (uint8_t)prev;
                }
            }
            if (lengths[256] == 0) return ERR_CODE;  // no end-of-block code
            int err = construct<STRIDE>(lcount, lsym, lengths, nlen);
            if (err < 0 || (err > 0 && nlen - lcount[0] != 1)) return ERR_CODE;
            err = construct<STRIDE>(dcount, dsym, lengths + nlen, ndist);
            if (err < 0 || (err > 0 && ndist - dcount[0] != 1)) return ERR_CODE;
        }
        for (;;) {  // the symbols of the block
            refill(b);
            int sym = decode_sym<STRIDE>(b, lcount, lsym);
            if (sym < 0) return ERR_CODE;
            if (sym < 256) {
                if (out >= cap) return ERR_OVERRUN;
                dst[out++] = (uint8_t)sym;
                nl += sym == '\n';
                continue;
            }
            if (sym == 256) break;
            sym -= 257;
            if (sym >= 29) return ERR_CODE;
            // length 3..258: codes 257..264 are 3..10, then groups of four codes per extra-bit count, 285 is 258
            uint32_t len;
            if (sym < 8) {
                len = 3u + (uint32_t)sym;
            } else if (sym == 28) {
                len = 258u;
            } else {
                const int e = (sym >> 2) - 1;
                len = ((4u + (uint32_t)(sym & 3)) << e) + 3u + take(b, e);
            }
            refill(b);
            const int ds = decode_sym<STRIDE>(b, dcount, dsym);
            if (ds < 0 || ds >= 30) return ERR_DIST;
            uint32_t dist;  // 1..32768: codes 0..3 are 1..4, then pairs of codes per extra-bit count
            if (ds < 4) {
                dist = 1u + (uint32_t)ds;
            } else {
                const int e = (ds >> 1) - 1;
                dist = ((2u + (uint32_t)(ds & 1)) << e) + 1u + take(b, e);
            }
            if (dist > out) return ERR_DIST;
            if (out + len > cap) return ERR_OVERRUN;
            for (uint32_t i = 0; i < len; i++) {
                const uint8_t c = dst[out - dist + i];
                dst[out + i] = c;
                nl += c == '\n';
            }
            out += len;
        }
    } while (!last);
    // trailer: CRC-32 (not verified here), ISIZE
    const uint8_t* tr = b.p - (b.cnt >> 3);
    if (tr + 8 > b.end) return ERR_TRAILER;
    const uint32_t isize = tr[4] | ((uint32_t)tr[5] << 8) | ((uint32_t)tr[6] << 16) | ((uint32_t)tr[7] << 24);
    if (isize != out) return ERR_TRAILER;
    *produced = out;
    *lines = nl;
    return OK;
}

}  // namespace gzi
