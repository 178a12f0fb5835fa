// DEFLATE decoder for ONE gzip member by ONE group of LANES cooperating threads (RFC 1951 / 1952).  On the device a
// group is a warp (gz_inflate.cu: one warp per BGZF member, every member of a batch in flight at once); the host twin
// runs the same code with LANES = 1 for the CPU tests.
//
// Every lane of the group runs the whole bit-level decode REDUNDANTLY on the same bits (bit buffer, code tables and
// control flow are warp-uniform: no divergence, no shuffles, shared-memory reads are broadcasts); what the lanes split
// is the byte traffic: literals are collected one per lane and leave as one store of LANES bytes, match and stored-block
// copies are strided over the lanes with their loads in flight together.  (The first version ran one THREAD per member:
// 32 different decoder states per warp serialise completely and a batch is only a few hundred warps - 0.4 GB/s.)
//
// Huffman codes: canonical decoding by code length (counts per length + symbols in code order, as in zlib's
// contrib/puff) is kept as the slow path for long codes and for validation; in front of it a direct table indexed by
// the next LBITS / DBITS input bits answers codes of up to that length in one lookup.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GZI_HD __host__ __device__ __forceinline__
#else
#define GZI_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define GZI_SYNC() __syncwarp()
#define GZI_UNROLL _Pragma("unroll")
#else
#define GZI_SYNC() ((void)0)
#define GZI_UNROLL
#endif

namespace gzi {

constexpr int MAXBITS = 15, MAXLCODES = 286, MAXDCODES = 30, FIXLCODES = 288;
constexpr int LBITS = 10, DBITS = 8;  // index bits of the direct tables

enum Status { OK = 0, ERR_HEADER = 1, ERR_BLOCK = 2, ERR_CODE = 3, ERR_DIST = 4, ERR_OVERRUN = 5, ERR_TRAILER = 6 };

struct Tables {  // per group (device: shared memory, 3.6 KB per warp)
    uint16_t llut[1 << LBITS];  // (symbol << 4) | code length; 0: the code is longer than LBITS (or invalid)
    uint16_t dlut[1 << DBITS];
    uint16_t lcount[MAXBITS + 1], dcount[MAXBITS + 1];  // codes per length
    uint16_t lsym[FIXLCODES], dsym[32];                 // symbols in code order
    uint8_t lengths[MAXLCODES + MAXDCODES + 4];
};

struct BitIn {
    const uint8_t* p;    // next input word (4-byte aligned)
    const uint8_t* end;  // one past the member
    uint64_t buf;
    int cnt;
};

// Input is read in aligned 32-bit words (the compressed buffer is padded, and reading the bytes in front of a member
// is harmless): starting at byte address q means loading the word that holds it and dropping the bytes in front.
GZI_HD void start(BitIn& b, const uint8_t* q) {
    const uint32_t mis = (uint32_t)((uintptr_t)q & 3u);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(q - mis);
    b.buf = (uint64_t)(w >> (8u * mis));
    b.cnt = 32 - 8 * (int)mis;
    b.p = q - mis + 4;
}
GZI_HD const uint8_t* byte_pos(const BitIn& b) { return b.p - (b.cnt >> 3); }  // next unread byte (at a byte boundary)
GZI_HD void refill(BitIn& b) {  // at least 33 bits afterwards (whatever lies behind the member: zeros or the next one)
    if (b.cnt <= 32) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(b.p);
        b.buf |= (uint64_t)w << b.cnt;
        b.cnt += 32;
        b.p += 4;
    }
}
GZI_HD uint32_t take(BitIn& b, int n) {  // n <= 16, after refill
    const uint32_t v = (uint32_t)b.buf & ((1u << n) - 1u);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

// canonical decode, one bit at a time (codes the direct table does not hold)
GZI_HD int decode_slow(BitIn& b, const uint16_t* count, const uint16_t* symbol) {
    int code = 0, first = 0, index = 0;
    uint32_t bits = (uint32_t)b.buf;
    for (int len = 1; len <= MAXBITS; len++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int c = count[len];
        if (code - c < first) {
            b.buf >>= len;
            b.cnt -= len;
            return symbol[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

template <int BITS>
GZI_HD int decode_sym(BitIn& b, const uint16_t* lut, const uint16_t* count, const uint16_t* symbol) {
    const uint32_t e = lut[(uint32_t)b.buf & ((1u << BITS) - 1u)];
    if (e) {
        const int n = (int)(e & 15u);
        b.buf >>= n;
        b.cnt -= n;
        return (int)(e >> 4);
    }
    return decode_slow(b, count, symbol);
}

GZI_HD uint32_t bit_reverse(uint32_t v, int n) {  // the low n bits of v, reversed
#if defined(__CUDA_ARCH__)
    return __brev(v) >> (32 - n);
#else
    uint32_t r = 0;
    for (int i = 0; i < n; i++) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
#endif
}

// Code lengths (already in shared memory, visible to every lane) -> counts per length, symbols in code order (lane 0)
// and the direct table (all lanes).  Returns < 0 for an over-subscribed set, > 0 for an incomplete one (only
// acceptable for a single code of one bit, checked by the caller), 0 when complete.  The group is synchronised on
// return.
template <int LANES, int BITS>
GZI_HD int construct(uint16_t* count, uint16_t* symbol, uint16_t* lut, const uint8_t* length, int n, int lane) {
    // every lane derives the counts per length for itself (registers), lane 0 publishes them
    uint32_t cnt[MAXBITS + 1];
    GZI_UNROLL
    for (int len = 0; len <= MAXBITS; len++) cnt[len] = 0;
    for (int s = 0; s < n; s++) {
        const int l = length[s];
        GZI_UNROLL
        for (int len = 0; len <= MAXBITS; len++) cnt[len] += (l == len) ? 1u : 0u;
    }
    int left = 1;
    bool over = false;
    GZI_UNROLL
    for (int len = 1; len <= MAXBITS; len++) {
        left <<= 1;
        left -= (int)cnt[len];
        over = over || left < 0;
    }
    if ((int)cnt[0] == n) left = 0;  // no codes at all
    // offs[len]: index of the first symbol of that length in code order; first[len]: its code
    uint32_t offs[MAXBITS + 2], first[MAXBITS + 2];
    offs[1] = 0;
    first[1] = 0;
    GZI_UNROLL
    for (int len = 1; len <= MAXBITS; len++) {
        offs[len + 1] = offs[len] + cnt[len];
        first[len + 1] = (first[len] + cnt[len]) << 1;
    }
    GZI_SYNC();  // the previous tables are not read any more
    if (lane == 0) {
        GZI_UNROLL
        for (int len = 0; len <= MAXBITS; len++) count[len] = (uint16_t)cnt[len];
    }
    for (int e = lane; e < (1 << BITS); e += LANES) lut[e] = 0;
    if (over) {
        GZI_SYNC();
        return -1;
    }
    if (lane == 0) {
        uint32_t next[MAXBITS + 1];
        GZI_UNROLL
        for (int len = 1; len <= MAXBITS; len++) next[len] = offs[len];
        for (int s = 0; s < n; s++) {
            const int l = length[s];
            if (l == 0) continue;
            uint32_t at = 0;
            GZI_UNROLL
            for (int len = 1; len <= MAXBITS; len++)
                if (l == len) at = next[len]++;
            symbol[at] = (uint16_t)s;
        }
    }
    GZI_SYNC();
    // direct table: the symbol at position idx of the code order has the code first[l] + (idx - offs[l]) of l bits,
    // MSB first in the stream, i.e. bit-reversed in the (LSB first) bit buffer; every index whose low l bits equal it
    const int total = (int)offs[MAXBITS + 1];
    for (int idx = lane; idx < total; idx += LANES) {
        const int s = symbol[idx];
        const int l = length[s];
        if (l > BITS) continue;
        uint32_t f = 0, o = 0;
        GZI_UNROLL
        for (int len = 1; len <= MAXBITS; len++)
            if (l == len) {
                f = first[len];
                o = offs[len];
            }
        const uint32_t code = f + ((uint32_t)idx - o);
        const uint16_t entry = (uint16_t)(((uint32_t)s << 4) | (uint32_t)l);
        for (uint32_t e = bit_reverse(code, l); e < (1u << BITS); e += 1u << l) lut[e] = entry;
    }
    GZI_SYNC();
    return left;
}

// One gzip member src[0 .. n) -> dst[0 .. cap).  *produced = bytes written (valid in every lane).  Checks ISIZE and
// that the member ends where it should; the CRC-32 of the text is verified by the caller (k_gz_check / the host twin).
template <int LANES>
GZI_HD int inflate_member(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap, Tables& T, int lane, uint32_t* produced) {
    *produced = 0;
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
    BitIn b;
    b.end = src + n;
    start(b, src + pos);
    // out: bytes in dst; behind them `pend` (< LANES) literals that still sit in `mine` of lanes 0 .. pend - 1
    uint32_t out = 0, pend = 0;
    uint32_t mine = 0;
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
            const uint8_t* q = byte_pos(b);  // the whole bytes in the bit buffer are given back
            if (q + len > b.end || out + pend + len > cap) return ERR_OVERRUN;
            if ((uint32_t)lane < pend) dst[out + (uint32_t)lane] = (uint8_t)mine;
            out += pend;
            pend = 0;
            for (uint32_t i = (uint32_t)lane; i < len; i += LANES) dst[out + i] = q[i];
            out += len;
            start(b, q + len);
            continue;
        }
        if (type == 3) return ERR_BLOCK;
        if (type == 1) {  // fixed codes
            GZI_SYNC();
            for (int s = lane; s < FIXLCODES; s += LANES) T.lengths[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
            for (int s = lane; s < MAXDCODES; s += LANES) T.lengths[FIXLCODES + s] = 5;
            GZI_SYNC();
            construct<LANES, LBITS>(T.lcount, T.lsym, T.llut, T.lengths, FIXLCODES, lane);
            construct<LANES, DBITS>(T.dcount, T.dsym, T.dlut, T.lengths + FIXLCODES, MAXDCODES, lane);
        } else {  // dynamic codes
            refill(b);
            const int nlen = (int)take(b, 5) + 257, ndist = (int)take(b, 5) + 1, ncode = (int)take(b, 4) + 4;
            if (nlen > MAXLCODES || ndist > MAXDCODES) return ERR_CODE;
            // the code-length code: 19 lengths of 3 bits in a fixed order, read by every lane, stored by lane 0
            uint32_t packed[2] = {0, 0};  // 19 x 3 bits, by position in the stream
            for (int idx = 0; idx < ncode; idx++) {
                refill(b);
                const uint32_t v = take(b, 3);
                if (idx < 10)
                    packed[0] |= v << (3 * idx);
                else
                    packed[1] |= v << (3 * (idx - 10));
            }
            GZI_SYNC();
            if (lane == 0) {
                const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                for (int idx = 0; idx < 19; idx++) {
                    const uint32_t v = idx < 10 ? (packed[0] >> (3 * idx)) & 7u : (packed[1] >> (3 * (idx - 10))) & 7u;
                    T.lengths[order[idx]] = (uint8_t)(idx < ncode ? v : 0u);
                }
            }
            GZI_SYNC();
            // its tables live where the distance tables will be (19 symbols, codes of at most 7 bits <= DBITS)
            if (construct<LANES, DBITS>(T.dcount, T.dsym, T.dlut, T.lengths, 19, lane) != 0) return ERR_CODE;  // must be complete
            int idx = 0, prev = 0;
            bool have_eob = false;
            while (idx < nlen + ndist) {
                refill(b);
                const int sym = decode_sym<DBITS>(b, T.dlut, T.dcount, T.dsym);
                if (sym < 0) return ERR_CODE;
                int rep = 1, val = sym;
                if (sym == 16) {
                    if (idx == 0) return ERR_CODE;
                    val = prev;
                    rep = 3 + (int)take(b, 2);
                } else if (sym == 17) {
                    val = 0;
                    rep = 3 + (int)take(b, 3);
                } else if (sym == 18) {
                    val = 0;
                    rep = 11 + (int)take(b, 7);
                }
                if (idx + rep > nlen + ndist) return ERR_CODE;
                if (val != 0 && idx <= 256 && idx + rep > 256) have_eob = true;
                // literal/length lengths at [0, nlen), distance lengths behind them at [nlen, nlen + ndist)
                for (int r = lane; r < rep; r += LANES) T.lengths[idx + r] = (uint8_t)val;
                idx += rep;
                prev = val;
            }
            if (!have_eob) return ERR_CODE;  // no end-of-block code
            GZI_SYNC();
            // the distance code first: the code-length tables in its place are done with
            int err = construct<LANES, DBITS>(T.dcount, T.dsym, T.dlut, T.lengths + nlen, ndist, lane);
            if (err < 0 || (err > 0 && ndist - (int)T.dcount[0] != 1)) return ERR_CODE;
            err = construct<LANES, LBITS>(T.lcount, T.lsym, T.llut, T.lengths, nlen, lane);
            if (err < 0 || (err > 0 && nlen - (int)T.lcount[0] != 1)) return ERR_CODE;
        }
        for (;;) {  // the symbols of the block
            // literals whose code sits in the direct table: the loop FASTQ text spends most of its symbols in
            uint32_t e;
            for (;;) {
                refill(b);
                e = T.llut[(uint32_t)b.buf & ((1u << LBITS) - 1u)];
                if (e - 1u >= (256u << 4) - 1u) break;  // not in the table, a length code or the end of the block
                const int nb = (int)(e & 15u);
                b.buf >>= nb;
                b.cnt -= nb;
                if ((uint32_t)lane == pend) mine = e >> 4;
                pend++;
                if (pend == (uint32_t)LANES) {
                    if (out + LANES > cap) return ERR_OVERRUN;
                    dst[out + (uint32_t)lane] = (uint8_t)mine;
                    out += LANES;
                    pend = 0;
                }
            }
            int sym;
            if (e) {
                const int nb = (int)(e & 15u);
                b.buf >>= nb;
                b.cnt -= nb;
                sym = (int)(e >> 4);
            } else {
                sym = decode_slow(b, T.lcount, T.lsym);
                if (sym < 0) return ERR_CODE;
            }
            if (sym < 256) {  // a literal with a long code
                if ((uint32_t)lane == pend) mine = (uint32_t)sym;
                pend++;
                if (pend == (uint32_t)LANES) {
                    if (out + LANES > cap) return ERR_OVERRUN;
                    dst[out + (uint32_t)lane] = (uint8_t)mine;
                    out += LANES;
                    pend = 0;
                }
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
                const int x = (sym >> 2) - 1;
                len = ((4u + (uint32_t)(sym & 3)) << x) + 3u + take(b, x);
            }
            refill(b);
            const int ds = decode_sym<DBITS>(b, T.dlut, T.dcount, T.dsym);
            if (ds < 0 || ds >= 30) return ERR_DIST;
            uint32_t dist;  // 1..32768: codes 0..3 are 1..4, then pairs of codes per extra-bit count
            if (ds < 4) {
                dist = 1u + (uint32_t)ds;
            } else {
                const int x = (ds >> 1) - 1;
                dist = ((2u + (uint32_t)(ds & 1)) << x) + 1u + take(b, x);
            }
            // the pending literals first
            if (out + pend + len > cap) return ERR_OVERRUN;
            if ((uint32_t)lane < pend) dst[out + (uint32_t)lane] = (uint8_t)mine;
            out += pend;
            pend = 0;
            if (dist > out) return ERR_DIST;
            // The copy: byte i of the match is byte (i mod dist) of the `dist` bytes in front of it, which all exist
            // already - no lane waits for another one's store.  The source was written a moment ago (every load is a
            // round trip to L2), so the loads of a long copy are issued before its first store.
            GZI_SYNC();  // the stores of the other lanes (literals, earlier copies) are visible
            uint8_t* d = dst + out;
            const uint8_t* sp = d - dist;
            if (dist >= len) {
                if (len <= (uint32_t)LANES) {
                    if ((uint32_t)lane < len) d[lane] = sp[lane];
                } else {
                    for (uint32_t base = 0; base < len; base += 4u * LANES) {
                        uint8_t v[4];
                        GZI_UNROLL
                        for (uint32_t k = 0; k < 4; k++) {
                            const uint32_t i = base + k * LANES + (uint32_t)lane;
                            if (i < len) v[k] = sp[i];
                        }
                        GZI_UNROLL
                        for (uint32_t k = 0; k < 4; k++) {
                            const uint32_t i = base + k * LANES + (uint32_t)lane;
                            if (i < len) d[i] = v[k];
                        }
                    }
                }
            } else if (dist == 1) {  // a run of one byte
                const uint8_t c = sp[0];
                for (uint32_t i = (uint32_t)lane; i < len; i += LANES) d[i] = c;
            } else {  // a period shorter than the match
                uint32_t r = (uint32_t)lane % dist;
                const uint32_t step = (uint32_t)LANES % dist;
                for (uint32_t i = (uint32_t)lane; i < len; i += LANES) {
                    d[i] = sp[r];
                    r += step;
                    if (r >= dist) r -= dist;
                }
            }
            out += len;
        }
    } while (!last);
    if (out + pend > cap) return ERR_OVERRUN;
    if ((uint32_t)lane < pend) dst[out + (uint32_t)lane] = (uint8_t)mine;
    out += pend;
    GZI_SYNC();
    // trailer: CRC-32 (verified by the caller), ISIZE
    const uint8_t* tr = byte_pos(b);
    if (tr + 8 > b.end) return ERR_TRAILER;
    const uint32_t isize = tr[4] | ((uint32_t)tr[5] << 8) | ((uint32_t)tr[6] << 16) | ((uint32_t)tr[7] << 24);
    if (isize != out) return ERR_TRAILER;
    *produced = out;
    return OK;
}

}  // namespace gzi
