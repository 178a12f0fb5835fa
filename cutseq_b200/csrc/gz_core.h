// DEFLATE / gzip pieces shared by the device codecs (gz_deflate.cu, gz_inflate.cu) and their host-side twins that the
// CPU tests run (csq_gz_* test entry points): CRC-32 arithmetic, length-limited Huffman code construction, the
// dynamic-block header.  Stands where xopen's gzip backends (zlib / isal / pigz) sit behind cutadapt's output files in
// the reference (run.py:449-470, 767-790): "@name\nseq\n+\nqual\n" text -> .fastq.gz.
// RFC 1951 (DEFLATE), RFC 1952 (gzip), SAM specification 4.1 (BGZF members).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define GZ_HD __host__ __device__ __forceinline__
#else
#define GZ_HD inline
#endif

namespace gz {

constexpr int N_LITLEN = 257;       // literals 0..255 and end-of-block: this encoder emits no length codes
constexpr int MAX_BITS = 15;
constexpr uint32_t CRC_POLY = 0xEDB88320u;

// ---- CRC-32 (reflected, as gzip) -------------------------------------------------------------------------------------
// a(x) * b(x) mod P in the reflected representation (bit 31 = x^0), as zlib's multmodp
GZ_HD uint32_t crc_mulmod(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1u)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 n) mod P: the operator "n more bytes follow" of crc32_combine
inline uint32_t crc_xpow8(uint64_t n) {
    uint32_t p = 1u << 31;        // x^0
    uint32_t sq = 0x00800000u;    // x^8 (reflected: bit 31 - 8)
    while (n) {
        if (n & 1) p = crc_mulmod(sq, p);
        sq = crc_mulmod(sq, sq);
        n >>= 1;
    }
    return p;
}
inline void crc_make_table(uint32_t* t /*[256]*/) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
        t[i] = c;
    }
}

// ---- Huffman code lengths -------------------------------------------------------------------------------------------
// Code lengths (<= max_bits) for n symbols with frequencies f[]; symbols with f == 0 get length 0.  A single used
// symbol gets length 1.  Minimum-redundancy lengths by Moffat & Katajainen's in-place algorithm on the sorted
// frequencies, then the usual length limiting on the histogram of lengths (fold the too-long codes into max_bits and
// repair the Kraft sum), lengths handed out by rank (rarest symbols get the longest codes).
// scratch: order[n] (uint16_t), work[n] (uint32_t).
GZ_HD void huff_lengths(const uint32_t* f, int n, int max_bits, uint8_t* len, uint16_t* order, uint32_t* work) {
    int used = 0;
    for (int i = 0; i < n; i++) {
        len[i] = 0;
        if (f[i]) order[used++] = (uint16_t)i;
    }
    if (used == 0) return;
    if (used == 1) {
        len[order[0]] = 1;
        return;
    }
    // insertion sort by (frequency, symbol) ascending - n <= 286, run by one thread per code
    for (int i = 1; i < used; i++) {
        const uint16_t s = order[i];
        const uint32_t fs = f[s];
        int j = i - 1;
        while (j >= 0 && (f[order[j]] > fs || (f[order[j]] == fs && order[j] > s))) {
            order[j + 1] = order[j];
            j--;
        }
        order[j + 1] = s;
    }
    uint32_t* A = work;
    for (int i = 0; i < used; i++) A[i] = f[order[i]];
    {  // Moffat & Katajainen, "In-place calculation of minimum-redundancy codes" (1995)
        const int m = used;
        int root = 0, leaf = 2, next;
        A[0] += A[1];
        for (next = 1; next < m - 1; next++) {
            if (leaf >= m || A[root] < A[leaf]) {
                A[next] = A[root];
                A[root++] = (uint32_t)next;
            } else {
                A[next] = A[leaf++];
            }
            if (leaf >= m || (root < next && A[root] < A[leaf])) {
                A[next] += A[root];
                A[root++] = (uint32_t)next;
            } else {
                A[next] += A[leaf++];
            }
        }
        A[m - 2] = 0;
        for (next = m - 3; next >= 0; next--) A[next] = A[A[next]] + 1;
        int avbl = 1, usedn = 0, dpth = 0;
        root = m - 2;
        next = m - 1;
        while (avbl > 0) {
            while (root >= 0 && (int)A[root] == dpth) {
                usedn++;
                root--;
            }
            while (avbl > usedn) {
                A[next--] = (uint32_t)dpth;
                avbl--;
            }
            avbl = 2 * usedn;
            dpth++;
            usedn = 0;
        }
    }
    // A[i] = depth of the i-th rarest symbol (non-increasing in i).  Histogram of lengths, limited to max_bits.
    uint32_t num[64];
    for (int i = 0; i < 64; i++) num[i] = 0;
    for (int i = 0; i < used; i++) num[A[i] < 63 ? A[i] : 63]++;
    for (int i = max_bits + 1; i < 64; i++) {
        num[max_bits] += num[i];
        num[i] = 0;
    }
    uint32_t total = 0;
    for (int i = max_bits; i > 0; i--) total += num[i] << (max_bits - i);
    while (total != (1u << max_bits)) {
        num[max_bits]--;
        for (int i = max_bits - 1; i > 0; i--)
            if (num[i]) {
                num[i]--;
                num[i + 1] += 2;
                break;
            }
        total--;
    }
    int r = 0;  // rarest first: longest codes
    for (int l = max_bits; l >= 1; l--)
        for (uint32_t c = 0; c < num[l]; c++) len[order[r++]] = (uint8_t)l;
}

GZ_HD uint32_t bit_reverse(uint32_t v, int n) {  // DEFLATE sends Huffman codes most significant bit first
    uint32_t r = 0;
    for (int i = 0; i < n; i++) {
        r = (r << 1) | (v & 1u);
        v >>= 1;
    }
    return r;
}

// Canonical codes (RFC 1951 3.2.2) for lengths len[0..n), bit-reversed so that they can be OR-ed into an LSB-first
// bit stream: code[i] | len[i] << 16.
GZ_HD void huff_codes(const uint8_t* len, int n, uint32_t* code) {
    uint32_t bl_count[MAX_BITS + 1], next_code[MAX_BITS + 2];
    for (int i = 0; i <= MAX_BITS; i++) bl_count[i] = 0;
    for (int i = 0; i < n; i++) bl_count[len[i]]++;
    bl_count[0] = 0;
    uint32_t c = 0;
    next_code[0] = 0;
    for (int b = 1; b <= MAX_BITS; b++) {
        c = (c + bl_count[b - 1]) << 1;
        next_code[b] = c;
    }
    for (int i = 0; i < n; i++) {
        const int l = len[i];
        code[i] = l ? (bit_reverse(next_code[l]++, l) | ((uint32_t)l << 16)) : 0u;
    }
}

// LSB-first bit writer into 32-bit words (zeroed by the caller)
struct BitSink {
    uint32_t* w;
    uint32_t pos;  // bits written
};
GZ_HD void put_bits(BitSink& s, uint32_t v, int n) {
    if (n == 0) return;
    const uint32_t word = s.pos >> 5, sh = s.pos & 31u;
    s.w[word] |= v << sh;
    if (sh + (uint32_t)n > 32u) s.w[word + 1] |= v >> (32u - sh);
    s.pos += (uint32_t)n;
}

// The bits of a dynamic-Huffman block header (RFC 1951 3.2.7) for a literal-only code: BFINAL = 1, BTYPE = 2,
// HLIT = 257 literal/length codes, HDIST = 2 distance codes of one bit each (what zlib emits for a block without
// matches), the code-length code, and the run-length coded lengths.  hdr must hold HDR_WORDS zeroed words.
constexpr int HDR_WORDS = 96;
GZ_HD uint32_t dyn_header(const uint8_t* litlen /*[257]*/, uint32_t* hdr) {
    uint8_t seq[N_LITLEN + 2];
    for (int i = 0; i < N_LITLEN; i++) seq[i] = litlen[i];
    seq[N_LITLEN] = 1;
    seq[N_LITLEN + 1] = 1;
    const int nseq = N_LITLEN + 2;
    // run-length code: (symbol, extra value) pairs
    uint8_t rsym[N_LITLEN + 2];
    uint8_t rext[N_LITLEN + 2];
    int nr = 0;
    uint32_t clf[19];
    for (int i = 0; i < 19; i++) clf[i] = 0;
    for (int i = 0; i < nseq;) {
        const int v = seq[i];
        int run = 1;
        while (i + run < nseq && seq[i + run] == v) run++;
        int left = run;
        if (v == 0) {
            while (left >= 11) {
                const int r = left > 138 ? 138 : left;
                rsym[nr] = 18;
                rext[nr++] = (uint8_t)(r - 11);
                left -= r;
            }
            if (left >= 3) {
                rsym[nr] = 17;
                rext[nr++] = (uint8_t)(left - 3);
                left = 0;
            }
        } else {
            rsym[nr] = (uint8_t)v;  // the length itself, then repeats of it
            rext[nr++] = 0;
            left--;
            while (left >= 3) {
                const int r = left > 6 ? 6 : left;
                rsym[nr] = 16;
                rext[nr++] = (uint8_t)(r - 3);
                left -= r;
            }
        }
        while (left > 0) {
            rsym[nr] = (uint8_t)v;
            rext[nr++] = 0;
            left--;
        }
        i += run;
    }
    for (int i = 0; i < nr; i++) clf[rsym[i]]++;
    uint8_t cll[19];
    uint16_t order[19];
    uint32_t work[19], clc[19];
    huff_lengths(clf, 19, 7, cll, order, work);
    huff_codes(cll, 19, clc);
    const uint8_t perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && cll[perm[hclen - 1]] == 0) hclen--;
    BitSink s = {hdr, 0};
    put_bits(s, 1, 1);                      // BFINAL
    put_bits(s, 2, 2);                      // BTYPE = dynamic
    put_bits(s, N_LITLEN - 257, 5);         // HLIT
    put_bits(s, 2 - 1, 5);                  // HDIST
    put_bits(s, (uint32_t)(hclen - 4), 4);  // HCLEN
    for (int i = 0; i < hclen; i++) put_bits(s, cll[perm[i]], 3);
    for (int i = 0; i < nr; i++) {
        const int sym = rsym[i];
        put_bits(s, clc[sym] & 0xFFFFu, (int)(clc[sym] >> 16));
        if (sym == 16) put_bits(s, rext[i], 2);
        else if (sym == 17) put_bits(s, rext[i], 3);
        else if (sym == 18) put_bits(s, rext[i], 7);
    }
    return s.pos;
}

// gzip member framing as BGZF (SAM specification 4.1): 18 header bytes with the 'BC' extra field holding
// (member size - 1), the DEFLATE data, CRC-32 and ISIZE.
constexpr int GZ_HEAD = 18, GZ_TAIL = 8;
GZ_HD void bgzf_header(uint8_t* p, uint32_t member_size) {
    const uint8_t h[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    for (int i = 0; i < 16; i++) p[i] = h[i];
    p[16] = (uint8_t)((member_size - 1u) & 0xFFu);
    p[17] = (uint8_t)((member_size - 1u) >> 8);
}

}  // namespace gz
