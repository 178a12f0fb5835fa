// gzip (RFC 1952) / DEFLATE (RFC 1951) decoder of the input side of the file driver.
//
// Stands where xopen's decompression backends (python-isal / zlib-ng / gzip) sit behind cutadapt's InputPaths
// in the reference (run.py:434, 751).  System zlib inflates FASTQ at ~0.16-0.23 GB/s per thread on the hosts this
// runs on, which bounds the whole .gz pipeline (the GPU chain takes 2 % of that time); this decoder is written
// for speed on a 64-bit little-endian host:
//   * the whole compressed input is visible at once (the file is mmap-ed), the output is produced straight into
//     the caller's (pinned) buffer, any number of bytes per call, resumable at every symbol;
//   * 64-bit bit buffer refilled with one unaligned 8-byte load (branch-free, at most once per symbol);
//   * 11-bit (literal/length) and 8-bit (distance) lookup tables with second-level tables for longer codes, one
//     32-bit entry per code: base value | kind | extra-bit count | code length;
//   * literal runs through a multi-literal table (up to 3 literals per lookup, 12 per refill); matches copied
//     with unaligned 8-byte moves (byte / short-period forms
//     for distances below 8);
//   * a careful byte-wise loop takes over within 300 bytes of the end of the caller's buffer or 16 bytes of the
//     end of the input, so the fast loop needs no bounds checks;
//   * history that lies in front of the caller's buffer (previous call) is kept in a 32 KiB side window;
//   * concatenated members, header flags (FEXTRA, FNAME, FCOMMENT, FHCRC), CRC-32 and ISIZE are checked.
// Errors are reported, never papered over: a corrupt or truncated stream makes read() return -1.
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>  // crc32() only

#include "host_io.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace csqio {

namespace {

// CRC-32 of the decompressed bytes (the gzip trailer check).  zlib's table-driven crc32() runs at ~2 GB/s here,
// a sixth of the decoder's time; on x86 with PCLMULQDQ the bulk is folded 64 bytes per step with carry-less
// multiplications instead (V. Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ
// Instruction", Intel 2009; constants of the reflected CRC-32 polynomial 0x1DB710641), zlib takes what is left.
// tests/test_inflate.py checks it against zlib on every length and alignment class.
#if defined(__x86_64__)
__attribute__((target("pclmul,sse4.1"))) static uint32_t crc32_fold(uint32_t crc, const uint8_t* buf, size_t len) {
    // crc: the raw register (zlib's value inverted); len >= 64 and a multiple of 16
    const __m128i k1k2 = _mm_set_epi64x(0x01c6e41596LL, 0x0154442bd4LL);
    const __m128i k3k4 = _mm_set_epi64x(0x00ccaa009eLL, 0x01751997d0LL);
    const __m128i k5 = _mm_set_epi64x(0, 0x0163cd6124LL);
    const __m128i poly = _mm_set_epi64x(0x01f7011641LL, 0x01db710641LL);
    const __m128i mask32 = _mm_set_epi32(0, 0, 0, -1);
    __m128i x1 = _mm_loadu_si128((const __m128i*)(buf + 0)), x2 = _mm_loadu_si128((const __m128i*)(buf + 16));
    __m128i x3 = _mm_loadu_si128((const __m128i*)(buf + 32)), x4 = _mm_loadu_si128((const __m128i*)(buf + 48));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
    buf += 64;
    len -= 64;
    while (len >= 64) {
        const __m128i y1 = _mm_clmulepi64_si128(x1, k1k2, 0x00), y2 = _mm_clmulepi64_si128(x2, k1k2, 0x00);
        const __m128i y3 = _mm_clmulepi64_si128(x3, k1k2, 0x00), y4 = _mm_clmulepi64_si128(x4, k1k2, 0x00);
        x1 = _mm_clmulepi64_si128(x1, k1k2, 0x11);
        x2 = _mm_clmulepi64_si128(x2, k1k2, 0x11);
        x3 = _mm_clmulepi64_si128(x3, k1k2, 0x11);
        x4 = _mm_clmulepi64_si128(x4, k1k2, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, y1), _mm_loadu_si128((const __m128i*)(buf + 0)));
        x2 = _mm_xor_si128(_mm_xor_si128(x2, y2), _mm_loadu_si128((const __m128i*)(buf + 16)));
        x3 = _mm_xor_si128(_mm_xor_si128(x3, y3), _mm_loadu_si128((const __m128i*)(buf + 32)));
        x4 = _mm_xor_si128(_mm_xor_si128(x4, y4), _mm_loadu_si128((const __m128i*)(buf + 48)));
        buf += 64;
        len -= 64;
    }
    // four accumulators -> one
    __m128i y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), y), x2);
    y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), y), x3);
    y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
    x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), y), x4);
    while (len >= 16) {
        y = _mm_clmulepi64_si128(x1, k3k4, 0x00);
        x1 = _mm_xor_si128(_mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x11), y), _mm_loadu_si128((const __m128i*)buf));
        buf += 16;
        len -= 16;
    }
    // 128 -> 64 -> 32 bits, then Barrett reduction
    __m128i x = _mm_xor_si128(_mm_clmulepi64_si128(x1, k3k4, 0x10), _mm_srli_si128(x1, 8));
    __m128i hi = _mm_srli_si128(x, 4);
    x = _mm_xor_si128(_mm_clmulepi64_si128(_mm_and_si128(x, mask32), k5, 0x00), hi);
    __m128i t = _mm_clmulepi64_si128(_mm_and_si128(x, mask32), poly, 0x10);
    t = _mm_clmulepi64_si128(_mm_and_si128(t, mask32), poly, 0x00);
    return (uint32_t)_mm_extract_epi32(_mm_xor_si128(x, t), 1);
}
#endif

uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
#if defined(__x86_64__)
    static const bool have_clmul = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    if (have_clmul && n >= 256) {
        const size_t bulk = n & ~(size_t)15;
        crc = ~crc32_fold(~crc, p, bulk);
        p += bulk;
        n -= bulk;
    }
#endif
    while (n) {  // crc32() takes 32-bit lengths
        const uInt c = n > (1u << 30) ? (1u << 30) : (uInt)n;
        crc = (uint32_t)crc32(crc, p, c);
        p += c;
        n -= c;
    }
    return crc;
}

constexpr int LL_BITS = 11, D_BITS = 8;
constexpr uint32_t K_LIT = 0, K_LEN = 1, K_EOB = 2, K_SUB = 3;

inline uint32_t mk_entry(uint32_t value, uint32_t kind, uint32_t extra, uint32_t len) {
    return (value << 16) | (kind << 14) | (extra << 8) | len;
}
inline uint32_t e_len(uint32_t e) { return e & 0xFFu; }
inline uint32_t e_extra(uint32_t e) { return (e >> 8) & 0x3Fu; }
inline uint32_t e_kind(uint32_t e) { return (e >> 14) & 3u; }
inline uint32_t e_value(uint32_t e) { return e >> 16; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t bit_reverse(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) {
        r = (r << 1) | (code & 1u);
        code >>= 1;
    }
    return r;
}

}  // namespace

uint32_t crc32_fast(uint32_t crc, const uint8_t* p, size_t n) { return crc32_update(crc, p, n); }

struct Inflater::Tables {
    uint32_t ll[(1 << LL_BITS) + 2048];  // second-level tables behind the first 2^LL_BITS entries
    uint32_t dd[(1 << D_BITS) + 1024];
    // multi-literal table: what the next LL_BITS bits hold when they start with 1..3 complete literal codes:
    // literals in bytes 0..2 | total code length << 24 | number of literals << 28; 0 = not a literal (use ll).
    // Serial Huffman decoding is bound by the lookup -> shift -> lookup dependency chain (~7 cycles per symbol);
    // FASTQ text has 2-4 bit codes for its few frequent characters, so one chain step here yields 2-3 bytes.
    uint32_t ml[1 << LL_BITS];
};

static void build_multi_literal(const uint32_t* ll, uint32_t* ml) {
    for (uint32_t i = 0; i < (1u << LL_BITS); i++) {
        uint32_t e = ll[i];
        if (e_kind(e) != K_LIT || e_len(e) == 0) {
            ml[i] = 0;
            continue;
        }
        uint32_t lits = e_value(e), used = e_len(e), n = 1;
        while (n < 3) {
            e = ll[i >> used];  // only LL_BITS - used bits of this index are real: an entry is valid if its code fits in them
            if (e_kind(e) != K_LIT || e_len(e) == 0 || used + e_len(e) > (uint32_t)LL_BITS) break;
            lits |= e_value(e) << (8 * n);
            used += e_len(e);
            n++;
        }
        ml[i] = lits | (used << 24) | (n << 28);
    }
}

// Canonical Huffman code (lengths per symbol) -> lookup table.  kind_of(sym) gives the entry payload.
// Returns false for an over-subscribed code or one that does not fit the table space.
template <typename F>
static bool build_table(const uint8_t* lens, int n_sym, int table_bits, uint32_t* table, size_t table_cap, F payload) {
    int count[16] = {0};
    for (int s = 0; s < n_sym; s++) count[lens[s]]++;
    count[0] = 0;
    // over-subscription check (incomplete codes are allowed: unused entries stay invalid)
    int left = 1;
    for (int l = 1; l <= 15; l++) {
        left <<= 1;
        left -= count[l];
        if (left < 0) return false;
    }
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; l++) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
    }
    const size_t main_size = (size_t)1 << table_bits;
    for (size_t i = 0; i < main_size; i++) table[i] = 0;  // len 0 == invalid
    // pass 1: longest code behind every first-level prefix that needs a second level
    uint8_t sub_bits[1 << LL_BITS];
    memset(sub_bits, 0, main_size);
    uint32_t codes[320];
    for (int s = 0; s < n_sym; s++) {
        const int l = lens[s];
        if (!l) continue;
        codes[s] = bit_reverse(next_code[l]++, l);
        if (l > table_bits) {
            const uint32_t prefix = codes[s] & (uint32_t)(main_size - 1);
            if (l - table_bits > sub_bits[prefix]) sub_bits[prefix] = (uint8_t)(l - table_bits);
        }
    }
    // pass 2: allocate the second-level tables
    size_t next_free = main_size;
    for (size_t pfx = 0; pfx < main_size; pfx++)
        if (sub_bits[pfx]) {
            const size_t sz = (size_t)1 << sub_bits[pfx];
            if (next_free + sz > table_cap) return false;
            table[pfx] = mk_entry((uint32_t)next_free, K_SUB, sub_bits[pfx], (uint32_t)table_bits);
            for (size_t i = 0; i < sz; i++) table[next_free + i] = 0;
            next_free += sz;
        }
    // pass 3: fill
    for (int s = 0; s < n_sym; s++) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t c = codes[s];
        if (l <= table_bits) {
            const uint32_t e = payload(s, (uint32_t)l);
            for (size_t i = c; i < main_size; i += (size_t)1 << l) table[i] = e;
        } else {
            const uint32_t prefix = c & (uint32_t)(main_size - 1);
            const uint32_t sub = table[prefix];
            const uint32_t sb = e_extra(sub), base = e_value(sub);
            const uint32_t e = payload(s, (uint32_t)l);
            for (uint32_t i = c >> table_bits; i < (1u << sb); i += 1u << (l - table_bits)) table[base + i] = e;
        }
    }
    return true;
}

static uint32_t ll_payload(int s, uint32_t l) {
    if (s < 256) return mk_entry((uint32_t)s, K_LIT, 0, l);
    if (s == 256) return mk_entry(0, K_EOB, 0, l);
    if (s > 285) return 0;  // invalid symbol: decoding it is an error
    return mk_entry(LEN_BASE[s - 257], K_LEN, LEN_EXTRA[s - 257], l);
}
static uint32_t d_payload(int s, uint32_t l) {
    if (s > 29) return 0;
    return mk_entry(DIST_BASE[s], K_LEN, DIST_EXTRA[s], l);
}

Inflater::Inflater() : t_(new Tables()) {}
Inflater::~Inflater() { delete t_; }

void Inflater::reset(const uint8_t* data, size_t n) {
    in_ = data;
    in_end_ = data + n;
    bitbuf_ = 0;
    bits_ = 0;
    state_ = S_HEADER;
    last_block_ = false;
    stored_left_ = 0;
    pend_len_ = pend_dist_ = 0;
    win_len_ = 0;
    crc_ = 0;
    isize_ = 0;
    err_ = nullptr;
    members_ = 0;
}

// ---- careful bit access (byte-wise refill, bounds checked) ----
inline bool Inflater::need(int n) {
    while (bits_ < n) {
        if (in_ >= in_end_) return false;
        bitbuf_ |= (uint64_t)*in_++ << bits_;
        bits_ += 8;
    }
    return true;
}
inline uint32_t Inflater::take(int n) {
    const uint32_t v = (uint32_t)(bitbuf_ & ((1ull << n) - 1));
    bitbuf_ >>= n;
    bits_ -= n;
    return v;
}
// gives whole bytes of look-ahead back to the input (before byte-aligned structures)
inline void Inflater::align_to_byte() {
    const int drop = bits_ & 7;
    bitbuf_ >>= drop;
    bits_ -= drop;
    in_ -= bits_ >> 3;
    bitbuf_ = 0;
    bits_ = 0;
}

bool Inflater::fail(const char* msg) {
    err_ = msg;
    state_ = S_ERROR;
    return false;
}

bool Inflater::parse_header() {
    // in_ is byte aligned here
    if (in_end_ - in_ < 10) return fail("truncated gzip header");
    if (in_[0] != 0x1f || in_[1] != 0x8b) return fail("not a gzip member (bad magic)");
    if (in_[2] != 8) return fail("unknown gzip compression method");
    const uint8_t flg = in_[3];
    if (flg & 0xE0) return fail("reserved gzip header flags set");
    const uint8_t* p = in_ + 10;
    if (flg & 4) {  // FEXTRA
        if (in_end_ - p < 2) return fail("truncated gzip header");
        const size_t xlen = p[0] | ((size_t)p[1] << 8);
        p += 2;
        if ((size_t)(in_end_ - p) < xlen) return fail("truncated gzip header");
        p += xlen;
    }
    for (int f = 8; f <= 16; f <<= 1)  // FNAME, FCOMMENT: zero-terminated
        if (flg & f) {
            const uint8_t* z = (const uint8_t*)memchr(p, 0, (size_t)(in_end_ - p));
            if (!z) return fail("truncated gzip header");
            p = z + 1;
        }
    if (flg & 2) {  // FHCRC
        if (in_end_ - p < 2) return fail("truncated gzip header");
        p += 2;
    }
    in_ = p;
    crc_ = (uint32_t)crc32(0L, Z_NULL, 0);
    isize_ = 0;
    members_++;
    state_ = S_BLOCK;
    last_block_ = false;
    return true;
}

bool Inflater::parse_block_header() {
    if (!need(3)) return fail("compressed data ends inside a block header");
    last_block_ = take(1) != 0;
    const uint32_t type = take(2);
    if (type == 0) {
        align_to_byte();
        if (in_end_ - in_ < 4) return fail("truncated stored block");
        const uint32_t len = in_[0] | ((uint32_t)in_[1] << 8), nlen = in_[2] | ((uint32_t)in_[3] << 8);
        if ((len ^ 0xFFFFu) != nlen) return fail("stored block length check failed");
        in_ += 4;
        stored_left_ = len;
        state_ = S_STORED;
        return true;
    }
    uint8_t lens[320];
    int hlit, hdist;
    if (type == 1) {
        for (int s = 0; s < 144; s++) lens[s] = 8;
        for (int s = 144; s < 256; s++) lens[s] = 9;
        for (int s = 256; s < 280; s++) lens[s] = 7;
        for (int s = 280; s < 288; s++) lens[s] = 8;
        for (int s = 288; s < 320; s++) lens[s] = 5;
        hlit = 288;
        hdist = 32;
    } else if (type == 2) {
        if (!need(14)) return fail("compressed data ends inside a block header");
        hlit = (int)take(5) + 257;
        hdist = (int)take(5) + 1;
        const int hclen = (int)take(4) + 4;
        if (hlit > 286 || hdist > 30) return fail("too many length or distance symbols");
        uint8_t cl[19] = {0};
        for (int i = 0; i < hclen; i++) {
            if (!need(3)) return fail("compressed data ends inside a block header");
            cl[CL_ORDER[i]] = (uint8_t)take(3);
        }
        uint32_t pre[128 + 64];
        if (!build_table(cl, 19, 7, pre, 128 + 64, [](int s, uint32_t l) { return mk_entry((uint32_t)s, K_LIT, 0, l); }))
            return fail("invalid code-length code");
        int i = 0;
        while (i < hlit + hdist) {
            if (!need(7)) {
                // the last code may be shorter than 7 bits at the very end of the input
                if (bits_ == 0) return fail("compressed data ends inside a block header");
            }
            const uint32_t e = pre[bitbuf_ & 127u];
            if (e_len(e) == 0 || (int)e_len(e) > bits_) return fail("invalid code-length symbol");
            take((int)e_len(e));
            const uint32_t sym = e_value(e);
            if (sym < 16) {
                lens[i++] = (uint8_t)sym;
                continue;
            }
            int rep, val = 0;
            if (sym == 16) {
                if (i == 0) return fail("length repeat without a previous length");
                if (!need(2)) return fail("compressed data ends inside a block header");
                val = lens[i - 1];
                rep = 3 + (int)take(2);
            } else if (sym == 17) {
                if (!need(3)) return fail("compressed data ends inside a block header");
                rep = 3 + (int)take(3);
            } else {
                if (!need(7)) return fail("compressed data ends inside a block header");
                rep = 11 + (int)take(7);
            }
            if (i + rep > hlit + hdist) return fail("length repeat runs past the code lengths");
            while (rep--) lens[i++] = (uint8_t)val;
        }
        if (lens[256] == 0) return fail("no end-of-block code");
        // the distance lengths follow the literal/length lengths: move them to their own array position
        memmove(lens + 288, lens + hlit, (size_t)hdist);
        for (int s = hlit; s < 288; s++) lens[s] = 0;
        for (int s = 288 + hdist; s < 320; s++) lens[s] = 0;
        hlit = 288;
        hdist = 32;
    } else {
        return fail("invalid block type");
    }
    if (!build_table(lens, hlit, LL_BITS, t_->ll, sizeof(t_->ll) / 4, ll_payload)) return fail("invalid literal/length code");
    if (!build_table(lens + 288, hdist, D_BITS, t_->dd, sizeof(t_->dd) / 4, d_payload)) return fail("invalid distance code");
    build_multi_literal(t_->ll, t_->ml);
    state_ = S_HUFFMAN;
    return true;
}

// One byte of history at distance `dist` behind position `out` (out > base by `have` bytes; older history in win_).
inline uint8_t Inflater::hist_byte(const uint8_t* base, const uint8_t* out, uint32_t dist) const {
    const size_t have = (size_t)(out - base);
    if (dist <= have) return out[-(ptrdiff_t)dist];
    return win_[win_len_ - (dist - have)];
}

// Decodes into [out, out_end); returns the new out.  Leaves state_ != S_HUFFMAN at the end of the block.
uint8_t* Inflater::run_huffman(uint8_t* const base, uint8_t* out, uint8_t* const out_end) {
    const uint32_t* const ll = t_->ll;
    const uint32_t* const dd = t_->dd;
    const uint32_t* const ml = t_->ml;
    // a match cut by the end of the previous buffer
    while (pend_len_ && out < out_end) {
        *out = hist_byte(base, out, pend_dist_);
        out++;
        pend_len_--;
    }
    if (pend_len_) return out;

    // ---- fast loop: no bounds checks inside ----
    if (out_end - out >= 320 && in_end_ - in_ >= 32) {
        uint8_t* const out_fast_end = out_end - 300;
        const uint8_t* const in_fast_end = in_end_ - 16;
        uint64_t bb = bitbuf_;
        uint32_t bl = (uint32_t)bits_;
        const uint8_t* in = in_;
        const size_t hist0 = win_len_;
#define REFILL()                                             \
    do {                                                     \
        uint64_t w_;                                         \
        memcpy(&w_, in, 8);                                  \
        bb |= w_ << (bl & 63);                               \
        in += 7 - ((bl >> 3) & 7);                           \
        bl |= 56;                                            \
    } while (0)
        bool end_of_block = false;
        while (out < out_fast_end && in < in_fast_end) {
            REFILL();
            // One lookup decides: in text compressed at a low level (zlib -1 on FASTQ: 96 % of the symbols are
            // matches of ~6 bytes) the match path must not pay for a literal-table miss first.
            uint32_t e = ll[bb & ((1u << LL_BITS) - 1)];
            if (e_kind(e) == K_LIT && e_len(e) && e_len(e) <= (uint32_t)LL_BITS) {
                // runs of literals: up to 3 per lookup, up to 4 lookups (44 bits) per refill; the 32-bit store writes
                // up to 3 bytes more than it advances (room is kept)
                uint32_t m = ml[bb & ((1u << LL_BITS) - 1)];
                int rounds = 4;
                do {
                    memcpy(out, &m, 4);
                    out += m >> 28;
                    const uint32_t used = (m >> 24) & 15u;
                    bb >>= used;
                    bl -= used;
                    m = ml[bb & ((1u << LL_BITS) - 1)];
                } while (m && --rounds);
                continue;
            }
            if (e_kind(e) == K_SUB) e = ll[e_value(e) + ((bb >> LL_BITS) & ((1u << e_extra(e)) - 1))];
            if (e_kind(e) == K_LIT && e_len(e)) {  // a literal with a code longer than LL_BITS
                bb >>= e_len(e);
                bl -= e_len(e);
                *out++ = (uint8_t)e_value(e);
                continue;
            }
            if (e_len(e) == 0) {
                fail("invalid literal/length code in the data");
                break;
            }
            bb >>= e_len(e);
            bl -= e_len(e);
            if (e_kind(e) == K_EOB) {
                end_of_block = true;
                break;
            }
            // a length: at most 5 extra bits, then a distance of at most 15 + 13 bits: 48 bits with the code itself
            uint32_t len = e_value(e) + (uint32_t)(bb & ((1u << e_extra(e)) - 1));
            bb >>= e_extra(e);
            bl -= e_extra(e);
            uint32_t d = dd[bb & ((1u << D_BITS) - 1)];
            if (e_kind(d) == K_SUB) d = dd[e_value(d) + ((bb >> D_BITS) & ((1u << e_extra(d)) - 1))];
            if (e_len(d) == 0) {
                fail("invalid distance code in the data");
                break;
            }
            bb >>= e_len(d);
            bl -= e_len(d);
            const uint32_t dist = e_value(d) + (uint32_t)(bb & ((1u << e_extra(d)) - 1));
            bb >>= e_extra(d);
            bl -= e_extra(d);
            const size_t have = (size_t)(out - base);
            if (dist > have) {  // (part of) the match lies in the previous buffer
                if (dist > have + hist0) {
                    fail("distance reaches in front of the start of the data");
                    break;
                }
                while (len && dist > (size_t)(out - base)) {
                    *out = win_[hist0 - (dist - (size_t)(out - base))];
                    out++;
                    len--;
                }
                for (; len; len--, out++) *out = out[-(ptrdiff_t)dist];
                continue;
            }
            const uint8_t* src = out - dist;
            uint8_t* const end = out + len;
            if (dist >= 8) {
                // may write up to 7 bytes past `end`: the fast loop keeps 300 bytes of room; most matches are short,
                // the first 16 bytes go without a loop
                uint64_t w0, w1;
                memcpy(&w0, src, 8);
                memcpy(out, &w0, 8);
                if (len > 8) {
                    memcpy(&w1, src + 8, 8);
                    memcpy(out + 8, &w1, 8);
                    src += 16;
                    out += 16;
                    while (out < end) {
                        uint64_t w;
                        memcpy(&w, src, 8);
                        memcpy(out, &w, 8);
                        src += 8;
                        out += 8;
                    }
                }
            } else if (dist == 1) {
                memset(out, *src, len);
            } else {
                do {
                    *out++ = *src++;
                } while (out < end);
            }
            out = end;
        }
#undef REFILL
        // hand the look-ahead back in whole bytes: the careful loop refills byte by byte
        bl &= 63;
        in -= bl >> 3;
        bl &= 7;
        bitbuf_ = bb & ((1ull << bl) - 1);
        bits_ = (int)bl;
        in_ = in;
        if (state_ == S_ERROR) return out;
        if (end_of_block) {
            state_ = last_block_ ? S_TRAILER : S_BLOCK;
            return out;
        }
    }

    // ---- careful loop ----
    while (out < out_end) {
        need(15);  // as many bits as there are; the code check below catches a real shortage
        uint32_t e = ll[bitbuf_ & ((1u << LL_BITS) - 1)];
        if (e_kind(e) == K_SUB && e_len(e)) e = ll[e_value(e) + ((bitbuf_ >> LL_BITS) & ((1u << e_extra(e)) - 1))];
        if (e_len(e) == 0 || (int)e_len(e) > bits_) {
            fail(in_ >= in_end_ ? "compressed data ends inside a block" : "invalid literal/length code in the data");
            return out;
        }
        take((int)e_len(e));
        if (e_kind(e) == K_LIT) {
            *out++ = (uint8_t)e_value(e);
            continue;
        }
        if (e_kind(e) == K_EOB) {
            state_ = last_block_ ? S_TRAILER : S_BLOCK;
            return out;
        }
        if (!need((int)e_extra(e))) {
            fail("compressed data ends inside a block");
            return out;
        }
        uint32_t len = e_value(e) + take((int)e_extra(e));
        need(15);
        uint32_t d = dd[bitbuf_ & ((1u << D_BITS) - 1)];
        if (e_kind(d) == K_SUB && e_len(d)) d = dd[e_value(d) + ((bitbuf_ >> D_BITS) & ((1u << e_extra(d)) - 1))];
        if (e_len(d) == 0 || (int)e_len(d) > bits_) {
            fail(in_ >= in_end_ ? "compressed data ends inside a block" : "invalid distance code in the data");
            return out;
        }
        take((int)e_len(d));
        if (!need((int)e_extra(d))) {
            fail("compressed data ends inside a block");
            return out;
        }
        const uint32_t dist = e_value(d) + take((int)e_extra(d));
        if (dist > (size_t)(out - base) + win_len_) {
            fail("distance reaches in front of the start of the data");
            return out;
        }
        while (len && out < out_end) {
            *out = hist_byte(base, out, dist);
            out++;
            len--;
        }
        if (len) {  // the caller's buffer is full in the middle of a match
            pend_len_ = len;
            pend_dist_ = dist;
            return out;
        }
    }
    return out;
}

// Up to n decompressed bytes into dst; fewer only at the end of the input; -1 on error (error() has the text).
long Inflater::read(uint8_t* dst, size_t n) {
    if (state_ == S_ERROR) return -1;
    uint8_t* out = dst;
    uint8_t* const out_end = dst + n;
    uint8_t* member_from = dst;  // bytes of the current member produced in this call start here
    while (out < out_end && state_ != S_DONE) {
        switch (state_) {
            case S_HEADER:
                // zero padding behind the last member is tolerated (as gzip does)
                while (in_ < in_end_ && *in_ == 0 && members_ > 0) in_++;
                if (in_ >= in_end_) {
                    if (members_ == 0) {
                        fail("empty input is not a gzip file");
                        return -1;
                    }
                    state_ = S_DONE;
                    break;
                }
                if (!parse_header()) return -1;
                member_from = out;
                break;
            case S_BLOCK:
                if (!parse_block_header()) return -1;
                break;
            case S_STORED: {
                size_t c = stored_left_;
                if (c > (size_t)(out_end - out)) c = (size_t)(out_end - out);
                if (c > (size_t)(in_end_ - in_)) {
                    fail("truncated stored block");
                    return -1;
                }
                memcpy(out, in_, c);
                out += c;
                in_ += c;
                stored_left_ -= (uint32_t)c;
                if (stored_left_ == 0) state_ = last_block_ ? S_TRAILER : S_BLOCK;
                break;
            }
            case S_HUFFMAN:
                out = run_huffman(dst, out, out_end);
                if (state_ == S_ERROR) return -1;
                break;
            case S_TRAILER: {
                align_to_byte();
                if (in_end_ - in_ < 8) {
                    fail("truncated gzip trailer");
                    return -1;
                }
                crc_ = crc32_update(crc_, member_from, (size_t)(out - member_from));
                isize_ += (uint32_t)(out - member_from);
                member_from = out;
                const uint32_t want_crc = in_[0] | ((uint32_t)in_[1] << 8) | ((uint32_t)in_[2] << 16) | ((uint32_t)in_[3] << 24);
                const uint32_t want_size = in_[4] | ((uint32_t)in_[5] << 8) | ((uint32_t)in_[6] << 16) | ((uint32_t)in_[7] << 24);
                in_ += 8;
                if (want_crc != crc_) {
                    fail("CRC check failed (corrupt gzip data)");
                    return -1;
                }
                if (want_size != isize_) {
                    fail("length check failed (corrupt gzip data)");
                    return -1;
                }
                state_ = S_HEADER;
                break;
            }
            default: return -1;
        }
    }
    // account the bytes of the still open member, keep the last 32 KiB as history for the next call
    if (out > member_from && state_ != S_HEADER && state_ != S_DONE) {
        crc_ = crc32_update(crc_, member_from, (size_t)(out - member_from));
        isize_ += (uint32_t)(out - member_from);
    }
    const size_t produced = (size_t)(out - dst);
    if (produced >= sizeof(win_)) {
        memcpy(win_, out - sizeof(win_), sizeof(win_));
        win_len_ = sizeof(win_);
    } else if (produced) {
        const size_t keep = win_len_ + produced > sizeof(win_) ? sizeof(win_) - produced : win_len_;
        memmove(win_, win_ + (win_len_ - keep), keep);
        memcpy(win_ + keep, dst, produced);
        win_len_ = keep + produced;
    }
    return (long)produced;
}

}  // namespace csqio
