// Parallel decoder for ORDINARY gzip files - one long DEFLATE stream per member, the form the sequencers' software and
// `gzip` / `pigz` write and the one the reference's users feed it (xopen behind cutadapt's InputPaths, run.py:434, 751).
// BGZF inputs are inflated on the GPU (gz_inflate.cu); a single stream has no member boundaries to split at, and the
// serial decoder (inflate.cpp, ~0.3 - 0.45 GB/s of text per stream) caps a 2 x 150 run at ~1.2 M pairs/s on any number
// of GPUs.  Here the stream is decoded by several threads at once, in two passes (the scheme of pugz - Kerbiriou &
// Chikhi, "Parallel decompression of gzip-compressed files and random access to DNA sequences", 2019 - and rapidgzip):
//
//   1. the compressed file is cut into spans; for every span but the first a thread SEARCHES the first DEFLATE block
//      that starts behind the cut: a dynamic-Huffman block header whose code-length code and both codes are complete
//      and whose first symbols decode to text;
//   2. every span is decoded from its block start WITHOUT knowing the 32 KiB of text in front of it: the output is
//      kept as 16-bit symbols, and a back-reference that reaches in front of the span yields MARKERS (0x8000 | place
//      in the unknown window), which later copies move around like any other symbol;
//   3. in file order (cheap: 32 KiB per span) the window behind every span becomes known, then all spans replace their
//      markers at once and deliver bytes.
//
// Nothing rests on the search heuristic: a span is only accepted if the decoder of the span in front of it arrives
// EXACTLY at its start bit at a block boundary - DEFLATE decoding is deterministic, so from there on both would do the
// same.  A start that is not confirmed (a false positive, or text that is not ASCII so that no start is found) costs
// speed, never correctness: the predecessor's decoder simply goes on through that span.  CRC-32 and ISIZE of every
// member are verified (per-span CRCs combined in order).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <zlib.h>  // crc32_combine()

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_io.h"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace csqio {

namespace {

constexpr int LBITS = 11, DBITS = 9, MAXBITS = 15;
constexpr uint32_t E_LIT = 0x80000000u, E_EOB = 0x40000000u, E_BAD = 0xFFFFFFFFu;
constexpr size_t WIN = 32768;

inline uint32_t entry_litlen(uint32_t s) {
    if (s < 256u) return E_LIT | (s << 8);
    if (s == 256u) return E_EOB;
    const uint32_t ls = s - 257u;
    if (ls >= 29u) return E_BAD;
    if (ls < 8u) return (3u + ls) << 8;
    if (ls == 28u) return 258u << 8;
    const uint32_t x = (ls >> 2) - 1u;
    return ((((4u + (ls & 3u)) << x) + 3u) << 8) | (x << 4);
}
inline uint32_t entry_dist(uint32_t ds) {
    if (ds >= 30u) return E_BAD;
    if (ds < 4u) return (1u + ds) << 8;
    const uint32_t x = (ds >> 1) - 1u;
    return ((((2u + (ds & 1u)) << x) + 1u) << 8) | (x << 4);
}

struct Code {  // canonical Huffman code: direct table for short codes, counts + symbols in code order for the rest
    uint16_t count[MAXBITS + 1];
    uint16_t symbol[288];
    std::vector<uint32_t> lut;  // code length [3:0] | extra bits [7:4] | base [23:8] | flags; 0: longer than the index
    int bits = 0;
};

// -> 0 complete, > 0 incomplete, < 0 over-subscribed.  kind 0: plain symbols, 1: literal/length, 2: distance
int build_code(Code& c, const uint8_t* length, int n, int bits, int kind) {
    c.bits = bits;
    c.lut.assign((size_t)1 << bits, 0u);
    for (int l = 0; l <= MAXBITS; l++) c.count[l] = 0;
    for (int s = 0; s < n; s++) c.count[length[s]]++;
    if (c.count[0] == n) return 0;
    int left = 1;
    for (int l = 1; l <= MAXBITS; l++) {
        left <<= 1;
        left -= c.count[l];
        if (left < 0) return left;
    }
    uint16_t offs[MAXBITS + 2];
    uint32_t first[MAXBITS + 2];
    offs[1] = 0;
    first[1] = 0;
    for (int l = 1; l <= MAXBITS; l++) {
        offs[l + 1] = (uint16_t)(offs[l] + c.count[l]);
        first[l + 1] = (first[l] + c.count[l]) << 1;
    }
    uint16_t next[MAXBITS + 2];
    memcpy(next, offs, sizeof(next));
    for (int s = 0; s < n; s++)
        if (length[s]) c.symbol[next[length[s]]++] = (uint16_t)s;
    for (int l = 1; l <= bits; l++)
        for (int k = 0; k < c.count[l]; k++) {
            const uint32_t s = c.symbol[offs[l] + k];
            const uint32_t meaning = kind == 0 ? (s << 8) : kind == 1 ? entry_litlen(s) : entry_dist(s);
            if (meaning == E_BAD) continue;
            uint32_t code = first[l] + (uint32_t)k, rev = 0;
            for (int b = 0; b < l; b++) rev |= ((code >> b) & 1u) << (l - 1 - b);
            for (uint32_t e = rev; e < (1u << bits); e += 1u << l) c.lut[e] = meaning | (uint32_t)l;
        }
    return left;
}

struct Bits {
    const uint8_t *data, *end, *in;
    uint64_t buf = 0;
    int cnt = 0;
    void seek(const uint8_t* d, size_t n, size_t bit) {
        data = d;
        end = d + n;
        in = d + (bit >> 3);
        buf = 0;
        cnt = 0;
        refill();
        const int skip = (int)(bit & 7);
        buf >>= skip;
        cnt -= skip;
    }
    inline void refill() {  // >= 56 bits afterwards (zeros behind the end)
        if (in + 8 <= end) {
            uint64_t w;
            memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56) {
                if (in < end) buf |= (uint64_t)(*in) << cnt;
                in++;  // may run behind the end: pos() then reports a position behind it, the callers check
                cnt += 8;
            }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    inline void drop(int n) {
        buf >>= n;
        cnt -= n;
    }
    inline uint32_t take(int n) {
        const uint32_t v = peek(n);
        drop(n);
        return v;
    }
    size_t pos() const { return (size_t)(in - data) * 8 - (size_t)cnt; }  // bit position of the next unread bit
};

inline uint32_t decode_entry(Bits& b, const Code& c, int kind) {
    const uint32_t e = c.lut[b.peek(c.bits)];
    if (e) {
        b.drop((int)(e & 15u));
        return e & ~15u;
    }
    int code = 0, first = 0, index = 0;
    uint64_t bits = b.buf;
    for (int len = 1; len <= MAXBITS; len++) {
        code |= (int)(bits & 1u);
        bits >>= 1;
        const int cnt = c.count[len];
        if (code - cnt < first) {
            b.drop(len);
            const uint32_t s = c.symbol[index + (code - first)];
            return kind == 0 ? (s << 8) : kind == 1 ? entry_litlen(s) : entry_dist(s);
        }
        index += cnt;
        first += cnt;
        first <<= 1;
        code <<= 1;
    }
    return E_BAD;
}

// Dynamic block header at the reader's position (behind the 3 header bits) -> both codes.  false: not a valid header.
bool read_dynamic(Bits& b, Code& lit, Code& dist, Code& pre) {
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    b.refill();
    const int nlen = (int)b.take(5) + 257, ndist = (int)b.take(5) + 1, ncode = (int)b.take(4) + 4;
    if (nlen > 286 || ndist > 30) return false;
    uint8_t lengths[320];
    memset(lengths, 0, 19);
    for (int i = 0; i < ncode; i++) {
        if ((i & 7) == 0) b.refill();
        lengths[order[i]] = (uint8_t)b.take(3);
    }
    if (build_code(pre, lengths, 19, 7, 0) != 0) return false;
    int idx = 0, prev = 0;
    while (idx < nlen + ndist) {
        b.refill();
        const uint32_t e = decode_entry(b, pre, 0);
        if (e == E_BAD) return false;
        const int sym = (int)(e >> 8);
        int rep = 1, val = sym;
        if (sym == 16) {
            if (idx == 0) return false;
            val = prev;
            rep = 3 + (int)b.take(2);
        } else if (sym == 17) {
            val = 0;
            rep = 3 + (int)b.take(3);
        } else if (sym == 18) {
            val = 0;
            rep = 11 + (int)b.take(7);
        }
        if (idx + rep > nlen + ndist) return false;
        while (rep--) lengths[idx++] = (uint8_t)val;
        prev = val;
    }
    if (lengths[256] == 0) return false;
    int err = build_code(lit, lengths, nlen, LBITS, 1);
    if (err < 0 || (err > 0 && nlen - lit.count[0] != 1)) return false;
    err = build_code(dist, lengths + nlen, ndist, DBITS, 2);
    if (err < 0 || (err > 0 && ndist - dist.count[0] != 1)) return false;
    return true;
}

void fixed_codes(Code& lit, Code& dist) {
    uint8_t lengths[320];
    for (int s = 0; s < 288; s++) lengths[s] = (uint8_t)(s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8);
    build_code(lit, lengths, 288, LBITS, 1);
    for (int s = 0; s < 30; s++) lengths[s] = 5;
    build_code(dist, lengths, 30, DBITS, 2);
}

inline bool texty(uint32_t c) { return (c >= 32 && c < 127) || c == '\n' || c == '\r' || c == '\t'; }

// Is there a dynamic block at this bit that decodes to text?  (The search of pass 1.)
bool plausible_block(const uint8_t* data, size_t n, size_t bit, Code& lit, Code& dist, Code& pre) {
    Bits b;
    b.seek(data, n, bit);
    const uint32_t h = b.take(3);
    if (h != 4u) return false;  // BFINAL = 0, BTYPE = 2
    // the code-length code must be complete: checked on the raw bits before anything is built
    {
        const uint64_t v = b.buf;
        const int nlen = (int)(v & 31) + 257, ndist = (int)((v >> 5) & 31) + 1, ncode = (int)((v >> 10) & 15) + 4;
        if (nlen > 286 || ndist > 30) return false;
        const int avail = std::min(ncode, (b.cnt - 14) / 3);
        uint32_t kraft = 0;
        for (int i = 0; i < avail; i++) {
            const int l = (int)((v >> (14 + 3 * i)) & 7);
            if (l) kraft += 128u >> l;
        }
        if (kraft > 128u || (avail == ncode && kraft != 128u)) return false;
    }
    if (!read_dynamic(b, lit, dist, pre)) return false;
    size_t produced = 0;
    for (int k = 0; k < 4096; k++) {
        b.refill();
        uint32_t e = decode_entry(b, lit, 1);
        if (e >= 0x01000000u) {
            if (e == E_BAD) return false;
            if (e & E_EOB) return produced >= 64;  // a whole (if short) block of text
            if (!texty((e >> 8) & 0xFFu)) return false;
            produced++;
            continue;
        }
        const int x = (int)(e >> 4 & 15u);
        const uint32_t len = (e >> 8) + b.take(x);
        b.refill();
        e = decode_entry(b, dist, 2);
        if (e == E_BAD) return false;
        b.drop((int)(e >> 4 & 15u));
        produced += len;
        if (b.pos() > n * 8) return false;
    }
    return true;
}

size_t find_block(const uint8_t* data, size_t n, size_t from_bit, size_t to_bit) {
    Code lit, dist, pre;
    for (size_t bit = from_bit; bit < to_bit; bit++) {
        // cheap test on three bits before anything else
        const uint32_t three = ((uint32_t)data[bit >> 3] | ((uint32_t)((bit >> 3) + 1 < n ? data[(bit >> 3) + 1] : 0) << 8)) >> (bit & 7) & 7u;
        if (three != 4u) continue;
        if (plausible_block(data, n, bit, lit, dist, pre)) return bit;
    }
    return SIZE_MAX;
}

// Big working buffers: mapped (never zero-filled by us, grown in place by mremap), huge pages where the system gives
// them, and kept for the next round - the page faults of fresh memory were 2/3 of the first version's run time.
template <typename T>
struct MapBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    MapBuf() = default;
    MapBuf(const MapBuf&) = delete;
    MapBuf& operator=(const MapBuf&) = delete;
    MapBuf(MapBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    MapBuf& operator=(MapBuf&& o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            cap = o.cap;
            o.p = nullptr;
            o.cap = 0;
        }
        return *this;
    }
    ~MapBuf() { release(); }
    void release() {
        if (p) munmap(p, cap * sizeof(T));
        p = nullptr;
        cap = 0;
    }
    bool reserve(size_t elems) {  // keeps the contents
        if (elems <= cap) return true;
        const size_t bytes = ((elems * sizeof(T)) + ((2u << 20) - 1)) & ~(size_t)((2u << 20) - 1);
        void* q;
        if (p) {
            q = mremap(p, cap * sizeof(T), bytes, MREMAP_MAYMOVE);
        } else {
            q = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        }
        if (q == MAP_FAILED) return false;
        madvise(q, bytes, MADV_HUGEPAGE);
        p = (T*)q;
        cap = bytes / sizeof(T);
        return true;
    }
};

struct MemberEnd {
    size_t out_off;  // bytes of the span in front of the member's end
    uint32_t crc, isize;
};

struct Span {
    size_t start_bit = 0;       // where decoding starts (a block boundary, or a gzip header when header_first)
    bool header_first = false;  // the span starts at a gzip member header (byte aligned)
    bool known_window = false;  // sym[0 .. WIN) holds real bytes (the first span of a round)
    bool found = true;          // pass 1 found a start (else the span is empty and its predecessor covers it)
    MapBuf<uint16_t> sym;       // WIN window places, then the output
    size_t n_out = 0;
    size_t end_bit = 0;         // block boundary (or end of the data) where decoding stopped
    bool at_eof = false;        // the data ended with a complete member
    std::vector<MemberEnd> members;
    std::string error;
    MapBuf<uint8_t> bytes;           // resolved output (n_out bytes)
    std::vector<uint32_t> seg_crc;   // CRC-32 of the segments between member ends (one more than members)
};

bool parse_gzip_header(const uint8_t* data, size_t n, size_t& byte) {
    if (byte + 18 > n || data[byte] != 0x1f || data[byte + 1] != 0x8b || data[byte + 2] != 8) return false;
    const uint32_t flg = data[byte + 3];
    size_t p = byte + 10;
    if (flg & 4u) {
        if (p + 2 > n) return false;
        p += 2 + (size_t)(data[p] | (data[p + 1] << 8));
    }
    if (flg & 8u) {
        while (p < n && data[p]) p++;
        p++;
    }
    if (flg & 16u) {
        while (p < n && data[p]) p++;
        p++;
    }
    if (flg & 2u) p += 2;
    if (p >= n) return false;
    byte = p;
    return true;
}

// Pass 2: decode from the span's start up to the first block boundary at or behind stop_bit (or the end of the data).
void decode_span(const uint8_t* data, size_t n, Span& sp, size_t stop_bit) {
    const size_t total_bits = n * 8;
    if (!sp.sym.reserve(WIN + ((size_t)4 << 20) + 5 * (size_t)std::min<size_t>(n, (stop_bit == SIZE_MAX ? n * 8 : stop_bit) / 8 - std::min(sp.start_bit / 8, n)))) {
        sp.error = "out of memory";
        return;
    }
    if (!sp.known_window)
        for (size_t j = 0; j < WIN; j++) sp.sym.p[j] = (uint16_t)(0x8000u | j);
    size_t cap = sp.sym.cap;
    uint16_t* out = sp.sym.p;
    size_t o = WIN;
    bool oom = false;
    auto grow = [&](size_t need) {
        if (o + need + 320 <= cap) return;
        if (!sp.sym.reserve(std::max(cap + cap / 2, o + need + 320))) {
            oom = true;
            return;
        }
        cap = sp.sym.cap;
        out = sp.sym.p;
    };
    Bits b;
    size_t start = sp.start_bit;
    if (sp.header_first) {
        size_t byte = start >> 3;
        if (!parse_gzip_header(data, n, byte)) {
            sp.error = "not a gzip header";
            return;
        }
        start = byte * 8;
    }
    b.seek(data, n, start);
    Code lit, dist, pre;
    bool first_block = true;
    for (;;) {
        const size_t at = b.pos();
        if (!first_block && at >= stop_bit) {
            sp.end_bit = at;
            break;
        }
        first_block = false;
        if (at + 3 > total_bits) {
            sp.error = "the DEFLATE stream ends in the middle of a member";
            return;
        }
        b.refill();
        const uint32_t last = b.take(1), type = b.take(2);
        if (type == 0) {
            b.drop(b.cnt & 7);
            b.refill();
            const uint32_t len = b.take(16), nlen = b.take(16);
            if ((len ^ 0xFFFFu) != nlen) {
                sp.error = "stored block with inconsistent length";
                return;
            }
            const size_t byte = b.pos() >> 3;
            if (byte + len > n) {
                sp.error = "stored block runs past the end of the file";
                return;
            }
            grow(len);
            if (oom) {
                sp.error = "out of memory";
                return;
            }
            for (uint32_t i = 0; i < len; i++) out[o + i] = data[byte + i];
            o += len;
            b.seek(data, n, (byte + len) * 8);
        } else if (type == 3) {
            sp.error = "invalid block type";
            return;
        } else {
            if (type == 1) {
                fixed_codes(lit, dist);
            } else if (!read_dynamic(b, lit, dist, pre)) {
                sp.error = "invalid dynamic block header";
                return;
            }
            const uint32_t* const llut = lit.lut.data();
            const uint32_t* const dlut = dist.lut.data();
            for (;;) {
                if (o + 640 > cap) {
                    grow(258);
                    if (oom) {
                        sp.error = "out of memory";
                        return;
                    }
                }
                b.refill();
                if (b.in > b.end + 16) {
                    sp.error = "the DEFLATE stream ends in the middle of a block";
                    return;
                }
                uint32_t e = llut[(uint32_t)b.buf & ((1u << LBITS) - 1u)];
                if (e) {
                    b.drop((int)(e & 15u));
                    e &= ~15u;
                } else {
                    e = decode_entry(b, lit, 1);
                }
                if (e >= 0x01000000u) {
                    if (e == E_BAD) {
                        sp.error = "invalid literal/length code";
                        return;
                    }
                    if (e & E_EOB) break;
                    out[o++] = (uint16_t)((e >> 8) & 0xFFu);
                    // literal runs: up to two more without a refill (15 + 11 + 11 bits of the 56 on hand)
                    e = llut[(uint32_t)b.buf & ((1u << LBITS) - 1u)];
                    if (e >= 0x80000000u) {
                        b.drop((int)(e & 15u));
                        out[o++] = (uint16_t)((e >> 8) & 0xFFu);
                        e = llut[(uint32_t)b.buf & ((1u << LBITS) - 1u)];
                        if (e >= 0x80000000u) {
                            b.drop((int)(e & 15u));
                            out[o++] = (uint16_t)((e >> 8) & 0xFFu);
                        }
                    }
                    continue;
                }
                // a match: length (<= 15 + 5 bits) and distance (<= 15 + 13) come out of the 56 bits on hand
                const int x = (int)(e >> 4 & 15u);
                const uint32_t len = (e >> 8) + b.take(x);
                e = dlut[(uint32_t)b.buf & ((1u << DBITS) - 1u)];
                if (e) {
                    b.drop((int)(e & 15u));
                    e &= ~15u;
                } else {
                    e = decode_entry(b, dist, 2);
                }
                if (e == E_BAD) {
                    sp.error = "invalid distance code";
                    return;
                }
                const int y = (int)(e >> 4 & 15u);
                const uint32_t d = (e >> 8) + b.take(y);
                if (d > o) {  // (o counts the window places too: a reference may reach WIN back at most)
                    sp.error = "distance reaches in front of the window";
                    return;
                }
                const uint16_t* s = out + o - d;
                uint16_t* t = out + o;
                if (d >= 8) {  // 8 symbols per step; the copy may run up to 7 symbols past the match (slack behind o)
                    for (uint32_t i = 0; i < len; i += 8) {
                        uint64_t lo, hi;
                        memcpy(&lo, s + i, 8);
                        memcpy(&hi, s + i + 4, 8);
                        memcpy(t + i, &lo, 8);
                        memcpy(t + i + 4, &hi, 8);
                    }
                } else {
                    for (uint32_t i = 0; i < len; i++) t[i] = s[i];
                }
                o += len;
            }
        }
        if (last) {  // trailer of this member, then the next member's header or the end of the data
            b.drop(b.cnt & 7);
            size_t byte = b.pos() >> 3;
            if (byte + 8 > n) {
                sp.error = "gzip trailer missing";
                return;
            }
            MemberEnd me;
            me.out_off = o - WIN;
            me.crc = (uint32_t)data[byte] | ((uint32_t)data[byte + 1] << 8) | ((uint32_t)data[byte + 2] << 16) | ((uint32_t)data[byte + 3] << 24);
            me.isize = (uint32_t)data[byte + 4] | ((uint32_t)data[byte + 5] << 8) | ((uint32_t)data[byte + 6] << 16) | ((uint32_t)data[byte + 7] << 24);
            sp.members.push_back(me);
            byte += 8;
            while (byte < n && data[byte] == 0) byte++;  // zero padding between / behind members (gzip accepts it)
            if (byte >= n) {
                sp.end_bit = n * 8;
                sp.at_eof = true;
                break;
            }
            if (!parse_gzip_header(data, n, byte)) {
                sp.error = "trailing garbage behind a gzip member";
                return;
            }
            b.seek(data, n, byte * 8);
        }
    }
    sp.n_out = o - WIN;
}

}  // namespace

// ---- the round pipeline ----------------------------------------------------------------------------------------------
struct ParallelInflater::Impl {
    const uint8_t* data = nullptr;
    size_t n = 0;
    int threads = 4;
    size_t span_bytes = 2u << 20;  // compressed bytes per thread and round
    // decoder position
    size_t next_bit = 0;        // where the next round starts (exact)
    bool at_header = true;      // ... at a gzip header
    bool finished = false;
    std::vector<uint8_t> window;  // the 32 KiB in front of next_bit (shorter at the start of a member)
    uint32_t crc_run = 0;         // CRC-32 of the current member so far
    uint64_t len_run = 0;         // its length so far
    std::string error;
    // rounds: one being consumed, one being produced
    struct Piece {
        MapBuf<uint8_t> buf;
        size_t n = 0;
    };
    struct Round {
        std::vector<Piece> pieces;
        bool last = false;
        std::string error;
    };
    // working buffers that go round: symbol buffers between the rounds, byte buffers between producer and consumer
    std::mutex pool_m;
    std::vector<MapBuf<uint16_t>> sym_pool;
    std::vector<MapBuf<uint8_t>> byte_pool;
    template <typename T>
    static MapBuf<T> take(std::mutex& m, std::vector<MapBuf<T>>& pool) {
        std::lock_guard<std::mutex> g(m);
        if (pool.empty()) return MapBuf<T>();
        // the largest one first: spans are of similar size
        size_t best = 0;
        for (size_t i = 1; i < pool.size(); i++)
            if (pool[i].cap > pool[best].cap) best = i;
        MapBuf<T> b = std::move(pool[best]);
        pool.erase(pool.begin() + (long)best);
        return b;
    }
    template <typename T>
    static void give(std::mutex& m, std::vector<MapBuf<T>>& pool, MapBuf<T>&& b) {
        if (!b.p) return;
        std::lock_guard<std::mutex> g(m);
        pool.push_back(std::move(b));
    }
    std::unique_ptr<Round> ready;
    bool producing = false;
    std::thread producer;
    std::mutex m;
    std::condition_variable cv;
    // consumer position
    std::unique_ptr<Round> cur;
    size_t cur_piece = 0, cur_off = 0;
    bool eof = false;

    void produce(Round& r);
    void start_round();
};

void ParallelInflater::Impl::produce(Round& r) {
    const bool trace = getenv("CSQ_PINFLATE_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto stamp = [&](const char* what) {
        if (trace) fprintf(stderr, "[pinflate] %8.1f ms  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
    };
    const size_t total_bits = n * 8;
    const int T = threads;
    std::vector<Span> spans((size_t)T);
    // the spans of this round: nominal starts at multiples of span_bytes behind next_bit
    const size_t base_byte = next_bit >> 3;
    std::vector<size_t> nominal((size_t)T + 1);
    for (int k = 0; k <= T; k++) nominal[(size_t)k] = std::min(n, base_byte + (size_t)k * span_bytes) * 8;
    spans[0].start_bit = next_bit;
    spans[0].header_first = at_header;
    spans[0].known_window = true;
    for (int k = 0; k < T; k++) spans[(size_t)k].sym = take(pool_m, sym_pool);
    if (!spans[0].sym.reserve(WIN + ((size_t)4 << 20))) {
        r.error = "out of memory";
        return;
    }
    memset(spans[0].sym.p, 0, (WIN - window.size()) * 2);
    for (size_t j = 0; j < window.size(); j++) spans[0].sym.p[WIN - window.size() + j] = window[j];
    // pass 1 + 2, one thread per span: find the start, then decode up to the next span's start.  A span waits for
    // its successor's start (the stop position) before it decodes.
    std::vector<size_t> start((size_t)T + 1, SIZE_MAX);
    std::vector<char> have((size_t)T + 1, 0);
    std::mutex sm;
    std::condition_variable scv;
    start[0] = next_bit;
    have[0] = 1;
    start[(size_t)T] = nominal[(size_t)T];  // the last span stops at the first block boundary behind the round
    have[(size_t)T] = 1;
    auto work = [&](int k) {
        Span& sp = spans[(size_t)k];
        if (k > 0) {
            size_t found = SIZE_MAX;
            if (nominal[(size_t)k] < total_bits) found = find_block(data, n, nominal[(size_t)k], nominal[(size_t)k + 1]);
            {
                std::lock_guard<std::mutex> g(sm);
                start[(size_t)k] = found;
                have[(size_t)k] = 1;
            }
            scv.notify_all();
            if (found == SIZE_MAX) {
                sp.found = false;
                return;
            }
            sp.start_bit = found;
            if (trace) fprintf(stderr, "[pinflate] span %d: start found %.1f ms\n", k, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
        // stop at the start of the next span that has one
        size_t stop = SIZE_MAX;
        {
            std::unique_lock<std::mutex> g(sm);
            for (int q = k + 1; q <= T; q++) {
                scv.wait(g, [&] { return have[(size_t)q] != 0; });
                if (start[(size_t)q] != SIZE_MAX) {
                    stop = start[(size_t)q];
                    break;
                }
            }
        }
        decode_span(data, n, sp, stop);
        if (trace) fprintf(stderr, "[pinflate] span %d: decoded %zu bytes, %.1f ms\n", k, sp.n_out, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    {
        std::vector<std::thread> ts;
        for (int k = 1; k < T; k++) ts.emplace_back(work, k);
        work(0);
        for (auto& t : ts) t.join();
    }
    stamp("spans decoded");
    // pass 3a, in file order: which spans are confirmed, and the window behind each of them
    std::vector<int> chain;  // the spans that make up the output, in order
    {
        int k = 0;
        for (;;) {
            Span& sp = spans[(size_t)k];
            if (!sp.error.empty()) {
                if (k == 0 || chain.empty()) {
                    r.error = sp.error;
                    return;
                }
                // cannot happen for a confirmed span of a valid file: its predecessor arrived here
                r.error = sp.error;
                return;
            }
            chain.push_back(k);
            if (sp.at_eof) break;
            // the successor whose start equals this span's end, if any; spans in between were skipped over
            const size_t here = sp.end_bit;
            int nxt = -1;
            for (int q = 1; q < T; q++)
                if (spans[(size_t)q].found && spans[(size_t)q].start_bit == here) {
                    nxt = q;
                    break;
                }
            if (nxt < 0) {
                // Either this was the last span of the round, or a later start was a false positive that this span
                // ran past.  Go on from here with a fresh span up to the next start behind end_bit (or the round's end).
                size_t stop = nominal[(size_t)T];
                bool more = false;
                for (int q = 1; q < T; q++)
                    if (spans[(size_t)q].found && spans[(size_t)q].start_bit > here) {
                        stop = spans[(size_t)q].start_bit;
                        more = true;
                        break;
                    }
                if (!more && here >= nominal[(size_t)T]) break;  // the round is complete
                // (rare) serial continuation: decode [end_bit, stop) with markers like any other span
                spans.emplace_back();
                Span& ext = spans.back();
                ext.sym = take(pool_m, sym_pool);
                ext.start_bit = here;
                decode_span(data, n, ext, stop);
                k = (int)spans.size() - 1;
                continue;
            }
            k = nxt;
        }
    }
    stamp("chain built");
    // pass 3b: windows in order (serial, 32 KiB each), then every span resolves its symbols (parallel)
    std::vector<std::vector<uint8_t>> win_before(chain.size());
    {
        std::vector<uint8_t> w(WIN, 0);
        memcpy(w.data() + WIN - window.size(), window.data(), window.size());
        for (size_t c = 0; c < chain.size(); c++) {
            Span& sp = spans[(size_t)chain[c]];
            win_before[c] = w;
            // the window behind this span: the last WIN bytes of (w ++ output)
            const size_t no = sp.n_out;
            std::vector<uint8_t> nw(WIN);
            for (size_t j = 0; j < WIN; j++) {
                // place j of the new window = position (no - WIN + j) of the output, or of w if negative
                const long p = (long)no - (long)WIN + (long)j;
                if (p < 0) {
                    nw[j] = w[(size_t)((long)WIN + p)];
                } else {
                    const uint16_t s = sp.sym.p[WIN + (size_t)p];
                    nw[j] = (s & 0x8000u) ? w[s & 0x7FFFu] : (uint8_t)s;
                }
            }
            w.swap(nw);
        }
        window = w;  // (members shorter than the window: the bytes in front are never referenced by a valid stream)
    }
    {
        std::atomic<size_t> next{0};
        auto resolve = [&] {
            for (;;) {
                const size_t c = next.fetch_add(1);
                if (c >= chain.size()) return;
                Span& sp = spans[(size_t)chain[c]];
                const std::vector<uint8_t>& w = win_before[c];
                sp.bytes = take(pool_m, byte_pool);
                if (!sp.bytes.reserve(sp.n_out + 64)) {
                    sp.error = "out of memory";
                    continue;
                }
                const uint16_t* s = sp.sym.p + WIN;
                uint8_t* d = sp.bytes.p;
                size_t i = 0;
#if defined(__SSE2__)
                // 16 symbols per step; markers (bit 15) are rare behind the first 32 KiB of a span
                for (; i + 16 <= sp.n_out; i += 16) {
                    const __m128i a = _mm_loadu_si128((const __m128i*)(s + i)), b2 = _mm_loadu_si128((const __m128i*)(s + i + 8));
                    if (_mm_movemask_epi8(_mm_or_si128(a, b2)) & 0xAAAA) {
                        for (size_t q = i; q < i + 16; q++) d[q] = (s[q] & 0x8000u) ? w[s[q] & 0x7FFFu] : (uint8_t)s[q];
                    } else {
                        _mm_storeu_si128((__m128i*)(d + i), _mm_packus_epi16(a, b2));
                    }
                }
#endif
                for (; i < sp.n_out; i++) d[i] = (s[i] & 0x8000u) ? w[s[i] & 0x7FFFu] : (uint8_t)s[i];
                give(pool_m, sym_pool, std::move(sp.sym));
                // CRC-32 of the pieces between member ends
                size_t from = 0;
                for (size_t q = 0; q <= sp.members.size(); q++) {
                    const size_t to = q < sp.members.size() ? sp.members[q].out_off : sp.n_out;
                    sp.seg_crc.push_back(crc32_fast(0u, d + from, to - from));
                    from = to;
                }
            }
        };
        std::vector<std::thread> ts;
        for (int t = 1; t < T; t++) ts.emplace_back(resolve);
        resolve();
        for (auto& t : ts) t.join();
    }
    stamp("resolved");
    // member checks in order
    for (size_t c = 0; c < chain.size(); c++) {
        Span& sp = spans[(size_t)chain[c]];
        size_t from = 0;
        for (size_t q = 0; q <= sp.members.size(); q++) {
            const size_t to = q < sp.members.size() ? sp.members[q].out_off : sp.n_out;
            crc_run = (uint32_t)crc32_combine(crc_run, sp.seg_crc[q], (z_off_t)(to - from));
            len_run += to - from;
            if (q < sp.members.size()) {
                if (crc_run != sp.members[q].crc) {
                    r.error = "CRC-32 mismatch in a gzip member";
                    return;
                }
                if ((uint32_t)len_run != sp.members[q].isize) {
                    r.error = "length mismatch (ISIZE) in a gzip member";
                    return;
                }
                crc_run = 0;
                len_run = 0;
            }
            from = to;
        }
        if (!sp.error.empty()) {
            r.error = sp.error;
            return;
        }
        Piece pc;
        pc.buf = std::move(sp.bytes);
        pc.n = sp.n_out;
        r.pieces.push_back(std::move(pc));
    }
    for (Span& sp : spans) give(pool_m, sym_pool, std::move(sp.sym));  // (spans that were skipped over)
    Span& lastsp = spans[(size_t)chain.back()];
    next_bit = lastsp.end_bit;
    at_header = false;
    if (lastsp.at_eof) {
        r.last = true;
        finished = true;
    }
    stamp("round done");
}

void ParallelInflater::Impl::start_round() {
    producing = true;
    producer = std::thread([this] {
        std::unique_ptr<Round> r(new Round());
        produce(*r);
        {
            std::lock_guard<std::mutex> g(m);
            ready = std::move(r);
            producing = false;
        }
        cv.notify_all();
    });
}

ParallelInflater::ParallelInflater(const uint8_t* data, size_t n, int threads) : p_(new Impl()) {
    p_->data = data;
    p_->n = n;
    p_->threads = std::max(2, threads);
    const char* env = getenv("CSQ_PINFLATE_SPAN");
    if (env && atol(env) > 0) p_->span_bytes = (size_t)atol(env);
    p_->start_round();
}

ParallelInflater::~ParallelInflater() {
    if (p_->producer.joinable()) p_->producer.join();
    delete p_;
}

const char* ParallelInflater::error() const { return p_->error.c_str(); }

long ParallelInflater::read(uint8_t* dst, size_t n) {
    Impl& I = *p_;
    size_t done = 0;
    while (done < n) {
        if (!I.cur) {
            if (I.eof) break;
            {
                std::unique_lock<std::mutex> g(I.m);
                I.cv.wait(g, [&] { return I.ready != nullptr; });
                I.cur = std::move(I.ready);
            }
            I.producer.join();
            I.cur_piece = I.cur_off = 0;
            if (!I.cur->error.empty()) {
                I.error = I.cur->error;
                return -1;
            }
            if (I.cur->last) I.eof = true;
            else I.start_round();  // the next round is decoded while this one is consumed
        }
        if (I.cur_piece >= I.cur->pieces.size()) {
            I.cur.reset();
            continue;
        }
        Impl::Piece& pc = I.cur->pieces[I.cur_piece];
        const size_t c = std::min(n - done, pc.n - I.cur_off);
        memcpy(dst + done, pc.buf.p + I.cur_off, c);
        done += c;
        I.cur_off += c;
        if (I.cur_off == pc.n) {
            Impl::give(I.pool_m, I.byte_pool, std::move(pc.buf));
            I.cur_piece++;
            I.cur_off = 0;
        }
    }
    return (long)done;
}

}  // namespace csqio

void csq_set_error(const char* msg);  // plan.cu

// test hook: a whole gzip file in memory through the parallel decoder (span size in bytes: small values exercise many
// spans and rounds on small inputs)
extern "C" int csq_pinflate_mem(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, int threads, uint64_t span_bytes,
                                uint64_t* out_n) {
    if (!src || !dst || !out_n) return CSQ_ERR_INVALID;
    char env[32];
    snprintf(env, sizeof(env), "%llu", (unsigned long long)span_bytes);
    if (span_bytes) setenv("CSQ_PINFLATE_SPAN", env, 1);
    csqio::ParallelInflater pi(src, (size_t)n, threads);
    if (span_bytes) unsetenv("CSQ_PINFLATE_SPAN");
    uint64_t total = 0;
    for (;;) {
        if (total == cap) {
            uint8_t extra;
            const long more = pi.read(&extra, 1);
            if (more < 0) break;
            if (more > 0) return CSQ_ERR_CAPACITY;
            *out_n = total;
            return 0;
        }
        const long got = pi.read(dst + total, (size_t)(cap - total));
        if (got < 0) break;
        if (got == 0) {
            *out_n = total;
            return 0;
        }
        total += (uint64_t)got;
    }
    csq_set_error(pi.error());
    return CSQ_ERR_IO;
}
