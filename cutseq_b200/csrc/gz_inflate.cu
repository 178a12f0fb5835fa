// Device gzip reader for BGZF inputs (bgzip / this library's own output): one WARP per member, all members of a
// batch in flight at once.  Replaces the host inflate of the first version for such files (xopen's gzip backends
// behind cutadapt's InputPaths in the reference, run.py:434, 751): the compressed bytes cross PCIe (~1/3 of the
// text) and no host core decodes anything.  Members are independent DEFLATE streams of at most 64 KiB with their
// compressed size in the 'BC' extra field and ISIZE in the trailer, so the host lays out source and destination
// offsets by walking the member headers, without looking at the payload.
// k_gz_inflate: the decoder of gz_inflate_core.h, its tables (3.6 KB per warp) in shared memory.
// k_gz_check:   one CTA per member reads the text back (L2 resident) - CRC-32 against the trailer, as zlib / isal do
//               behind the reference's reader, and the '\n' count per member from which the host cuts batches at
//               record boundaries (pass 1 of the file driver; pass 2 produces the text where the parse kernels expect it).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "csq_internal.h"
#include "gz_core.h"
#include "gz_inflate_core.h"

namespace {

constexpr int INF_LANES = 32;    // lanes per member (16: two members per warp - measured, no gain: the halves of a warp do not stay on one path)
constexpr int INF_GROUPS = 4;    // members per CTA

__global__ void __launch_bounds__(INF_GROUPS * INF_LANES) k_gz_inflate(const __grid_constant__ InflateParams P) {
    __shared__ gzi::Tables tabs[INF_GROUPS];
    const int g = threadIdx.x / INF_LANES;
    gzi::Group<INF_LANES> G;
    G.lane = threadIdx.x % INF_LANES;
    G.base = (threadIdx.x & 31) - G.lane;
    G.mask = (INF_LANES == 32 ? 0xFFFFFFFFu : ((1u << (INF_LANES & 31)) - 1u)) << G.base;
    const uint32_t i = blockIdx.x * INF_GROUPS + g;
    if (i >= P.n_members) return;
    const uint32_t s0 = P.moff[i], s1 = P.moff[i + 1], o0 = P.ooff[i], o1 = P.ooff[i + 1];
    uint32_t produced = 0;
    const int rc = gzi::inflate_member<INF_LANES>(P.comp + s0, s1 - s0, P.out + o0, o1 - o0, tabs[g], G, &produced);
    if (G.lane == 0 && (rc != gzi::OK || produced != o1 - o0)) atomicMin(P.status, (rc ? rc : (int)gzi::ERR_TRAILER) + 16 * (int)min(i, 0x7FFFFFu));
}

// Thread t takes the 256 bytes at 256 t of the member's text (members hold at most 64 KiB): line ends, and the CRC-32
// of its piece, folded with the "n bytes follow" operators x^(8n) mod P (crc32_combine): n = 256 q + r comes from two
// 256-entry tables.
__global__ void __launch_bounds__(256) k_gz_check(const __grid_constant__ InflateParams P) {
    __shared__ uint32_t tab[256];
    __shared__ uint32_t wcrc[8], wnl[8];
    const uint32_t i = blockIdx.x, t = threadIdx.x;
    tab[t] = P.crc_tables[t];
    __syncthreads();
    const uint32_t o0 = P.ooff[i], len = P.ooff[i + 1] - o0;
    const uint8_t* __restrict__ text = P.out + o0;
    const uint32_t lo = min(t * 256u, len), hi = min(lo + 256u, len);
    uint32_t crc = 0, nl = 0;
    if (hi > lo) {
        crc = 0xFFFFFFFFu;
        uint32_t k = lo;
        for (; k < hi && ((o0 + k) & 3u); k++) {
            const uint32_t c = text[k];
            nl += c == '\n';
            crc = tab[(crc ^ c) & 0xFFu] ^ (crc >> 8);
        }
        for (; k + 4 <= hi; k += 4) {
            uint32_t w = *reinterpret_cast<const uint32_t*>(text + k);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t c = w & 0xFFu;
                w >>= 8;
                nl += c == '\n';
                crc = tab[(crc ^ c) & 0xFFu] ^ (crc >> 8);
            }
        }
        for (; k < hi; k++) {
            const uint32_t c = text[k];
            nl += c == '\n';
            crc = tab[(crc ^ c) & 0xFFu] ^ (crc >> 8);
        }
        crc = ~crc;
        const uint32_t after = len - hi;
        if (after) crc = gz::crc_mulmod(gz::crc_mulmod(P.crc_tables[256 + (after >> 8)], P.crc_tables[512 + (after & 255u)]), crc);
    }
    for (int o = 16; o > 0; o >>= 1) {
        crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
        nl += __shfl_xor_sync(0xffffffffu, nl, o);
    }
    if ((t & 31u) == 0) {
        wcrc[t >> 5] = crc;
        wnl[t >> 5] = nl;
    }
    __syncthreads();
    if (t == 0) {
        uint32_t all = 0, lines = 0;
        for (int w = 0; w < 8; w++) {
            all ^= wcrc[w];
            lines += wnl[w];
        }
        const uint8_t* tr = P.comp + P.moff[i + 1] - 8;
        const uint32_t want = tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24);
        if (all != want) atomicMin(P.status, 7 + 16 * (int)min(i, 0x7FFFFFu));
        // bit 31: the member's text does not end in a line end (the host needs that for the last member of a file)
        if (P.lines) P.lines[i] = lines | ((len && text[len - 1] != '\n') ? 0x80000000u : 0u);
    }
}

}  // namespace

// [0, 256): CRC-32 byte table; [256, 512): x^(8 * 256 * q) mod P; [512, 768): x^(8 * r) mod P
void csq_gz_crc_check_tables(uint32_t* t /*[768]*/) {
    gz::crc_make_table(t);
    for (int q = 0; q < 256; q++) t[256 + q] = gz::crc_xpow8((uint64_t)q * 256);
    for (int r = 0; r < 256; r++) t[512 + r] = gz::crc_xpow8((uint64_t)r);
}

cudaError_t csq_launch_inflate(const InflateParams& p, cudaStream_t stream) {
    if (p.n_members == 0) return cudaSuccess;
    k_gz_inflate<<<(p.n_members + INF_GROUPS - 1) / INF_GROUPS, INF_GROUPS * INF_LANES, 0, stream>>>(p);
    k_gz_check<<<p.n_members, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// host twin for the CPU tests: the same decoder, member by member (BGZF framing walked as the file driver does)
extern "C" int csq_gz_inflate_host(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t* out_n, uint64_t* lines) {
    if (!src || !dst || !out_n) return CSQ_ERR_INVALID;
    std::vector<gzi::Tables> tab(1);
    uint32_t crc_tab[256];
    gz::crc_make_table(crc_tab);
    std::vector<uint8_t> padded;
    uint64_t pos = 0, out = 0, nl = 0;
    while (pos < n) {
        if (n - pos < 18 || src[pos] != 0x1f || src[pos + 1] != 0x8b || !(src[pos + 3] & 4) || src[pos + 12] != 'B' || src[pos + 13] != 'C') return CSQ_ERR_FORMAT;
        const uint32_t size = (uint32_t)(src[pos + 16] | (src[pos + 17] << 8)) + 1u;
        if (pos + size > n) return CSQ_ERR_FORMAT;
        const uint32_t isize = src[pos + size - 4] | (src[pos + size - 3] << 8) | (src[pos + size - 2] << 16) | ((uint32_t)src[pos + size - 1] << 24);
        if (out + isize > cap) return CSQ_ERR_CAPACITY;
        padded.assign(src + pos, src + pos + size);
        padded.resize(size + 16, 0);  // the decoder may look a few bytes behind the member
        uint32_t produced = 0;
        gzi::Group<1> one = {0, 1u, 0};
        const int rc = gzi::inflate_member<1>(padded.data(), size, dst + out, isize, tab[0], one, &produced);
        if (rc != gzi::OK) return -100 - rc;
        uint32_t crc = 0xFFFFFFFFu;
        for (uint32_t q = 0; q < produced; q++) {
            nl += dst[out + q] == '\n';
            crc = crc_tab[(crc ^ dst[out + q]) & 0xFFu] ^ (crc >> 8);
        }
        const uint8_t* tr = src + pos + size - 8;
        if (~crc != ((uint32_t)tr[0] | ((uint32_t)tr[1] << 8) | ((uint32_t)tr[2] << 16) | ((uint32_t)tr[3] << 24))) return -107;  // CRC-32 mismatch
        out += produced;
        pos += size;
    }
    *out_n = out;
    if (lines) *lines = nl;
    return 0;
}
