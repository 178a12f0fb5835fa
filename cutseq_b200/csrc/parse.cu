// FASTQ text -> record index, on the device.
//
// In text mode (csq_submit_text / csq_upload_text) the host ships the raw FASTQ bytes of a batch - whole
// records, cut at a record boundary by counting line ends - and these kernels do what dnaio's FASTQ
// parser (_core.pyx, behind cutadapt's InputPaths in reference run.py:434, 751) does on the CPU:
//
//   k_nl_count     reads the text once: a line-end bit mask per 16-byte chunk, line ends per 16 KiB tile
//   k_tile_scan    exclusive scan of the tile counts (one CTA)
//   k_nl_index     from the masks: byte position of every line end, in order, nl[r] = offset of the r-th '\n'
//   k_records      record i = lines 4i .. 4i+3: '@' / '+' checks, '\r' stripping, equal sequence / quality
//                  lengths, the read-length limit; writes name / sequence / quality offsets and lengths
//
// The trimming kernels then work on the text where it lies (seq == qual == name pool == the text buffer);
// nothing is copied into a packed layout.  Bound: HBM - the text is read once (plus 1/8 of it written and read
// back as masks, 16 bytes of line-end offsets per record written, and the six boundary bytes of each record
// looked at by k_records).  A malformed record is reported as the smallest (record, kind) key through `perr`.
#include <cuda_runtime.h>
#include <stdint.h>

#include "csq_internal.h"

namespace {

constexpr int TILE_THREADS = 256;
constexpr int TILE_BYTES = 16384;               // 1024 chunks of 16 bytes; thread t owns the 64 bytes at 64 t
constexpr int TILE_CHUNKS = TILE_BYTES / 16;

// bit i set <=> byte i of the 16-byte chunk is '\n'
__device__ __forceinline__ uint32_t nl_mask16(uint4 v) {
    auto nib = [](uint32_t w) {
        const uint32_t x = (__vcmpeq4(w, 0x0A0A0A0Au) >> 7) & 0x01010101u;  // bits 0, 8, 16, 24
        return ((x * 0x00204081u) >> 21) & 0xFu;                             // gathered into bits 0..3
    };
    return nib(v.x) | (nib(v.y) << 4) | (nib(v.z) << 8) | (nib(v.w) << 12);
}

// Pass 1: the text is read ONCE, fully coalesced (chunk c of the tile by thread c % 256); what survives is a
// 16-bit line-end mask per chunk (1/8 of the text) and the number of line ends per tile.
// The text buffer is 16-byte aligned and zero padded, so whole chunks never leave the allocation and padding
// never counts as a line end.
__global__ void __launch_bounds__(TILE_THREADS) k_nl_count(const uint8_t* __restrict__ text, uint64_t bytes,
                                                           uint16_t* __restrict__ masks, uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t wsum[TILE_THREADS / 32];
    const uint64_t tile_off = (uint64_t)blockIdx.x * TILE_BYTES;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < TILE_CHUNKS / TILE_THREADS; i++) {
        const uint32_t chunk = i * TILE_THREADS + threadIdx.x;
        const uint64_t off = tile_off + 16ull * chunk;
        uint32_t m = 0;
        if (off < bytes) m = nl_mask16(*reinterpret_cast<const uint4*>(text + off));
        masks[(size_t)blockIdx.x * TILE_CHUNKS + chunk] = (uint16_t)m;
        c += __popc(m);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < TILE_THREADS / 32; w++) t += wsum[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// One CTA: exclusive scan of n_tiles counts in place; total[0] = number of line ends.
__global__ void __launch_bounds__(1024) k_tile_scan(uint32_t n_tiles, uint32_t* __restrict__ tile_cnt, uint32_t* __restrict__ total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? tile_cnt[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const uint32_t before = carry + (wid ? warp_sums[wid - 1] : 0u) + (x - v);
        if (i < n_tiles) tile_cnt[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = carry;
}

// Pass 2 reads only the masks: thread t of a tile owns the 64 bytes at 64 t (four masks = one 64-bit word),
// ranks its line ends by a CTA-wide scan and writes their byte offsets, nl[r] = offset of the r-th '\n'.
__global__ void __launch_bounds__(TILE_THREADS) k_nl_index(const uint16_t* __restrict__ masks, const uint32_t* __restrict__ tile_base,
                                                           uint32_t* __restrict__ nl, uint32_t nl_cap) {
    __shared__ uint32_t wsum[TILE_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long m = reinterpret_cast<const unsigned long long*>(masks + (size_t)blockIdx.x * TILE_CHUNKS)[threadIdx.x];
    const uint32_t c = __popcll(m);
    uint32_t x = c;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    uint32_t r = tile_base[blockIdx.x] + (x - c);
    for (int w = 0; w < wid; w++) r += wsum[w];
    const uint32_t off = blockIdx.x * (uint32_t)TILE_BYTES + threadIdx.x * 64u;
    while (m) {
        const int bit = __ffsll((long long)m) - 1;
        m &= m - 1;
        if (r < nl_cap) nl[r] = off + (uint32_t)bit;
        r++;
    }
}

// kinds of a malformed record, in the order dnaio would meet them
enum { PERR_AT = 1, PERR_PLUS = 2, PERR_LEN = 3, PERR_LIMIT = 4, PERR_COUNT = 5 };

__device__ __forceinline__ void report(unsigned long long* perr, uint32_t rec, int kind) {
    atomicMin(perr, ((unsigned long long)rec << 3) | (unsigned long long)kind);
}

__global__ void __launch_bounds__(256) k_records(const ParseParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool count_ok = P.nl_total[0] == 4u * P.n;
    if (i == 0 && !count_ok) report(P.perr, 0, PERR_COUNT);
    if (i >= P.n) return;
    if (!count_ok) {  // nl[] is not fully defined: every record becomes an empty one, the host rejects the batch
        P.name_off[i] = P.name_end[i] = P.seq_off[i] = P.seq_len[i] = P.qual_off[i] = 0;
        return;
    }
    const uint4 e = reinterpret_cast<const uint4*>(P.nl)[i];  // the four line ends of record i
    uint32_t s0 = i ? P.nl[4u * i - 1] + 1u : 0u;
    const uint8_t* __restrict__ t = P.text;
    uint32_t e0 = e.x, e1 = e.y, e2 = e.z, e3 = e.w;
    const uint32_t s1 = e.x + 1, s2 = e.y + 1, s3 = e.z + 1;
    if (e0 > s0 && t[e0 - 1] == '\r') e0--;
    if (e1 > s1 && t[e1 - 1] == '\r') e1--;
    if (e2 > s2 && t[e2 - 1] == '\r') e2--;
    if (e3 > s3 && t[e3 - 1] == '\r') e3--;
    const uint32_t slen = e1 - s1, qlen = e3 - s3;
    bool ok = true;
    if (e0 == s0 || t[s0] != '@') {
        report(P.perr, i, PERR_AT);
        ok = false;
    } else if (e2 == s2 || t[s2] != '+') {
        report(P.perr, i, PERR_PLUS);
        ok = false;
    } else if (slen != qlen) {
        report(P.perr, i, PERR_LEN);
        ok = false;
    } else if (slen > CSQ_MAX_READ_LEN) {
        report(P.perr, i, PERR_LIMIT);
        ok = false;
    }
    // a bad record becomes an empty one: the rest of the chain stays in bounds, the batch is rejected by the host
    P.name_off[i] = ok ? s0 + 1 : s0;
    P.name_end[i] = ok ? e0 : s0;
    P.seq_off[i] = s1;
    P.seq_len[i] = ok ? slen : 0u;
    P.qual_off[i] = s3;
}

}  // namespace

uint32_t csq_parse_tiles(uint64_t bytes) { return (uint32_t)((bytes + TILE_BYTES - 1) / TILE_BYTES); }

cudaError_t csq_launch_parse(const ParseParams& p, uint32_t* tile_cnt, uint16_t* masks, cudaStream_t stream) {
    const uint32_t tiles = csq_parse_tiles(p.bytes);
    if (tiles) {
        k_nl_count<<<tiles, TILE_THREADS, 0, stream>>>(p.text, p.bytes, masks, tile_cnt);
        k_tile_scan<<<1, 1024, 0, stream>>>(tiles, tile_cnt, p.nl_total);
        k_nl_index<<<tiles, TILE_THREADS, 0, stream>>>(masks, tile_cnt, p.nl, 4u * p.n);
    } else {
        cudaError_t e = cudaMemsetAsync(p.nl_total, 0, 4, stream);
        if (e != cudaSuccess) return e;
    }
    k_records<<<(p.n + 255) / 256 + (p.n == 0 ? 1 : 0), 256, 0, stream>>>(p);
    return cudaGetLastError();
}
