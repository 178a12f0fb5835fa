// Device gzip writer: the emitted FASTQ text of a batch -> concatenated BGZF-framed gzip members, on the GPU.
// Replaces the host zlib deflate of the first version behind the output files (xopen's gzip backends behind
// cutadapt's OutputFiles in the reference, run.py:449-470, 767-790; its default output names end in .fastq.gz,
// run.py:1058-1093).  The host then only moves compressed bytes: ~3.5x fewer over PCIe and no deflate on its cores.
//
//   k_gz_hist      byte histogram of every output stream of the batch (every 8th 16-byte vector: frequencies only
//                  steer code lengths; all 256 literals get a code, so unsampled bytes stay encodable)
//   k_gz_build     one thread per stream: length-limited Huffman code (gz_core.h), its canonical codes and the bits of
//                  the dynamic-block header - ONE code per stream and batch, shared by all of its members
//   k_gz_deflate   one CTA per member (GZ_PIECE input bytes): stage the piece in shared memory, every thread encodes a
//                  run of GZ_RUN literals into the shared bit image (bit offsets by a CTA-wide scan of the code
//                  lengths), CRC-32 per run folded with the "bytes that follow" operator x^(8n) mod P
//                  (crc32_combine), BGZF header / trailer, member -> its slot; a piece that does not shrink is stored
//   k_gz_scan      exclusive scan of the member sizes per stream -> member offsets and stream totals
//   k_gz_pack      members -> their place in the packed stream (what crosses PCIe and lands in the file)
// Literal-only DEFLATE (dynamic Huffman, no LZ77 matches): on FASTQ text - four bases at ~2 bits, few distinct
// quality values, digits in the ids - that is most of what level-1 zlib achieves, at HBM speed.
// Every member is a complete gzip member of at most 64 KiB with the 'BC' extra field, i.e. the file is also BGZF.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "csq_internal.h"
#include "gz_core.h"

namespace {

constexpr int DF_THREADS = GZ_THREADS;
constexpr int DF_RUN = GZ_RUN;
constexpr int DF_PIECE = GZ_PIECE;
constexpr int DF_SLOT = GZ_SLOT;
constexpr int DF_OUT_WORDS = DF_SLOT / 4;
constexpr int N_STREAMS = CSQ_N_DEST * 2;

__device__ __forceinline__ int stream_of_member(const GzParams& P, uint32_t member) {
    int s = 0;
#pragma unroll
    for (int i = 1; i < N_STREAMS; i++)
        if (member >= P.first_member[i]) s = i;
    return s;
}

// ---- histogram -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gz_hist(const __grid_constant__ GzParams P) {
    __shared__ uint32_t h[8][256];  // one histogram per warp: the few FASTQ symbols would serialise a shared one
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t member = blockIdx.x;
    const int s = stream_of_member(P, member);
    const uint64_t off = (uint64_t)(member - P.first_member[s]) * DF_PIECE;
    const uint64_t left = P.bytes[s] - off;
    const uint32_t len = left < (uint64_t)DF_PIECE ? (uint32_t)left : (uint32_t)DF_PIECE;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(P.text[s] + off);
    uint32_t* mine = h[threadIdx.x >> 5];
    for (uint32_t v = threadIdx.x * 8u; v * 16u < len; v += 256u * 8u) {  // every 8th vector
        const uint4 x = src[v];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
        const uint32_t nb = len - v * 16u < 16u ? len - v * 16u : 16u;
#pragma unroll
        for (int i = 0; i < 16; i++)
            if ((uint32_t)i < nb) atomicAdd(&mine[(w[i >> 2] >> (8 * (i & 3))) & 0xFFu], 1u);
    }
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) t += h[w][threadIdx.x];
    if (t) atomicAdd(P.hist + s * 256 + threadIdx.x, t);
}

// ---- one Huffman code per stream ---------------------------------------------------------------------------------------
__global__ void k_gz_build(const __grid_constant__ GzParams P) {
    const int s = threadIdx.x;
    if (s >= N_STREAMS) return;
    uint32_t f[gz::N_LITLEN];
    // every literal keeps a code (the histogram is a sample); the sampled counts are scaled up so that the +1 of
    // absent symbols does not cost the frequent ones a bit
    unsigned long long sum = 0;
    for (int i = 0; i < 256; i++) sum += P.hist[s * 256 + i];
    int shift = 0;
    while ((sum >> shift) > (1ull << 26)) shift++;  // the frequencies of a code must add up below 2^32
    for (int i = 0; i < 256; i++) f[i] = ((P.hist[s * 256 + i] >> shift) << 4) + 1u;
    f[256] = 1u;  // end of block, once per member
    uint8_t len[gz::N_LITLEN];
    uint16_t order[gz::N_LITLEN];
    uint32_t work[gz::N_LITLEN];
    gz::huff_lengths(f, gz::N_LITLEN, gz::MAX_BITS, len, order, work);
    gz::huff_codes(len, gz::N_LITLEN, P.codes + s * GZ_CODE_STRIDE);
    uint32_t* hdr = P.hdr + s * GZ_HDR_STRIDE;
    for (int i = 0; i < gz::HDR_WORDS; i++) hdr[i] = 0;
    hdr[gz::HDR_WORDS] = gz::dyn_header(len, hdr);
}

// x^(8 n) mod P on the device (short last pieces only; full pieces use the table)
__device__ uint32_t crc_xpow8_dev(uint32_t n) {
    uint32_t p = 1u << 31, sq = 0x00800000u;
    while (n) {
        if (n & 1u) p = gz::crc_mulmod(sq, p);
        sq = gz::crc_mulmod(sq, sq);
        n >>= 1;
    }
    return p;
}

__device__ __forceinline__ void or_bits(uint32_t* w, uint32_t pos, uint32_t v, uint32_t n) {  // n <= 32, shared memory
    if (n == 0) return;
    const uint32_t word = pos >> 5, sh = pos & 31u;
    atomicOr(&w[word], v << sh);
    if (sh + n > 32u) atomicOr(&w[word + 1], v >> (32u - sh));
}

// ---- one member per CTA ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DF_THREADS) k_gz_deflate(const __grid_constant__ GzParams P) {
    extern __shared__ __align__(16) uint8_t df_smem[];
    uint8_t* const in = df_smem;                                              // DF_PIECE (+16)
    uint32_t* const outw = reinterpret_cast<uint32_t*>(df_smem + DF_PIECE + 16);  // DF_OUT_WORDS
    uint32_t* const codes = outw + DF_OUT_WORDS;                              // 257 (+3)
    uint32_t* const crc_tab = codes + 260;                                    // 256
    __shared__ uint32_t wsum[DF_THREADS / 32], wcrc[DF_THREADS / 32];
    __shared__ uint32_t s_total_bits, s_crc;
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const uint32_t member = blockIdx.x;
    const int s = stream_of_member(P, member);
    const uint64_t off = (uint64_t)(member - P.first_member[s]) * DF_PIECE;
    const uint64_t left = P.bytes[s] - off;
    const uint32_t len = left < (uint64_t)DF_PIECE ? (uint32_t)left : (uint32_t)DF_PIECE;
    // stage the piece (16-byte aligned in the stream: DF_PIECE is a multiple of 16), tables, zeroed bit image
    {
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(P.text[s] + off);
        uint4* dst = reinterpret_cast<uint4*>(in);
        for (uint32_t v = t; v * 16u < len; v += DF_THREADS) dst[v] = src[v];
        for (int i = t; i < DF_OUT_WORDS; i += DF_THREADS) outw[i] = 0u;
        for (int i = t; i < gz::N_LITLEN; i += DF_THREADS) codes[i] = P.codes[s * GZ_CODE_STRIDE + i];
        crc_tab[t] = P.crc_tab[t];
    }
    __syncthreads();
    // run of this thread, its CRC and its bit count
    const uint32_t lo = min((uint32_t)t * DF_RUN, len), hi = min(lo + DF_RUN, len);
    uint32_t crc = 0xFFFFFFFFu, bits = 0;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t b = in[i];
        crc = crc_tab[(crc ^ b) & 0xFFu] ^ (crc >> 8);
        bits += codes[b] >> 16;
    }
    crc = hi > lo ? ~crc : 0u;
    // fold: crc(A || B) = crc(A) * x^(8 |B|) + crc(B); full pieces take the operator from the table
    if (hi > lo) {
        const uint32_t after = len - hi;
        const uint32_t op = len == (uint32_t)DF_PIECE ? P.crc_pow[DF_THREADS - 1 - t] : crc_xpow8_dev(after);
        crc = gz::crc_mulmod(op, crc);
    }
    uint32_t x = bits, c = crc;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 31) wsum[wid] = x;
    if (lane == 0) wcrc[wid] = c;
    __syncthreads();
    uint32_t before = x - bits;
    for (int w = 0; w < wid; w++) before += wsum[w];
    if (t == DF_THREADS - 1) {
        s_total_bits = before + bits;
        uint32_t all = 0;
        for (int w = 0; w < DF_THREADS / 32; w++) all ^= wcrc[w];
        s_crc = all;
    }
    __syncthreads();
    const uint32_t hdr_bits = P.hdr[s * GZ_HDR_STRIDE + gz::HDR_WORDS];
    const uint32_t eob = codes[256];
    const uint32_t total_bits = hdr_bits + s_total_bits + (eob >> 16);
    const uint32_t huff_bytes = (total_bits + 7u) >> 3;
    const bool stored = huff_bytes >= len + 5u;
    uint8_t* const outb = reinterpret_cast<uint8_t*>(outw);
    uint32_t data_bytes;
    if (!stored) {
        data_bytes = huff_bytes;
        const uint32_t base = gz::GZ_HEAD * 8u;
        if ((uint32_t)t * 32u < hdr_bits) {
            const uint32_t nb = hdr_bits - (uint32_t)t * 32u < 32u ? hdr_bits - (uint32_t)t * 32u : 32u;
            uint32_t v = P.hdr[s * GZ_HDR_STRIDE + t];
            if (nb < 32u) v &= (1u << nb) - 1u;
            or_bits(outw, base + (uint32_t)t * 32u, v, nb);
        }
        uint32_t pos = base + hdr_bits + before;
        unsigned long long acc = 0;
        uint32_t nacc = pos & 31u, word = pos >> 5;
        for (uint32_t i = lo; i < hi; i++) {
            const uint32_t cd = codes[in[i]];
            acc |= (unsigned long long)(cd & 0xFFFFu) << nacc;
            nacc += cd >> 16;
            if (nacc >= 32u) {
                atomicOr(&outw[word++], (uint32_t)acc);
                acc >>= 32;
                nacc -= 32u;
            }
        }
        if (t == DF_THREADS - 1) {  // end of block behind the last literal
            acc |= (unsigned long long)(eob & 0xFFFFu) << nacc;
            nacc += eob >> 16;
            if (nacc >= 32u) {
                atomicOr(&outw[word++], (uint32_t)acc);
                acc >>= 32;
                nacc -= 32u;
            }
        }
        if (nacc) atomicOr(&outw[word], (uint32_t)acc);
    } else {
        // stored block: BFINAL = 1, BTYPE = 00, pad to the byte, LEN, NLEN, the bytes
        data_bytes = len + 5u;
        if (t == 0) {
            outb[gz::GZ_HEAD + 0] = 1;
            outb[gz::GZ_HEAD + 1] = (uint8_t)(len & 0xFFu);
            outb[gz::GZ_HEAD + 2] = (uint8_t)(len >> 8);
            outb[gz::GZ_HEAD + 3] = (uint8_t)(~len & 0xFFu);
            outb[gz::GZ_HEAD + 4] = (uint8_t)((~len >> 8) & 0xFFu);
        }
        for (uint32_t i = t; i < len; i += DF_THREADS) outb[gz::GZ_HEAD + 5 + i] = in[i];
    }
    __syncthreads();
    const uint32_t member_size = gz::GZ_HEAD + data_bytes + gz::GZ_TAIL;
    if (t == 0) {
        gz::bgzf_header(outb, member_size);
        uint8_t* tail = outb + gz::GZ_HEAD + data_bytes;
        const uint32_t crc_all = s_crc;
        for (int i = 0; i < 4; i++) {
            tail[i] = (uint8_t)(crc_all >> (8 * i));
            tail[4 + i] = (uint8_t)(len >> (8 * i));
        }
        P.msize[member] = member_size;
    }
    __syncthreads();
    uint4* __restrict__ slot = reinterpret_cast<uint4*>(P.slots + (size_t)member * DF_SLOT);
    const uint4* img = reinterpret_cast<const uint4*>(outw);
    for (uint32_t v = t; v * 16u < member_size; v += DF_THREADS) slot[v] = img[v];
}

// ---- member offsets ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_gz_scan(const __grid_constant__ GzParams P) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry;
    const int s = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t first = P.first_member[s], n = P.first_member[s + 1] - first;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? P.msize[first + i] : 0ull;
        unsigned long long x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long ws = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const unsigned long long before = carry + (wid ? warp_sums[wid - 1] : 0ull) + (x - v);
        if (i < n) P.moff[first + i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) P.totals[s] = carry;
}

// ---- members to their place in the packed stream ----------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_gz_pack(const __grid_constant__ GzParams P) {
    const uint32_t member = blockIdx.x;
    const int s = stream_of_member(P, member);
    const uint32_t size = P.msize[member];
    const uint8_t* __restrict__ src = P.slots + (size_t)member * DF_SLOT;
    uint8_t* __restrict__ dst = P.packed[s] + P.moff[member];
    // head up to the next 4-byte boundary of the destination, whole words built from two aligned source words, tail
    const uint32_t head = min((4u - ((uint32_t)(uintptr_t)dst & 3u)) & 3u, size);
    if (threadIdx.x < head) dst[threadIdx.x] = src[threadIdx.x];
    const uint32_t nw = (size - head) >> 2;
    const uint32_t* __restrict__ sw = reinterpret_cast<const uint32_t*>(src);  // slots are 16-byte aligned
    uint32_t* __restrict__ dw = reinterpret_cast<uint32_t*>(dst + head);
    const uint32_t sh = head * 8u;  // source byte offset of the first whole word == head (0..3)
    for (uint32_t k = threadIdx.x; k < nw; k += 128) dw[k] = __funnelshift_r(sw[k], sw[k + 1], sh);
    const uint32_t done = head + 4u * nw;
    if (threadIdx.x < size - done) dst[done + threadIdx.x] = src[done + threadIdx.x];
}

}  // namespace

size_t csq_gz_deflate_smem() { return (size_t)DF_PIECE + 16 + (size_t)DF_SLOT + (260 + 256) * 4; }

// Host part of the device tables: CRC-32 byte table and the operators x^(8 * GZ_RUN * k) mod P, k = 0..GZ_THREADS-1
void csq_gz_host_tables(uint32_t* crc_tab /*[256]*/, uint32_t* crc_pow /*[GZ_THREADS]*/) {
    gz::crc_make_table(crc_tab);
    for (int k = 0; k < GZ_THREADS; k++) crc_pow[k] = gz::crc_xpow8((uint64_t)k * GZ_RUN);
}

cudaError_t csq_launch_gz(const GzParams& p, uint32_t n_members, cudaStream_t stream) {
    if (n_members == 0) return cudaSuccess;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_gz_deflate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csq_gz_deflate_smem());
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    cudaError_t e = cudaMemsetAsync(p.hist, 0, (size_t)N_STREAMS * 256 * 4, stream);
    if (e != cudaSuccess) return e;
    k_gz_hist<<<n_members, 256, 0, stream>>>(p);
    k_gz_build<<<1, 32, 0, stream>>>(p);
    k_gz_deflate<<<n_members, DF_THREADS, csq_gz_deflate_smem(), stream>>>(p);
    k_gz_scan<<<N_STREAMS, 1024, 0, stream>>>(p);
    k_gz_pack<<<n_members, 128, 0, stream>>>(p);
    return cudaGetLastError();
}

// ---- host twin for the CPU tests: the same pieces, code construction, header bits, bit packing and CRC folding, run
// sequentially (what thread t of the kernel does is done for t = 0 .. GZ_THREADS-1 in turn) ---------------------------------
extern "C" int csq_gz_deflate_host(const uint8_t* text, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_n) {
    if ((!text && n) || !out || !out_n) return CSQ_ERR_INVALID;
    uint32_t crc_tab[256], crc_pow[GZ_THREADS];
    csq_gz_host_tables(crc_tab, crc_pow);
    uint32_t hist[256] = {0};
    for (uint64_t off = 0; off < n; off += DF_PIECE) {  // the kernel's sample: every 8th 16-byte vector of every piece
        const uint64_t len = n - off < (uint64_t)DF_PIECE ? n - off : (uint64_t)DF_PIECE;
        for (uint64_t v = 0; v * 16 < len; v += 8)
            for (uint64_t i = v * 16; i < v * 16 + 16 && i < len; i++) hist[text[off + i]]++;
    }
    uint32_t f[gz::N_LITLEN];
    unsigned long long sum = 0;
    for (int i = 0; i < 256; i++) sum += hist[i];
    int shift = 0;
    while ((sum >> shift) > (1ull << 26)) shift++;
    for (int i = 0; i < 256; i++) f[i] = ((hist[i] >> shift) << 4) + 1u;
    f[256] = 1u;
    uint8_t len_[gz::N_LITLEN];
    uint16_t order[gz::N_LITLEN];
    uint32_t work[gz::N_LITLEN], codes[gz::N_LITLEN], hdr[gz::HDR_WORDS + 1] = {0};
    gz::huff_lengths(f, gz::N_LITLEN, gz::MAX_BITS, len_, order, work);
    gz::huff_codes(len_, gz::N_LITLEN, codes);
    const uint32_t hdr_bits = gz::dyn_header(len_, hdr);
    uint64_t pos = 0;
    std::vector<uint32_t> img(DF_OUT_WORDS + 2);
    for (uint64_t off = 0; off < n; off += DF_PIECE) {
        const uint32_t len = (uint32_t)(n - off < (uint64_t)DF_PIECE ? n - off : (uint64_t)DF_PIECE);
        const uint8_t* in = text + off;
        std::fill(img.begin(), img.end(), 0u);
        uint32_t crc_all = 0, total = 0;
        for (int t = 0; t < DF_THREADS; t++) {
            const uint32_t lo = std::min<uint32_t>((uint32_t)t * DF_RUN, len), hi = std::min<uint32_t>(lo + DF_RUN, len);
            uint32_t crc = 0xFFFFFFFFu;
            for (uint32_t i = lo; i < hi; i++) {
                crc = crc_tab[(crc ^ in[i]) & 0xFFu] ^ (crc >> 8);
                total += codes[in[i]] >> 16;
            }
            if (hi > lo) {
                crc = ~crc;
                const uint32_t op = len == (uint32_t)DF_PIECE ? crc_pow[DF_THREADS - 1 - t] : gz::crc_xpow8(len - hi);
                crc_all ^= gz::crc_mulmod(op, crc);
            }
        }
        const uint32_t total_bits = hdr_bits + total + (codes[256] >> 16);
        const uint32_t huff_bytes = (total_bits + 7u) >> 3;
        uint8_t* outb = reinterpret_cast<uint8_t*>(img.data());
        uint32_t data_bytes;
        if (huff_bytes < len + 5u) {
            data_bytes = huff_bytes;
            gz::BitSink sink = {img.data(), gz::GZ_HEAD * 8u};
            for (uint32_t w = 0; w * 32u < hdr_bits; w++) gz::put_bits(sink, hdr_bits - w * 32u < 32u ? hdr[w] & ((1u << (hdr_bits - w * 32u)) - 1u) : hdr[w],
                                                                     (int)(hdr_bits - w * 32u < 32u ? hdr_bits - w * 32u : 32u));
            for (uint32_t i = 0; i < len; i++) gz::put_bits(sink, codes[in[i]] & 0xFFFFu, (int)(codes[in[i]] >> 16));
            gz::put_bits(sink, codes[256] & 0xFFFFu, (int)(codes[256] >> 16));
        } else {
            data_bytes = len + 5u;
            outb[gz::GZ_HEAD] = 1;
            outb[gz::GZ_HEAD + 1] = (uint8_t)(len & 0xFFu);
            outb[gz::GZ_HEAD + 2] = (uint8_t)(len >> 8);
            outb[gz::GZ_HEAD + 3] = (uint8_t)(~len & 0xFFu);
            outb[gz::GZ_HEAD + 4] = (uint8_t)((~len >> 8) & 0xFFu);
            memcpy(outb + gz::GZ_HEAD + 5, in, len);
        }
        const uint32_t member_size = gz::GZ_HEAD + data_bytes + gz::GZ_TAIL;
        gz::bgzf_header(outb, member_size);
        for (int i = 0; i < 4; i++) {
            outb[gz::GZ_HEAD + data_bytes + i] = (uint8_t)(crc_all >> (8 * i));
            outb[gz::GZ_HEAD + data_bytes + 4 + i] = (uint8_t)(len >> (8 * i));
        }
        if (pos + member_size > cap) return CSQ_ERR_CAPACITY;
        memcpy(out + pos, outb, member_size);
        pos += member_size;
    }
    *out_n = pos;
    return 0;
}
