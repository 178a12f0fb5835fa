// Device gzip reader for BGZF inputs (bgzip / this library's own output): one THREAD per member, all members of a
// batch in flight at once.  Replaces the host inflate of the first version for such files (xopen's gzip backends
// behind cutadapt's InputPaths in the reference, run.py:434, 751): the compressed bytes cross PCIe (~1/4 of the
// text) and no host core decodes anything.  Members are independent DEFLATE streams of at most 64 KiB with their
// compressed size in the 'BC' extra field and ISIZE in the trailer, so the host lays out source and destination
// offsets by walking the member headers, without looking at the payload.
// The decoder (gz_inflate_core.h) keeps its four canonical-Huffman tables (352 16-bit elements per thread) in shared
// memory, interleaved across the threads of the CTA.  '\n' are counted per member on the way: the host cuts batches
// at record boundaries from those counts (pass 1 of the file driver), then the text of a batch is produced where the
// parse kernels expect it (pass 2).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "csq_internal.h"
#include "gz_inflate_core.h"

namespace {

constexpr int INF_THREADS = 64;

__global__ void __launch_bounds__(INF_THREADS) k_gz_inflate(const __grid_constant__ InflateParams P) {
    __shared__ uint16_t tabs[gzi::TAB_ELEMS * INF_THREADS];
    const uint32_t i = blockIdx.x * INF_THREADS + threadIdx.x;
    if (i >= P.n_members) return;
    const uint32_t s0 = P.moff[i], s1 = P.moff[i + 1], o0 = P.ooff[i], o1 = P.ooff[i + 1];
    uint32_t produced = 0, lines = 0;
    const int rc = gzi::inflate_member<INF_THREADS>(P.comp + s0, s1 - s0, P.out + o0, o1 - o0, tabs + threadIdx.x, &produced, &lines);
    if (rc != gzi::OK || produced != o1 - o0) atomicMin(P.status, (rc ? rc : (int)gzi::ERR_TRAILER) + 16 * (int)min(i, 0x7FFFFFu));
    // bit 31: the member's text does not end in a line end (the host needs that for the last member of a file)
    if (P.lines) P.lines[i] = lines | ((produced && P.out[o0 + produced - 1] != '\n') ? 0x80000000u : 0u);
}

}  // namespace

cudaError_t csq_launch_inflate(const InflateParams& p, cudaStream_t stream) {
    if (p.n_members == 0) return cudaSuccess;
    k_gz_inflate<<<(p.n_members + INF_THREADS - 1) / INF_THREADS, INF_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}

// host twin for the CPU tests: the same decoder, member by member (BGZF framing walked as the file driver does)
extern "C" int csq_gz_inflate_host(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, uint64_t* out_n, uint64_t* lines) {
    if (!src || !dst || !out_n) return CSQ_ERR_INVALID;
    std::vector<uint16_t> tab(gzi::TAB_ELEMS);
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
        uint32_t produced = 0, l = 0;
        const int rc = gzi::inflate_member<1>(padded.data(), size, dst + out, isize, tab.data(), &produced, &l);
        if (rc != gzi::OK) return -100 - rc;
        out += produced;
        nl += l;
        pos += size;
    }
    *out_n = out;
    if (lines) *lines = nl;
    return 0;
}
