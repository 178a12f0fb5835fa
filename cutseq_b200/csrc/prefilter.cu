// k_prefilter<NW>: bit-parallel (Myers 1999 / Hyyro 2001) REJECT-ONLY filter in front of k_align.
//
// One thread per read computes the exact semiglobal edit-distance matrix of Aligner.locate (same
// boundary conditions, same column range) column by column with vertical-delta bit-vectors of
// NW x 32 bits, i.e. the `cost` component of cutadapt's cells and nothing else.  A read survives
// when some cell that Aligner.locate would examine COULD be acceptable:
//   row m, column j (only with QUERY_STOP):  cost <= thr[Lmax], Lmax = min(m, j + cost) >= min_overlap
//   last column, row i >= first_i:           cost <= thr[Lmax], Lmax = min(i, (n - min_n) + cost) >= min_overlap
// Lmax bounds the aligned adapter length from above (an alignment with `cost` errors that ends in
// column j covers at most (j - min_n) + cost adapter characters) and thr[] is monotone, so the test is
// a necessary condition for `length >= min_overlap and cost <= length * max_error_rate`.
// Survivors are compacted into a list and handed to the exact DP (k_align), which alone decides
// whether there is a match and where: the prefilter can only say "no match".
// The kernel also performs the scalar ops in front of the ALIGN op, so the exact pass starts from
// the stored state.
#include <cuda_runtime.h>
#include <stdint.h>

#include "csq_internal.h"
#include "device_common.cuh"

namespace {

template <int NW>
__global__ void __launch_bounds__(256) k_prefilter(const __grid_constant__ AlignParams P, uint32_t* __restrict__ list,
                                                   uint32_t* __restrict__ list_count) {
    __shared__ uint32_t lut[256 * NW];
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        const int u = c & 0xDF;
        const int li = u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : u == 'T' ? 3 : -1;
#pragma unroll
        for (int w = 0; w < NW; w++) lut[c * NW + w] = li >= 0 ? P.peq[li][w] : 0u;
    }
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < P.n;
    bool pass = false;
    unsigned int cells = 0;
    if (valid) {
        ReadState st = P.first ? fresh_state(P.md.seq_len[idx]) : load_state(P.md.state + idx);
        for (int q = 0; q < P.n_pre; q++) apply_scalar(P.pre[q], st);
        store_state(P.md.state + idx, st);

        const int m = P.m, k = P.k;
        const int a = st.a, b = st.b, n = b - a;
        const bool sir = P.flags & 1, siq = P.flags & 2, eir = P.flags & 4, eiq = P.flags & 8;
        int max_n = n, min_n = 0;
        if (!siq) max_n = min(n, m + k);
        if (!eiq) min_n = max(0, n - m - k);
        cells = (unsigned int)(m * (max_n - min_n));

        // column min_n: cost[i] = i (vertical deltas all +1) unless the adapter start is free (all 0)
        uint32_t Pv[NW], Mv[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            Pv[w] = sir ? 0u : 0xFFFFFFFFu;
            Mv[w] = 0u;
        }
        int score = sir ? 0 : m;             // cost[m][min_n]   (min_n == 0 whenever sir, see csq_adapter_kind)
        const int hin0 = siq ? 0 : 1;        // row 0: cost stays 0 (free read prefix) or grows by 1 per column
        const int top = (m - 1) & 31;
        const uint8_t* s = P.md.seq + P.md.seq_off[idx];
        const int step = P.reversed ? -1 : 1;
        const uint8_t* p = P.reversed ? (s + b - 1 - min_n) : (s + a + min_n);
        for (int j = min_n + 1; j <= max_n; j++, p += step) {
            const uint32_t c = *p;
            int hin = hin0;
#pragma unroll
            for (int w = 0; w < NW; w++) {
                uint32_t Eq = lut[c * NW + w];
                const uint32_t hneg = hin < 0 ? 1u : 0u, hpos = hin > 0 ? 1u : 0u;
                const uint32_t Xv = Eq | Mv[w];
                Eq |= hneg;
                const uint32_t Xh = (((Eq & Pv[w]) + Pv[w]) ^ Pv[w]) | Eq;
                uint32_t Ph = Mv[w] | ~(Xh | Pv[w]);
                uint32_t Mh = Pv[w] & Xh;
                const int bit = (w == NW - 1) ? top : 31;
                hin = (int)((Ph >> bit) & 1u) - (int)((Mh >> bit) & 1u);  // horizontal delta leaving this word
                Ph = (Ph << 1) | hpos;
                Mh = (Mh << 1) | hneg;
                Pv[w] = Mh | ~(Xv | Ph);
                Mv[w] = Ph & Xv;
            }
            score += hin;
            if (eiq && score <= k) {
                const int L = min(m, j + score);
                if (L >= P.min_overlap && score <= (int)P.thr[L]) pass = true;
            }
        }
        if (!pass && max_n == n) {
            const int first_i = eir ? 0 : m;
            int d = siq ? 0 : max_n;  // cost[0][max_n]
            const int span = n - min_n;
            for (int i = 1; i <= m; i++) {
                const int w = (i - 1) >> 5, bt = (i - 1) & 31;
                uint32_t pv = Pv[0], mv = Mv[0];
#pragma unroll
                for (int x = 1; x < NW; x++)
                    if (w == x) {
                        pv = Pv[x];
                        mv = Mv[x];
                    }
                d += (int)((pv >> bt) & 1u) - (int)((mv >> bt) & 1u);
                if (i >= first_i && d <= k) {
                    const int L = min(i, span + d);
                    if (L >= P.min_overlap && d <= (int)P.thr[L]) pass = true;
                }
            }
        }
        if (!pass && P.matches) {
            csq_match r;
            r.found = r.ref_start = r.ref_stop = r.query_start = r.query_stop = r.score = r.errors = r.reserved = 0;
            P.matches[idx] = r;
        }
    }
    // nominal DP cells of the launch (GCUPS numerator) and warp-aggregated append of the survivors
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_down_sync(0xffffffffu, cells, o);
    const int lane = threadIdx.x & 31;
    if (lane == 0 && cells) atomicAdd(P.counters + P.counter_index + (CNT_DP_CELLS - CNT_WITH_ADAPTERS), (unsigned long long)cells);
    const unsigned int ballot = __ballot_sync(0xffffffffu, pass);
    if (ballot) {
        const int leader = __ffs(ballot) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(list_count, (unsigned int)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pass) list[base + __popc(ballot & ((1u << lane) - 1u))] = idx;
    }
}

// Homopolymer adapters (the poly-A / poly-T 100-mers of reference run.py:389-404, 674-707) in their two
// non-internal kinds need no bit-vectors.  With e(l) = number of read characters different from the adapter
// base among the l characters next to the anchored end (read start for NonInternalFront, read end for
// NonInternalBack), every alignment that covers those l characters costs at least e(l), and covers at most
// l + cost adapter characters; thr[] grows by at most 1 per step, so "cost <= thr[min(m, l + cost)]" can
// only hold for some cost >= e(l) if it holds for cost == e(l).  e(l) never decreases, hence the scan
// stops at the first l with e(l) > k (about 20 characters into a random read instead of m + k = 115 columns).
__global__ void __launch_bounds__(256) k_prefilter_homo(const __grid_constant__ AlignParams P, uint32_t* __restrict__ list,
                                                        uint32_t* __restrict__ list_count) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < P.n;
    bool pass = false;
    unsigned int cells = 0;
    if (valid) {
        ReadState st = P.first ? fresh_state(P.md.seq_len[idx]) : load_state(P.md.state + idx);
        for (int q = 0; q < P.n_pre; q++) apply_scalar(P.pre[q], st);
        store_state(P.md.state + idx, st);
        const int m = P.m, k = P.k;
        const int a = st.a, b = st.b, n = b - a;
        const int span = min(n, m + k);  // columns the aligner visits (max_n - min_n for both kinds)
        cells = (unsigned int)(m * span);
        const bool from_end = (P.flags & 2) != 0;  // NonInternalBack: anchored at the read end
        const uint8_t* s = P.md.seq + P.md.seq_off[idx];
        const uint8_t* p = from_end ? (s + b - 1) : (s + a);
        const int step = from_end ? -1 : 1;
        int e = 0;
        for (int l = 1; l <= span; l++, p += step) {
            e += ((uint32_t)(*p & 0xDFu) != (uint32_t)P.letter) ? 1 : 0;
            if (e > k) break;
            const int L = min(m, l + e);
            if (L >= P.min_overlap && e <= (int)P.thr[L]) {
                pass = true;
                break;
            }
        }
        if (!pass && P.matches) {
            csq_match r;
            r.found = r.ref_start = r.ref_stop = r.query_start = r.query_stop = r.score = r.errors = r.reserved = 0;
            P.matches[idx] = r;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cells += __shfl_down_sync(0xffffffffu, cells, o);
    const int lane = threadIdx.x & 31;
    if (lane == 0 && cells) atomicAdd(P.counters + P.counter_index + (CNT_DP_CELLS - CNT_WITH_ADAPTERS), (unsigned long long)cells);
    const unsigned int ballot = __ballot_sync(0xffffffffu, pass);
    if (ballot) {
        const int leader = __ffs(ballot) - 1;
        unsigned int base = 0;
        if (lane == leader) base = atomicAdd(list_count, (unsigned int)__popc(ballot));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pass) list[base + __popc(ballot & ((1u << lane) - 1u))] = idx;
    }
}

}  // namespace

cudaError_t csq_launch_prefilter(const AlignParams& p, uint32_t* list, uint32_t* list_count, cudaStream_t stream) {
    if (p.n == 0) return cudaSuccess;
    const dim3 grid((p.n + 255) / 256), block(256);
    if (p.homopolymer && !p.reversed && (p.flags == 9 || p.flags == 6)) {  // NonInternalFront / NonInternalBack
        k_prefilter_homo<<<grid, block, 0, stream>>>(p, list, list_count);
        return cudaGetLastError();
    }
    const int nw = (p.m + 31) / 32;
    switch (nw) {
        case 1: k_prefilter<1><<<grid, block, 0, stream>>>(p, list, list_count); break;
        case 2: k_prefilter<2><<<grid, block, 0, stream>>>(p, list, list_count); break;
        case 3: k_prefilter<3><<<grid, block, 0, stream>>>(p, list, list_count); break;
        default: k_prefilter<4><<<grid, block, 0, stream>>>(p, list, list_count); break;
    }
    return cudaGetLastError();
}
