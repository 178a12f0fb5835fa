// Order-preserving FASTQ text emission ("@name\nseq\n+\nqual\n", dnaio's fastq_bytes), the sink behind
// PairedEndSink / SingleEndSink and the two filters of reference run.py:446-471 / 763-792.
// k_emit_rec (CSQ_PLAN_EMIT_REC) is an alternative to the direct k_emit<G> of kernels.cu, kept for A/B runs (the
// default emitter is k_emit_stage, emit_stage.cu): it needs 4x fewer instructions than k_emit<32>, but its
// per-thread access pattern makes it latency / L1-sector bound and it measures 7 % slower (1.26 ms against
// 1.17 ms per 2 M pairs, profiles/r01_emit_variants.md).
//
// k_emit_rec: one THREAD per pair (both mates).  Phase 1 is a CTA-wide exclusive scan of the record sizes per output
// stream (destination x mate), which places every record; phase 2 streams each record into its place
// through a 16-byte staging chunk held in registers:
//
//   * the output of a record is cut at the 16-byte boundaries of the output buffer; every full chunk is
//     written with ONE 128-bit store; only the first and the last chunk of a record, which it shares with
//     its neighbours, are written narrower (words / bytes), so no two threads ever write the same byte;
//   * a data piece (id, UMI parts, bases, qualities) is fetched 16 bytes at a time from an arbitrary byte
//     address: five aligned 32-bit loads and four funnel shifts; chunks that lie inside one piece - about
//     85 % of them - cost 4 loads + 4 shifts + 1 store;
//   * at a piece boundary the 16 source bytes are merged into the staging chunk under a byte mask
//     (prefix-mask table in shared memory), separators are OR-ed in as single bytes.
//
// A warp instruction therefore moves 32 x 16 bytes, against 32 x 4 in the warp-per-record kernel this one
// replaces (k_emit in kernels.cu, kept behind CSQ_PLAN_EMIT_WARP for A/B runs): that one spent ~250 warp
// instructions per record, mostly warp-uniform bookkeeping replicated in 32 lanes, and was bound by
// instruction issue (78 % issue-slot utilisation at 27 % of the HBM peak).  Here the bookkeeping is per
// thread, i.e. paid once per record.  Bound: HBM (every input byte read once, every output byte written once).
//
// Reads of up to 15 bytes in front of a piece and 19 behind it are harmless: device pools carry 64 bytes of
// padding on both sides (plan.cu).
#include <cuda_runtime.h>
#include <stdint.h>

#include "csq_internal.h"
#include "device_common.cuh"

namespace {

struct Stage {
    uint4 acc;      // the chunk being assembled; bytes [lo, fill) are set, all others zero
    uint8_t* cptr;  // its 16-byte aligned address
    uint32_t fill;  // next byte position in the chunk, 0..15
    uint32_t lo;    // first byte of the chunk that belongs to this record (non-zero only in its first chunk)
};

__device__ __forceinline__ void stage_begin(Stage& st, uint8_t* out) {
    st.acc = make_uint4(0, 0, 0, 0);
    st.cptr = reinterpret_cast<uint8_t*>((uintptr_t)out & ~(uintptr_t)15);
    st.fill = st.lo = (uint32_t)((uintptr_t)out & 15u);
}

// bytes [lo, hi) of the chunk, as whole words where the record owns them
__device__ __forceinline__ void store_partial(const Stage& st, uint32_t hi) {
    const uint32_t w[4] = {st.acc.x, st.acc.y, st.acc.z, st.acc.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t b0 = 4u * i;
        if (st.lo <= b0 && hi >= b0 + 4u) {
            reinterpret_cast<uint32_t*>(st.cptr)[i] = w[i];
        } else if (st.lo < b0 + 4u && hi > b0) {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (b0 + j >= st.lo && b0 + j < hi) st.cptr[b0 + j] = (uint8_t)(w[i] >> (8 * j));
        }
    }
}

__device__ __forceinline__ void stage_advance(Stage& st) {  // the chunk is complete up to byte 16
    if (st.lo == 0)
        *reinterpret_cast<uint4*>(st.cptr) = st.acc;
    else
        store_partial(st, 16);
    st.cptr += 16;
    st.fill = st.lo = 0;
    st.acc = make_uint4(0, 0, 0, 0);
}

__device__ __forceinline__ void append_byte(Stage& st, uint32_t c) {
    const uint32_t v = c << ((st.fill & 3u) * 8u), w = st.fill >> 2;
    st.acc.x |= w == 0 ? v : 0u;
    st.acc.y |= w == 1 ? v : 0u;
    st.acc.z |= w == 2 ? v : 0u;
    st.acc.w |= w == 3 ? v : 0u;
    if (++st.fill == 16) stage_advance(st);
}

// 16 bytes at byte offset k (0..15) of the 32 bytes (a, b): word select by k >> 2 (a two-level SEL network, no
// dynamic register indexing), then the funnel shift by k & 3.
__device__ __forceinline__ uint4 shift32(const uint4 a, const uint4 b, uint32_t k) {
    const bool q2 = (k & 8u) != 0, q1 = (k & 4u) != 0;
    const uint32_t sh = (k & 3u) * 8u;
    const uint32_t t0 = q2 ? a.z : a.x, t1 = q2 ? a.w : a.y, t2 = q2 ? b.x : a.z, t3 = q2 ? b.y : a.w, t4 = q2 ? b.z : b.x,
                   t5 = q2 ? b.w : b.y;
    const uint32_t x0 = q1 ? t1 : t0, x1 = q1 ? t2 : t1, x2 = q1 ? t3 : t2, x3 = q1 ? t4 : t3, x4 = q1 ? t5 : t4;
    return make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh),
                      __funnelshift_r(x3, x4, sh));
}

// pm[k]: the first k bytes of a chunk
__device__ __forceinline__ void append_data(Stage& st, const uint8_t* __restrict__ src, uint32_t len, const uint4* __restrict__ pm) {
    while (len) {
        if (st.fill == 0 && len >= 16) {
            // inside one piece, chunk by chunk: ONE aligned 128-bit load per chunk (read-only path, so the loads
            // of the next chunks are not ordered behind the stores), the previous load supplies the low bytes
            const uint32_t k = (uint32_t)(uintptr_t)src & 15u;
            const uint4* __restrict__ p = reinterpret_cast<const uint4*>((uintptr_t)src & ~(uintptr_t)15);
            uint4 prev = __ldg(p);
            while (len >= 32) {
                const uint4 n1 = __ldg(p + 1), n2 = __ldg(p + 2);
                reinterpret_cast<uint4*>(st.cptr)[0] = shift32(prev, n1, k);
                reinterpret_cast<uint4*>(st.cptr)[1] = shift32(n1, n2, k);
                prev = n2;
                p += 2;
                st.cptr += 32;
                src += 32;
                len -= 32;
            }
            if (len >= 16) {
                const uint4 n1 = __ldg(p + 1);
                *reinterpret_cast<uint4*>(st.cptr) = shift32(prev, n1, k);
                st.cptr += 16;
                src += 16;
                len -= 16;
            }
            continue;
        }
        const uint32_t take = min(16u - st.fill, len);
        const uint4 v = fetch16(src - st.fill);  // byte i of v is chunk byte i
        const uint4 mh = pm[st.fill + take], ml = pm[st.fill];
        st.acc.x |= v.x & mh.x & ~ml.x;
        st.acc.y |= v.y & mh.y & ~ml.y;
        st.acc.z |= v.z & mh.z & ~ml.z;
        st.acc.w |= v.w & mh.w & ~ml.w;
        st.fill += take;
        src += take;
        len -= take;
        if (st.fill == 16) stage_advance(st);
    }
}

__device__ __forceinline__ void stage_end(Stage& st) {
    if (st.fill > st.lo) store_partial(st, st.fill);
}

__global__ void __launch_bounds__(CSQ_PAIR_BLOCK) k_emit_rec(const __grid_constant__ EmitParams E) {
    const PairParams& P = E.pp;
    __shared__ uint4 pm[17];
    __shared__ unsigned int wtot[8][CSQ_PAIR_BLOCK / 32];  // per-stream totals of every warp
    const uint32_t base = blockIdx.x * CSQ_PAIR_BLOCK;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool paired = P.n_mates == 2;
    const int n_mates = paired ? 2 : 1;
    if (threadIdx.x < 17) {
        const uint32_t k = threadIdx.x;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t nb = k > 4u * i ? min(k - 4u * i, 4u) : 0u;  // bytes of word i below k
            w[i] = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
        }
        pm[k] = make_uint4(w[0], w[1], w[2], w[3]);
    }

    // ---- phase 1: where does every record go ----
    const uint32_t idx = base + threadIdx.x;
    const bool live = idx < P.n;
    int dest = -1;
    uint32_t len[2] = {0, 0}, incl[2] = {0, 0};
    ReadState st[2];
    RecordShape shape[2];
    if (live) {
        dest = P.dest[idx];
        st[0] = load_state(P.md[0].state + idx);
        st[1] = paired ? load_state(P.md[1].state + idx) : st[0];
        for (int mt = 0; mt < n_mates; mt++) {
            shape[mt] = record_shape(P, st[mt], st[0], st[1]);
            len[mt] = shape[mt].total;
        }
    }
    for (int mt = 0; mt < n_mates; mt++)
        for (int d = 0; d < CSQ_N_DEST; d++) {
            const uint32_t v = dest == d ? len[mt] : 0u;
            uint32_t x = v;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (dest == d) incl[mt] = x;
            if (lane == 31) wtot[d * 2 + mt][wid] = x;
        }
    __syncthreads();
    if (!live) return;

    // ---- phase 2: stream the records out ----
    const uint8_t* const poolA = (P.rename_parts & CSQ_REN_R1_PREFIX) ? P.md[0].seq : nullptr;
    const uint8_t* const poolB = (P.rename_parts & CSQ_REN_R2_PREFIX) ? P.md[1].seq : nullptr;
    const uint8_t* id_ptr[2] = {nullptr, nullptr};
    for (int mt = 0; mt < n_mates; mt++) {
        const int stream = dest * 2 + mt;
        uint32_t off = incl[mt] - len[mt];
        for (int w = 0; w < wid; w++) off += wtot[stream][w];
        const MateDev& md = P.md[mt];
        const ReadState& own = st[mt];
        const RecordShape& sh = shape[mt];
        uint8_t* out = E.out[dest][mt] + E.block_off[(size_t)blockIdx.x * 8 + stream] + off;
        const uint32_t so = md.seq_off[idx], qo = md.qual_off[idx];
        const uint8_t* nm = md.name + md.name_off[idx] + own.id_start;
        id_ptr[mt] = nm;
        // UMI parts: slices of the mate they come from (own read for the single-end template)
        const uint8_t *pa = nullptr, *pb = nullptr;
        if (P.rename_parts & CSQ_REN_OWN_PREFIX) pa = md.seq + so + (own.ren_cp >> 16);
        if (P.rename_parts & CSQ_REN_OWN_SUFFIX) pb = md.seq + so + (own.ren_cs >> 16);
        if (P.rename_parts & CSQ_REN_R1_PREFIX) pa = poolA + P.md[0].seq_off[idx] + (st[0].ren_cp >> 16);
        if (P.rename_parts & CSQ_REN_R2_PREFIX) pb = poolB + P.md[1].seq_off[idx] + (st[1].ren_cp >> 16);
        if (P.revcomp) {  // single-end --auto-rc on the '-' strand (run.py:420-426): bytewise, rare
            uint8_t* o = out;
            *o++ = '@';
            for (uint32_t x = 0; x < sh.id_len; x++) *o++ = nm[x];
            if (sh.umi_len) {
                *o++ = '_';
                for (uint32_t x = 0; x < sh.lenA; x++) *o++ = pa[x];
                for (uint32_t x = 0; x < sh.lenB; x++) *o++ = pb[x];
            }
            *o++ = '\n';
            for (uint32_t x = 0; x < sh.seq_len; x++) *o++ = complement_base(md.seq[so + own.b - 1 - x]);
            *o++ = '\n';
            *o++ = '+';
            *o++ = '\n';
            for (uint32_t x = 0; x < sh.seq_len; x++) *o++ = md.qual[qo + own.b - 1 - x];
            *o++ = '\n';
            continue;
        }
        Stage sg;
        stage_begin(sg, out);
        append_byte(sg, '@');
        append_data(sg, nm, sh.id_len, pm);
        if (sh.umi_len) {
            append_byte(sg, '_');
            append_data(sg, pa, sh.lenA, pm);
            append_data(sg, pb, sh.lenB, pm);
        }
        append_byte(sg, '\n');
        append_data(sg, md.seq + so + own.a, sh.seq_len, pm);
        append_byte(sg, '\n');
        append_byte(sg, '+');
        append_byte(sg, '\n');
        append_data(sg, md.qual + qo + own.a, sh.seq_len, pm);
        append_byte(sg, '\n');
        stage_end(sg);
    }
    if (P.check_ids) {  // PairedEndRenamer: the ids of the two mates must be identical
        bool bad = shape[0].id_len != shape[1].id_len;
        for (uint32_t x = 0; !bad && x < shape[0].id_len; x += 16) {
            const uint4 u = fetch16(id_ptr[0] + x), v = fetch16(id_ptr[1] + x), m = pm[min(16u, shape[0].id_len - x)];
            bad = (((u.x ^ v.x) & m.x) | ((u.y ^ v.y) & m.y) | ((u.z ^ v.z) & m.z) | ((u.w ^ v.w) & m.w)) != 0;
        }
        if (bad) atomicExch(P.error_flag, (int)CSQ_ERR_PAIRING);
    }
}

}  // namespace

cudaError_t csq_launch_emit_rec(const EmitParams& p, cudaStream_t stream) {
    if (p.pp.n == 0) return cudaSuccess;
    k_emit_rec<<<(p.pp.n + CSQ_PAIR_BLOCK - 1) / CSQ_PAIR_BLOCK, CSQ_PAIR_BLOCK, 0, stream>>>(p);
    return cudaGetLastError();
}
