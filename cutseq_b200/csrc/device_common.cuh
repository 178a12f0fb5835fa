// Device helpers shared by kernels.cu and prefilter.cu: per-read interval state and the O(1) ops.
#pragma once

#include "csq_internal.h"

namespace {

// CUT / COND_CUT / RENAME(capture) on the interval state.
__device__ __forceinline__ void apply_scalar(const DevOp& op, ReadState& st) {
    const int len = (int)st.b - (int)st.a;
    if (op.kind == CSQ_OP_RENAME) {
        st.ren_cp = st.cp;
        st.ren_cs = st.cs;
        return;
    }
    if (op.kind == CSQ_OP_COND_CUT && !(st.matched & 0x80000000u) && len < op.fmin) return;  // run.py:154-155
    if (op.kind == CSQ_OP_CUT || op.kind == CSQ_OP_COND_CUT) {
        if (op.length > 0) {
            const int c = min(op.length, len);
            st.cp = ((uint32_t)st.a << 16) | (uint32_t)c;
            st.a = (uint16_t)(st.a + c);
        } else if (op.length < 0) {
            const int c = min(-op.length, len);
            st.cs = ((uint32_t)(st.b - c) << 16) | (uint32_t)c;
            st.b = (uint16_t)(st.b - c);
        }
    }
}

__device__ __forceinline__ ReadState fresh_state(uint32_t len) {
    ReadState st;
    st.a = 0;
    st.b = (uint16_t)len;
    st.matched = 0;
    st.cp = st.cs = st.ren_cp = st.ren_cs = 0;
    st.id_start = st.id_end = 0;
    st.qtrim = 0;
    return st;
}

__device__ __forceinline__ ReadState load_state(const ReadState* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 x = q[0], y = q[1];
    ReadState st;
    st.a = (uint16_t)(x.x & 0xFFFFu);
    st.b = (uint16_t)(x.x >> 16);
    st.matched = x.y;
    st.cp = x.z;
    st.cs = x.w;
    st.ren_cp = y.x;
    st.ren_cs = y.y;
    st.id_start = (uint16_t)(y.z & 0xFFFFu);
    st.id_end = (uint16_t)(y.z >> 16);
    st.qtrim = y.w;
    return st;
}

__device__ __forceinline__ void store_state(ReadState* p, const ReadState& st) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4((uint32_t)st.a | ((uint32_t)st.b << 16), st.matched, st.cp, st.cs);
    q[1] = make_uint4(st.ren_cp, st.ren_cs, (uint32_t)st.id_start | ((uint32_t)st.id_end << 16), st.qtrim);
}


// 16 bytes from an arbitrary byte address: five aligned 32-bit loads and four funnel shifts.
// Touches up to 3 bytes in front of q and 4 behind q + 16; the device pools are padded for it.
__device__ __forceinline__ uint4 fetch16(const uint8_t* __restrict__ q) {
    const uint32_t* __restrict__ w = reinterpret_cast<const uint32_t*>((uintptr_t)q & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)(uintptr_t)q & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                      __funnelshift_r(w3, w4, sh));
}

struct RecordShape {
    uint32_t id_len, umi_len, seq_len, total;
    uint32_t lenA, lenB;
};

__device__ __forceinline__ RecordShape record_shape(const PairParams& P, const ReadState& own, const ReadState& r1,
                                                    const ReadState& r2) {
    RecordShape rs;
    rs.id_len = (uint32_t)own.id_end - (uint32_t)own.id_start;
    rs.lenA = rs.lenB = 0;
    if (P.rename_parts & CSQ_REN_OWN_PREFIX) rs.lenA = own.ren_cp & 0xFFFFu;
    if (P.rename_parts & CSQ_REN_OWN_SUFFIX) rs.lenB = own.ren_cs & 0xFFFFu;
    if (P.rename_parts & CSQ_REN_R1_PREFIX) rs.lenA = r1.ren_cp & 0xFFFFu;
    if (P.rename_parts & CSQ_REN_R2_PREFIX) rs.lenB = r2.ren_cp & 0xFFFFu;
    rs.umi_len = (P.rename_parts ? 1u : 0u) + rs.lenA + rs.lenB;
    rs.seq_len = (uint32_t)own.b - (uint32_t)own.a;
    rs.total = 1 + rs.id_len + rs.umi_len + 1 + rs.seq_len + 3 + rs.seq_len + 1;
    return rs;
}

__device__ __forceinline__ uint8_t complement_base(uint8_t c) {
    // dnaio reverse_complement table: ACGTUMRWSYKVHDBN -> TGCAAKYWSRMBDHVN (case kept), others unchanged
    const char* from = "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn";
    const char* to = "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn";
#pragma unroll
    for (int i = 0; i < 32; i++)
        if (c == (uint8_t)from[i]) return (uint8_t)to[i];
    return c;
}

}  // namespace
