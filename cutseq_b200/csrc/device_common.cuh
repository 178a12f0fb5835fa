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


}  // namespace
