// Host FASTQ side: reader (plain / gzip) -> packed SoA batches in pinned memory, and the
// writer (plain / gzip members) used by the whole-file driver in pipeline.cpp.
// Stands where dnaio (_core.pyx FASTQ parser, SequenceRecord.fastq_bytes) and xopen sit behind
// cutadapt's InputPaths / OutputFiles in the reference (run.py:434-441, 751-758).
#include <cuda_runtime.h>
#include <errno.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <zlib.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "host_io.h"

namespace csqio {

thread_local char g_io_err[512] = "";

int io_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_io_err, sizeof(g_io_err), fmt, ap);
    va_end(ap);
    return code;
}
const char* io_error() { return g_io_err; }

// ---- pinned (or plain, when no CUDA context can be had) host memory ----------------------
// cudaHostAlloc pins ~2 GB/s (one thread faults, zeroes and locks every page); a whole-file run needs a few GB of
// batch buffers before its first batch, so large buffers are mapped here, first-touched by several threads and
// then registered (cudaHostRegister only has to lock and map pages that exist): 5 - 8 GB/s, and several buffers can
// be prepared side by side (measured: scripts/pin_probe2.py).
static std::atomic<int> g_pinning(1);  // cleared after the first failure (CPU-only host): do not retry per buffer
constexpr size_t MAP_THRESHOLD = 4u << 20;

static void first_touch(uint8_t* p, size_t bytes, int threads) {
    const size_t page = 4096;
    if (threads < 1) threads = 1;
    if ((size_t)threads > bytes / (8u << 20) + 1) threads = (int)(bytes / (8u << 20) + 1);
    const size_t per = ((bytes + (size_t)threads - 1) / (size_t)threads + page - 1) & ~(page - 1);
    auto work = [&](int t) {
        const size_t lo = (size_t)t * per, hi = lo + per < bytes ? lo + per : bytes;
        for (size_t o = lo; o < hi; o += page) ((volatile uint8_t*)p)[o] = 0;
    };
    std::vector<std::thread> helpers;
    for (int t = 1; t < threads; t++) helpers.emplace_back(work, t);
    work(0);
    for (auto& h : helpers) h.join();
}

PinnedBuf::~PinnedBuf() { release(); }
void PinnedBuf::release() {
    if (p) {
        if (kind == 1) {
            cudaFreeHost(p);
        } else if (kind == 2 || kind == 3) {
            if (kind == 2) cudaHostUnregister(p);
            munmap(p, map_len);
        } else {
            free(p);
        }
    }
    p = nullptr;
    cap = 0;
    kind = 0;
    pinned = false;
}
bool PinnedBuf::reserve(size_t bytes, size_t keep, int touch_threads) {
    if (bytes <= cap) return true;
    // pinned allocations are expensive: grow geometrically (x2) and never below 64 KiB
    size_t want = bytes > 2 * cap ? bytes : 2 * cap;
    if (want < (64u << 10)) want = 64u << 10;
    void* q = nullptr;
    int k = 0;
    size_t len = 0;
    const bool try_pin = g_pinning.load(std::memory_order_relaxed) != 0;
    if (try_pin && want >= MAP_THRESHOLD) {
        len = (want + ((2u << 20) - 1)) & ~(size_t)((2u << 20) - 1);
        void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (m != MAP_FAILED) {
            madvise(m, len, MADV_HUGEPAGE);
            first_touch((uint8_t*)m, len, touch_threads);
            if (cudaHostRegister(m, len, cudaHostRegisterPortable) == cudaSuccess) {
                q = m;
                k = 2;
            } else {
                cudaGetLastError();
                munmap(m, len);
            }
        }
    }
    if (!q && try_pin) {
        if (cudaHostAlloc(&q, want, cudaHostAllocPortable) == cudaSuccess) {
            k = 1;
        } else {
            cudaGetLastError();
            g_pinning.store(0, std::memory_order_relaxed);
            q = nullptr;
        }
    }
    if (!q) {
        q = malloc(want);
        k = 0;
    }
    if (!q) return false;
    if (p && keep) memcpy(q, p, keep);
    release();
    p = (uint8_t*)q;
    cap = want;
    kind = k;
    map_len = len;
    pinned = k == 1 || k == 2;
    return true;
}

// ---- input ------------------------------------------------------------------------------
ByteSource::~ByteSource() { close(); }
void ByteSource::close() {
    if (gz) gzclose((gzFile)gz);
    gz = nullptr;
}
int ByteSource::open(const char* path) {
    // gzopen reads plain files transparently and gzip files incl. concatenated members
    gzFile f = gzopen(path, "rb");
    if (!f) return io_fail(CSQ_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    gzbuffer(f, 1 << 20);
    gz = f;
    name = path;
    return 0;
}
long ByteSource::read(uint8_t* dst, size_t n) {
    int got = gzread((gzFile)gz, dst, (unsigned)n);
    if (got < 0) {
        int errnum = 0;
        const char* msg = gzerror((gzFile)gz, &errnum);
        io_fail(CSQ_ERR_IO, "read error in %s: %s", name.c_str(), msg ? msg : "?");
        return -1;
    }
    return got;
}

// FASTQ text -> SoA. dnaio semantics: 4 lines per record, '@' header, '+' separator line,
// equal sequence/quality lengths, '\r' before '\n' dropped, a missing final newline accepted.
int MateParser::open(const char* path) {
    eof = false;
    carry.clear();
    carry_pos = 0;
    line_no = 0;
    return src.open(path);
}

static inline size_t round16(size_t x) { return (x + 15) & ~(size_t)15; }

int parse_records(const uint8_t* text, size_t n_bytes, bool at_eof, uint32_t max_reads, MateSoA& out, size_t* consumed,
                  uint64_t* line_no, const char* fname) {
    size_t pos = 0;
    while (out.n < max_reads) {
        // find the four line ends of the next record
        size_t p = pos;
        const uint8_t* ls[4];
        size_t ll[4];
        int k = 0;
        for (; k < 4; k++) {
            if (p >= n_bytes) break;
            const uint8_t* nl = (const uint8_t*)memchr(text + p, '\n', n_bytes - p);
            size_t end;
            if (nl) {
                end = (size_t)(nl - text);
            } else if (at_eof && k == 3) {
                end = n_bytes;  // last record without a final newline
            } else {
                break;
            }
            ls[k] = text + p;
            ll[k] = end - p;
            if (ll[k] && ls[k][ll[k] - 1] == '\r') ll[k]--;
            p = nl ? end + 1 : end;
        }
        if (k < 4) {
            if (at_eof) {
                // only blank space may remain
                for (size_t q = pos; q < n_bytes; q++)
                    if (text[q] != '\n' && text[q] != '\r')
                        return io_fail(CSQ_ERR_FORMAT, "%s: FASTQ file ended prematurely (line %llu)", fname,
                                       (unsigned long long)(*line_no + 1));
                pos = n_bytes;
            }
            break;
        }
        if (ll[0] == 0 || ls[0][0] != '@')
            return io_fail(CSQ_ERR_FORMAT, "%s: line %llu is expected to start with '@'", fname, (unsigned long long)(*line_no + 1));
        if (ll[2] == 0 || ls[2][0] != '+')
            return io_fail(CSQ_ERR_FORMAT, "%s: line %llu is expected to start with '+'", fname, (unsigned long long)(*line_no + 3));
        if (ll[2] > 1 && (ll[2] != ll[0] || memcmp(ls[2] + 1, ls[0] + 1, ll[0] - 1) != 0))
            return io_fail(CSQ_ERR_FORMAT, "%s: sequence descriptions don't match at line %llu (the second description must be empty or equal to the first)",
                           fname, (unsigned long long)(*line_no + 3));
        if (ll[0] - 1 > 65535)
            return io_fail(CSQ_ERR_LIMIT, "%s: header at line %llu exceeds the supported 65535 bytes", fname, (unsigned long long)(*line_no + 1));
        if (ll[1] != ll[3])
            return io_fail(CSQ_ERR_FORMAT, "%s: length of sequence and qualities differ (record at line %llu)", fname,
                           (unsigned long long)(*line_no + 1));
        if (ll[1] > CSQ_MAX_READ_LEN)
            return io_fail(CSQ_ERR_LIMIT, "%s: read of %zu bases at line %llu exceeds the supported %d", fname, ll[1],
                           (unsigned long long)(*line_no + 1), CSQ_MAX_READ_LEN);
        const size_t slen = ll[1], nlen = ll[0] - 1;
        const size_t need_seq = out.seq_bytes + round16(slen) + 16, need_name = out.name_bytes + nlen + 16;
        if (need_seq >= (1ull << 32) || need_name >= (1ull << 32)) break;  // batch pools stay below 4 GiB
        if (!out.seq.reserve(need_seq, out.seq_bytes) || !out.qual.reserve(need_seq, out.seq_bytes) ||
            !out.name.reserve(need_name, out.name_bytes) || !out.seq_off.reserve((out.n + 2) * 4, out.n * 4) ||
            !out.seq_len.reserve((out.n + 2) * 4, out.n * 4) || !out.name_off.reserve((out.n + 3) * 4, (out.n + 1) * 4))
            return io_fail(CSQ_ERR_NOMEM, "out of host memory while parsing %s", fname);
        memcpy(out.seq.p + out.seq_bytes, ls[1], slen);
        memcpy(out.qual.p + out.seq_bytes, ls[3], slen);
        const size_t pad = round16(slen) - slen;
        if (pad) {
            memset(out.seq.p + out.seq_bytes + slen, 0, pad);
            memset(out.qual.p + out.seq_bytes + slen, 0, pad);
        }
        ((uint32_t*)out.seq_off.p)[out.n] = (uint32_t)out.seq_bytes;
        ((uint32_t*)out.seq_len.p)[out.n] = (uint32_t)slen;
        out.seq_bytes += round16(slen);
        memcpy(out.name.p + out.name_bytes, ls[0] + 1, nlen);
        ((uint32_t*)out.name_off.p)[out.n] = (uint32_t)out.name_bytes;
        out.name_bytes += nlen;
        ((uint32_t*)out.name_off.p)[out.n + 1] = (uint32_t)out.name_bytes;
        out.n++;
        out.total_bases += slen;
        *line_no += 4;
        pos = p;
    }
    *consumed = pos;
    return 0;
}

// Reserve room for a batch of `reads` records of about `read_len` bases up front, so that steady-state
// parsing never reallocates pinned memory.
bool MateSoA::presize(uint32_t reads, uint32_t read_len, uint32_t name_len) {
    const size_t stride = ((size_t)read_len + 15) / 16 * 16;
    return seq.reserve(stride * reads + 64, seq_bytes) && qual.reserve(stride * reads + 64, seq_bytes) &&
           name.reserve((size_t)name_len * reads + 64, name_bytes) && seq_off.reserve(((size_t)reads + 2) * 4, (size_t)n * 4) &&
           seq_len.reserve(((size_t)reads + 2) * 4, (size_t)n * 4) && name_off.reserve(((size_t)reads + 3) * 4, ((size_t)n + 1) * 4);
}

void MateSoA::clear() {
    n = 0;
    seq_bytes = name_bytes = 0;
    total_bases = 0;
    if (name_off.reserve(16, 0)) ((uint32_t*)name_off.p)[0] = 0;
}

void MateSoA::view(csq_mate_in* mi) const {
    mi->seq = seq.p;
    mi->qual = qual.p;
    mi->seq_off = (const uint32_t*)seq_off.p;
    mi->seq_len = (const uint32_t*)seq_len.p;
    mi->seq_bytes = seq_bytes;
    mi->name = name.p;
    mi->name_off = (const uint32_t*)name_off.p;
    mi->name_bytes = name_bytes;
}

int MateParser::next(uint32_t max_reads, MateSoA& out) {
    out.clear();
    if (hint_read_len && !out.presize(max_reads, hint_read_len, hint_name_len))
        return io_fail(CSQ_ERR_NOMEM, "out of host memory for a batch of %u reads", max_reads);
    const size_t CHUNK = 8u << 20;
    for (;;) {
        size_t used = 0;
        int rc = parse_records(carry.data() + carry_pos, carry.size() - carry_pos, eof, max_reads, out, &used, &line_no, src.name.c_str());
        if (rc) return rc;
        carry_pos += used;
        if (!hint_read_len && out.n >= 64) {
            // first batch of the file: size the pools from what the first records look like (+10 % slack)
            hint_read_len = (uint32_t)(out.total_bases / out.n) + 16;
            hint_read_len += hint_read_len / 10;
            hint_name_len = (uint32_t)(out.name_bytes / out.n) + 8;
            hint_name_len += hint_name_len / 4;
            if (!out.presize(max_reads, hint_read_len, hint_name_len))
                return io_fail(CSQ_ERR_NOMEM, "out of host memory for a batch of %u reads", max_reads);
        }
        if (out.n >= max_reads || eof) break;
        // need more text: drop what was consumed, append the next chunk
        if (carry_pos) {
            carry.erase(carry.begin(), carry.begin() + (long)carry_pos);
            carry_pos = 0;
        }
        const size_t have = carry.size();
        carry.resize(have + CHUNK);
        long got = src.read(carry.data() + have, CHUNK);
        if (got < 0) return CSQ_ERR_IO;
        carry.resize(have + (size_t)got);
        if (got == 0) eof = true;
    }
    return 0;
}

// ---- output -----------------------------------------------------------------------------
static bool ends_with(const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

int OutFile::open(const char* p, int level) {
    path = p;
    gzip = ends_with(path, ".gz");
    gz_level = level > 0 ? level : 1;  // cutadapt/xopen default compression level 1
    f = fopen(p, "wb");
    if (!f) return io_fail(CSQ_ERR_IO, "cannot create %s: %s", p, strerror(errno));
    setvbuf(f, nullptr, _IOFBF, 1 << 20);
    wrote_any = false;
    return 0;
}

// One complete gzip member per call; concatenated members are a valid gzip file.
int gzip_member(const uint8_t* src, size_t n, int level, std::vector<uint8_t>& dst) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return io_fail(CSQ_ERR_IO, "deflateInit2 failed");
    dst.resize(deflateBound(&zs, (uLong)n) + 64);
    zs.next_in = (Bytef*)src;
    zs.avail_in = (uInt)n;
    zs.next_out = dst.data();
    zs.avail_out = (uInt)dst.size();
    int rc = deflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END) {
        deflateEnd(&zs);
        return io_fail(CSQ_ERR_IO, "deflate failed (%d)", rc);
    }
    dst.resize(zs.total_out);
    deflateEnd(&zs);
    return 0;
}

int OutFile::write_raw(const uint8_t* data, size_t n) {
    if (!f || n == 0) return 0;
    if (fwrite(data, 1, n, f) != n) return io_fail(CSQ_ERR_IO, "write to %s failed: %s", path.c_str(), strerror(errno));
    wrote_any = true;
    return 0;
}

int OutFile::close() {
    if (!f) return 0;
    int rc = 0;
    if (gzip && !wrote_any) {  // an empty gzip file is still a valid (empty) member, like xopen writes
        std::vector<uint8_t> z;
        rc = gzip_member((const uint8_t*)"", 0, gz_level, z);
        if (!rc) rc = write_raw(z.data(), z.size());
    }
    if (fclose(f) != 0 && !rc) rc = io_fail(CSQ_ERR_IO, "closing %s failed: %s", path.c_str(), strerror(errno));
    f = nullptr;
    return rc;
}

}  // namespace csqio

// ---- C ABI: reader and in-memory parser ---------------------------------------------------
using namespace csqio;

struct csq_reader {
    MateParser parser[2];
    int n_mates = 1;
    MateSoA soa[CSQ_N_SLOTS][2];
};

extern const char* csq_last_error(void);
void csq_set_error(const char* msg);  // plan.cu

static int report(int rc) {
    if (rc) csq_set_error(io_error());
    return rc;
}

extern "C" int csq_reader_open(const char* path1, const char* path2, csq_reader** out) {
    if (!path1 || !out) {
        csq_set_error("null argument");
        return CSQ_ERR_INVALID;
    }
    csq_reader* r = new csq_reader();
    r->n_mates = path2 ? 2 : 1;
    int rc = r->parser[0].open(path1);
    if (!rc && path2) rc = r->parser[1].open(path2);
    if (rc) {
        delete r;
        return report(rc);
    }
    *out = r;
    return 0;
}

int csq_reader_next_into(csq_reader* r, csqio::MateSoA* soa /*[2]*/, uint32_t max_reads, csq_batch_in* in, double* seconds) {
    int rcs[2] = {0, 0};
    char errs[2][512] = {"", ""};
    auto work = [&](int m) {
        rcs[m] = r->parser[m].next(max_reads, soa[m]);
        if (rcs[m]) snprintf(errs[m], sizeof(errs[m]), "%s", io_error());
    };
    if (r->n_mates == 2) {
        std::thread t(work, 1);
        work(0);
        t.join();
    } else {
        work(0);
    }
    (void)seconds;
    for (int m = 0; m < r->n_mates; m++)
        if (rcs[m]) {
            csq_set_error(errs[m]);
            return rcs[m];
        }
    if (r->n_mates == 2 && soa[0].n != soa[1].n) {
        csq_set_error("paired input files have different numbers of records");
        return CSQ_ERR_FORMAT;
    }
    memset(in, 0, sizeof(*in));
    in->n_reads = soa[0].n;
    in->n_mates = (uint32_t)r->n_mates;
    for (int m = 0; m < r->n_mates; m++) soa[m].view(&in->mate[m]);
    return 0;
}

extern "C" int csq_reader_next(csq_reader* r, int buffer, uint32_t max_reads, csq_batch_in* in) {
    if (!r || !in || buffer < 0 || buffer >= CSQ_N_SLOTS || max_reads == 0) {
        csq_set_error("bad argument");
        return CSQ_ERR_INVALID;
    }
    return csq_reader_next_into(r, r->soa[buffer], max_reads, in, nullptr);
}

extern "C" void csq_reader_close(csq_reader* r) { delete r; }

extern "C" int csq_parse_fastq_mem(const uint8_t* text, uint64_t n_bytes, uint32_t max_reads, uint8_t* seq, uint8_t* qual,
                                   uint64_t seq_cap, uint32_t* seq_off, uint32_t* seq_len, uint8_t* name, uint64_t name_cap,
                                   uint32_t* name_off, uint32_t* n_reads, uint64_t* seq_bytes, uint64_t* consumed) {
    if (!text || !seq || !qual || !seq_off || !seq_len || !name || !name_off || !n_reads) {
        csq_set_error("null argument");
        return CSQ_ERR_INVALID;
    }
    MateSoA soa;
    soa.clear();
    size_t used = 0;
    uint64_t line_no = 0;
    int rc = parse_records(text, (size_t)n_bytes, true, max_reads, soa, &used, &line_no, "<memory>");
    if (rc) return report(rc);
    if (soa.seq_bytes > seq_cap || soa.name_bytes > name_cap) {
        csq_set_error("caller buffers too small");
        return CSQ_ERR_CAPACITY;
    }
    if (soa.n) {
        memcpy(seq, soa.seq.p, soa.seq_bytes);
        memcpy(qual, soa.qual.p, soa.seq_bytes);
        memcpy(seq_off, soa.seq_off.p, (size_t)soa.n * 4);
        memcpy(seq_len, soa.seq_len.p, (size_t)soa.n * 4);
        memcpy(name, soa.name.p, soa.name_bytes);
    }
    memcpy(name_off, soa.name_off.p, ((size_t)soa.n + 1) * 4);
    *n_reads = soa.n;
    if (seq_bytes) *seq_bytes = soa.seq_bytes;
    if (consumed) *consumed = used;
    return 0;
}
