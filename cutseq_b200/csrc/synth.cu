// Synthetic workload generator for BASELINE.json configs 2-5.
// Record i depends only on (seed, first_index + i): any rank can regenerate any range.
// Output: packed SoA batches in library-owned pinned host memory (csq_batch_in layout).
// Host code only (multi-threaded); it feeds both the CUDA chain and the CPU oracle.
#ifndef CSQ_SYNTH_NO_CUDA
#include <cuda_runtime.h>
#else
// oracle/libsynth.so (g++ -x c++ -DCSQ_SYNTH_NO_CUDA): the same generator with plain host memory, so that the CPU arm
// of bench.py (--impl reference) maps nothing of the product library
#include <stdlib.h>
static int cudaFreeHost(void*) { return 1; }
static int cudaHostAlloc(void**, size_t, unsigned) { return 1; }
static int cudaGetLastError() { return 0; }
#define cudaSuccess 0
#define cudaHostAllocDefault 0u
#endif
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cutseq_b200.h"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }  // [0, n)
    bool chance(uint32_t per_100k) { return below(100000) < per_100k; }
    int range(int lo, int hi) { return lo + (int)below((uint32_t)(hi - lo + 1)); }  // inclusive
};

const char BASES[4] = {'A', 'C', 'G', 'T'};

char comp(char c) {
    switch (c) {
        case 'A': return 'T';
        case 'C': return 'G';
        case 'G': return 'C';
        case 'T': return 'A';
    }
    return c;
}

std::string revcomp(const std::string& s) {
    std::string r(s.rbegin(), s.rend());
    for (char& c : r) c = comp(c);
    return r;
}

struct Layout {  // library layout of one config
    std::string p5rc_tail, p7_tail;  // what follows the fragment in R2 / R1
    std::string inline5, inline3;
    int umi5 = 0, mask5 = 0, mask3 = 0, umi3 = 0;
    int readthrough_per_100k = 35000;
    int polyt_per_100k = 5000;
    bool minus_strand = true;
    int rt_min = 10, rt_mode = 130, rt_max = 149;  // read-through insert length (triangular)
    int long_min = 150, long_max = 400;
    int sub = 500, indel = 50, nrate = 100;  // per 100k bases
    int adapter_sub = 0;                     // extra error rate inside the adapter (config 4)
    int bc_err = 0, wrong_bc = 0;            // config 3
};

Layout layout_for(const csq_synth& cfg) {
    Layout L;
    const std::string p5 = "ACACGACGCTCTTCCGATCT", p7 = "AGATCGGAAGAGCACACGTC";
    L.p7_tail = p7 + "TGAACTCCAGTCAC";
    L.p5rc_tail = revcomp(p5) + "AGATCTCGGTGGTCGCCGTATCATT";
    if (cfg.config == 3) {  // ACACGACGCTCTTCCGATCT(ATCACG)NNNNNNNNXXX<XXX(CGTGAT)AGATCGGAAGAGCACACGTC
        L.inline5 = "ATCACG";
        L.inline3 = "CGTGAT";
        L.umi5 = 8;
        L.mask5 = 3;
        L.mask3 = 3;
        L.readthrough_per_100k = 30000;
        L.bc_err = 2000;
        L.wrong_bc = 3000;
    } else if (cfg.config == 4) {  // SMALLRNA, single-end 75 nt, insert 20-40 nt
        L.p7_tail = p7 + "ATCTCGTATGCCGTCTTCTGCTTG";
        L.readthrough_per_100k = 100000;
        L.polyt_per_100k = 0;
        L.minus_strand = false;
        L.rt_min = 20;
        L.rt_mode = 30;
        L.rt_max = 40;
        L.adapter_sub = 10000;
    } else {  // configs 2 / 5: TAKARAV3 ...XXX<XXXXXXNNNNNNNN...
        L.mask5 = 3;
        L.mask3 = 6;
        L.umi3 = 8;
    }
    return L;
}

void rand_bases(Rng& r, std::string& out, int n) {
    for (int i = 0; i < n; i++) out.push_back(BASES[r.below(4)]);
}

// integer triangular sample on [lo, hi] with the given mode
int triangular(Rng& r, int lo, int mode, int hi) {
    const int left = mode - lo, right = hi - mode;
    if ((int)r.below((uint32_t)(left + right + 1)) <= left) {
        int a = r.range(0, left), b = r.range(0, left);
        return lo + std::max(a, b);
    }
    int a = r.range(0, right), b = r.range(0, right);
    return mode + std::min(a, b);
}

// sequencing errors: substitutions, indels, N calls
void sequence(Rng& r, const std::string& tmpl, int read_len, const Layout& L, int adapter_from, std::string& out) {
    out.clear();
    size_t i = 0;
    while ((int)out.size() < read_len && i < tmpl.size()) {
        const char ch = tmpl[i];
        const int sub = L.sub + (((int)i >= adapter_from) ? L.adapter_sub : 0);
        const uint32_t x = r.below(100000);
        if ((int)x < sub * 8 / 10 || ((int)x < sub && L.adapter_sub == 0)) {
            char c;
            do c = BASES[r.below(4)];
            while (c == ch);
            out.push_back(c);
            i++;
        } else if ((int)x < sub) {  // adapter-region indels of config 4 (sub:ins:del = 8:1:1)
            if (x & 1) {
                i++;
            } else {
                out.push_back(BASES[r.below(4)]);
            }
        } else if ((int)x < sub + L.indel) {
            i++;  // deletion
        } else if ((int)x < sub + 2 * L.indel) {
            out.push_back(BASES[r.below(4)]);  // insertion (template position not consumed)
        } else if ((int)x < sub + 2 * L.indel + L.nrate) {
            out.push_back('N');
            i++;
        } else {
            out.push_back(ch);
            i++;
        }
    }
    while ((int)out.size() < read_len) out.push_back('G');
    out.resize((size_t)read_len);
}

void qualities(Rng& r, const std::string& seq, std::string& q) {
    q.resize(seq.size());
    for (size_t i = 0; i < seq.size(); i++) {
        const uint32_t x = r.below(100);
        q[i] = x < 84 ? 'I' : x < 92 ? '9' : '-';
    }
    if (r.below(5) == 0) {  // 20 %: low-quality 3' tail, geometric with mean 15
        int tail = 0;
        while (r.below(15) != 0 && tail < (int)seq.size()) tail++;
        for (int i = (int)seq.size() - tail; i < (int)seq.size(); i++) q[(size_t)i] = (r.below(2) ? '-' : '#');
    }
    for (size_t i = 0; i < seq.size(); i++)
        if (seq[i] == 'N') q[i] = '#';
}

void mutate_barcode(Rng& r, std::string& bc, const Layout& L) {
    if (bc.empty()) return;
    if (r.chance((uint32_t)L.wrong_bc)) {
        for (char& c : bc) c = BASES[r.below(4)];
        return;
    }
    for (char& c : bc)
        if (r.chance((uint32_t)L.bc_err)) c = BASES[r.below(4)];
}

struct Record {
    std::string name[2], seq[2], qual[2];
};

void make_record(const csq_synth& cfg, const Layout& L, uint64_t index, Record& rec) {
    // counter-based: hash (seed, index) into the stream's start state so that neighbouring
    // records do not walk the same splitmix lattice
    uint64_t h = cfg.seed * 0xD1B54A32D192ED03ull + index * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull;
    h = (h ^ (h >> 32)) * 0xD6E8FEB86659FD93ull;
    h = (h ^ (h >> 32)) * 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    Rng r(h);
    const int read_len = (int)cfg.read_len;
    std::string frag;
    std::string b5 = L.inline5, b3 = L.inline3;
    mutate_barcode(r, b5, L);
    mutate_barcode(r, b3, L);
    frag += b5;
    rand_bases(r, frag, L.umi5);
    rand_bases(r, frag, L.mask5);
    int ins_len = r.chance((uint32_t)L.readthrough_per_100k) ? triangular(r, L.rt_min, L.rt_mode, L.rt_max)
                                                             : r.range(L.long_min, L.long_max);
    std::string insert;
    rand_bases(r, insert, ins_len);
    if (L.polyt_per_100k && r.chance((uint32_t)L.polyt_per_100k)) {
        const int run = r.range(10, 40);
        if (L.minus_strand)
            insert = std::string((size_t)run, 'T') + insert;
        else
            insert += std::string((size_t)run, 'A');
    }
    frag += insert;
    rand_bases(r, frag, L.mask3);
    rand_bases(r, frag, L.umi3);
    frag += b3;
    const std::string t1 = frag + L.p7_tail + std::string((size_t)read_len, 'G');
    sequence(r, t1, read_len, L, (int)frag.size(), rec.seq[0]);
    qualities(r, rec.seq[0], rec.qual[0]);
    if (cfg.paired) {
        const std::string t2 = revcomp(frag) + L.p5rc_tail + std::string((size_t)read_len, 'G');
        sequence(r, t2, read_len, L, (int)frag.size(), rec.seq[1]);
        qualities(r, rec.seq[1], rec.qual[1]);
    }
    char buf[96];
    const unsigned tile = 1101 + (unsigned)((index / 100000) % 1000);
    const unsigned x = 1000 + r.below(30000), y = 1000 + r.below(30000);
    for (int m = 0; m < (cfg.paired ? 2 : 1); m++) {
        snprintf(buf, sizeof(buf), "SIM:1:FC:1:%u:%u:%u %d:N:0:ACGTACGT+TGCATGCA", tile, x, y, m + 1);
        rec.name[m] = buf;
    }
}

struct HostBatch {  // pinned SoA of one buffer index
    uint8_t *seq[2] = {nullptr, nullptr}, *qual[2] = {nullptr, nullptr}, *name[2] = {nullptr, nullptr};
    uint32_t *seq_off[2] = {nullptr, nullptr}, *seq_len[2] = {nullptr, nullptr}, *name_off[2] = {nullptr, nullptr};
    size_t seq_cap[2] = {0, 0}, name_cap[2] = {0, 0}, n_cap[2] = {0, 0};
};

constexpr int N_BUFFERS = 16;
HostBatch g_buffers[N_BUFFERS];

bool ensure_pinned(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return true;
    if (*p) cudaFreeHost(*p);
    *p = nullptr;
    *cap = 0;
    if (cudaHostAlloc(p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        // no CUDA context available (CPU-only host): plain memory still serves the oracle/tests
        *p = malloc(bytes);
        if (!*p) return false;
    }
    *cap = bytes;
    return true;
}

}  // namespace

extern "C" int csq_synth_batch(const csq_synth* cfg, uint64_t first_index, uint32_t n_reads, int buffer, csq_batch_in* in) {
    if (!cfg || !in || buffer < 0 || buffer >= N_BUFFERS || cfg->read_len < 20 || cfg->read_len > CSQ_MAX_READ_LEN)
        return CSQ_ERR_INVALID;
    const Layout L = layout_for(*cfg);
    const int n_mates = cfg->paired ? 2 : 1;
    HostBatch& hb = g_buffers[buffer];
    const size_t stride = ((size_t)cfg->read_len + 15) / 16 * 16;
    const size_t name_stride = 64;  // upper bound while generating; compacted below
    std::vector<std::string> names[2];
    for (int m = 0; m < n_mates; m++) {
        size_t c;
        c = hb.seq_cap[m];
        if (!ensure_pinned((void**)&hb.seq[m], &c, stride * n_reads + 16)) return CSQ_ERR_NOMEM;
        c = hb.seq_cap[m];
        if (!ensure_pinned((void**)&hb.qual[m], &c, stride * n_reads + 16)) return CSQ_ERR_NOMEM;
        hb.seq_cap[m] = c;
        c = hb.name_cap[m];
        if (!ensure_pinned((void**)&hb.name[m], &c, name_stride * n_reads + 16)) return CSQ_ERR_NOMEM;
        hb.name_cap[m] = c;
        size_t c1 = hb.n_cap[m], c2 = hb.n_cap[m], c3 = hb.n_cap[m];
        if (!ensure_pinned((void**)&hb.seq_off[m], &c1, ((size_t)n_reads + 1) * 4)) return CSQ_ERR_NOMEM;
        if (!ensure_pinned((void**)&hb.seq_len[m], &c2, ((size_t)n_reads + 1) * 4)) return CSQ_ERR_NOMEM;
        if (!ensure_pinned((void**)&hb.name_off[m], &c3, ((size_t)n_reads + 1) * 4)) return CSQ_ERR_NOMEM;
        hb.n_cap[m] = c1;
    }
    // names have variable length: generate per-thread into fixed 64-byte cells, then compact
    std::vector<uint8_t> name_cells[2];
    std::vector<uint8_t> name_lens[2];
    for (int m = 0; m < n_mates; m++) {
        name_cells[m].resize(name_stride * (size_t)n_reads);
        name_lens[m].resize(n_reads);
    }
    unsigned hw = std::thread::hardware_concurrency();
    const unsigned n_threads = std::max(1u, std::min(hw ? hw : 1u, 32u));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n_threads; t++) {
        pool.emplace_back([&, t]() {
            Record rec;
            const uint32_t lo = (uint32_t)((uint64_t)n_reads * t / n_threads), hi = (uint32_t)((uint64_t)n_reads * (t + 1) / n_threads);
            for (uint32_t i = lo; i < hi; i++) {
                make_record(*cfg, L, first_index + i, rec);
                for (int m = 0; m < n_mates; m++) {
                    const size_t off = stride * i;
                    memcpy(hb.seq[m] + off, rec.seq[m].data(), rec.seq[m].size());
                    memcpy(hb.qual[m] + off, rec.qual[m].data(), rec.qual[m].size());
                    memset(hb.seq[m] + off + rec.seq[m].size(), 0, stride - rec.seq[m].size());
                    memset(hb.qual[m] + off + rec.seq[m].size(), 0, stride - rec.seq[m].size());
                    hb.seq_off[m][i] = (uint32_t)off;
                    hb.seq_len[m][i] = (uint32_t)rec.seq[m].size();
                    const size_t nl = std::min(rec.name[m].size(), name_stride);
                    memcpy(name_cells[m].data() + name_stride * i, rec.name[m].data(), nl);
                    name_lens[m][i] = (uint8_t)nl;
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    memset(in, 0, sizeof(*in));
    in->n_reads = n_reads;
    in->n_mates = (uint32_t)n_mates;
    for (int m = 0; m < n_mates; m++) {
        uint32_t pos = 0;
        for (uint32_t i = 0; i < n_reads; i++) {
            hb.name_off[m][i] = pos;
            memcpy(hb.name[m] + pos, name_cells[m].data() + name_stride * i, name_lens[m][i]);
            pos += name_lens[m][i];
        }
        hb.name_off[m][n_reads] = pos;
        csq_mate_in& mi = in->mate[m];
        mi.seq = hb.seq[m];
        mi.qual = hb.qual[m];
        mi.seq_off = hb.seq_off[m];
        mi.seq_len = hb.seq_len[m];
        mi.seq_bytes = stride * n_reads;
        mi.name = hb.name[m];
        mi.name_off = hb.name_off[m];
        mi.name_bytes = pos;
    }
    return 0;
}

extern "C" void csq_synth_free(void) {
    // buffers may be pinned or malloc'ed (CPU-only host); cudaFreeHost fails harmlessly on the latter
    for (HostBatch& hb : g_buffers) {
        for (int m = 0; m < 2; m++) {
            void* ptrs[6] = {hb.seq[m], hb.qual[m], hb.name[m], hb.seq_off[m], hb.seq_len[m], hb.name_off[m]};
            for (void* p : ptrs)
                if (p && cudaFreeHost(p) != cudaSuccess) {
                    cudaGetLastError();
                    free(p);
                }
        }
        hb = HostBatch();
    }
}
