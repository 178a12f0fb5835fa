// Whole-file driver (csq_run_files): reader thread (raw FASTQ bytes, cut at record boundaries;
// the device parses them) -> per-GPU worker threads (two slots each, double-buffered
// csq_submit_text / csq_wait) -> ordered writer thread (parallel gzip members).
// The B200 analogue of cutadapt's reader / worker / ordered-writer runner behind
// runner.run(pipeline, Progress(), outfiles) in reference run.py:436-473 / 753-794: batches are
// contiguous record ranges, they are dealt to the GPUs in order and written back in input order.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_io.h"

void csq_set_error(const char* msg);
struct csq_text_reader;
int csq_text_reader_next_into(csq_text_reader* r, csqio::PinnedBuf* bufs, uint32_t max_reads, csq_batch_text* in);
extern "C" int csq_text_reader_open(const char* path1, const char* path2, csq_text_reader** out);
extern "C" void csq_text_reader_close(csq_text_reader* r);

namespace {

using Clock = std::chrono::steady_clock;
double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

template <typename T>
class Queue {
   public:
    void push(T v) {
        {
            std::lock_guard<std::mutex> g(m_);
            q_.push_back(std::move(v));
        }
        cv_.notify_one();
    }
    // returns false when the queue is closed and empty
    bool pop(T& out) {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop_front();
        return true;
    }
    void close() {
        {
            std::lock_guard<std::mutex> g(m_);
            closed_ = true;
        }
        cv_.notify_all();
    }

   private:
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<T> q_;
    bool closed_ = false;
};

struct Job {
    long index = -1;
    csqio::PinnedBuf text[2];
    csq_batch_text in;
    csqio::PinnedBuf outbuf[CSQ_N_DEST][2];
    csq_batch_out out;
    int slot = 0;
    float total_ms = 0, kernel_ms = 0;
};

struct Job;

struct Shared {
    std::mutex err_m;
    int err_code = 0;
    std::string err_msg;
    std::atomic<bool> stop{false};
    Queue<Job*>*free_q = nullptr, *ready_q = nullptr;
    void fail(int code, const char* msg) {
        {
            std::lock_guard<std::mutex> g(err_m);
            if (!err_code) {
                err_code = code;
                err_msg = msg ? msg : "";
            }
            stop = true;
        }
        // unblock the reader and the workers
        if (free_q) free_q->close();
        if (ready_q) ready_q->close();
    }
};

bool size_job_outputs(Job& j, bool first_try) {
    const int n_mates = (int)j.in.n_mates;
    for (int m = 0; m < 2; m++) {
        uint64_t full = 64;
        if (m < n_mates) full = j.in.mate[m].bytes + 64ull * j.in.n_reads + 4096;  // renaming can only shorten a record, bar "_" + UMI
        for (int d = 0; d < CSQ_N_DEST; d++) {
            uint64_t want = first_try ? (d == CSQ_DEST_TRIMMED ? full : full / 8 + 4096) : j.out.text[d][m].bytes + 4096;
            if (m >= n_mates) want = 64;
            // (allocate with 1/8 to spare: a batch a few bytes longer than the last must not cost a new pinned buffer)
            if (want > j.outbuf[d][m].cap && !j.outbuf[d][m].reserve(want + want / 8, 0)) return false;
            j.out.text[d][m].data = j.outbuf[d][m].p;
            j.out.text[d][m].capacity = j.outbuf[d][m].cap;
        }
    }
    return true;
}

}  // namespace

extern "C" int csq_run_files(const csq_op* ops_r1, int n1, const csq_op* ops_r2, int n2, const csq_filters* filters,
                             uint32_t plan_flags, const csq_files* files, csq_counters* counters, csq_timing* timing) {
    if (!files || !files->in[0] || !filters) {
        csq_set_error("null argument");
        return CSQ_ERR_INVALID;
    }
    const auto t_start = Clock::now();
    const bool trace = getenv("CSQ_TRACE") != nullptr;  // phase timestamps on stderr
    auto stamp = [&](const char* what) {
        if (trace) fprintf(stderr, "[csq_run_files] %8.3f s  %s\n", seconds_since(t_start), what);
    };
    const int n_mates = files->in[1] ? 2 : 1;
    if ((n_mates == 2) != (n2 > 0)) {
        csq_set_error("number of input files does not match the program (paired vs single-end)");
        return CSQ_ERR_INVALID;
    }
    const int n_dev = files->n_devices > 0 ? files->n_devices : 1;
    const uint32_t batch_reads = files->batch_reads ? files->batch_reads : (1u << 16);
    const int n_threads = files->n_threads > 0 ? files->n_threads : 4;

    // plans first: fails loudly when there is no usable GPU
    std::vector<csq_plan*> plans((size_t)n_dev, nullptr);
    auto destroy_plans = [&] {
        for (csq_plan* p : plans)
            if (p) csq_plan_destroy(p);
    };
    for (int d = 0; d < n_dev; d++) {
        const int dev = files->devices ? files->devices[d] : d;
        int rc = csq_plan_create(ops_r1, n1, ops_r2, n2, filters, dev, plan_flags, &plans[(size_t)d]);
        if (rc) {
            destroy_plans();
            return rc;
        }
    }
    stamp("plans created");
    csq_text_reader* reader = nullptr;
    int rc = csq_text_reader_open(files->in[0], files->in[1], &reader);
    if (rc) {
        destroy_plans();
        return rc;
    }
    csqio::OutFile outs[CSQ_N_DEST][2];
    for (int d = 0; d < CSQ_N_DEST && !rc; d++)
        for (int m = 0; m < n_mates && !rc; m++)
            if (files->out[d][m]) {
                rc = outs[d][m].open(files->out[d][m], files->gzip_level);
                if (rc) csq_set_error(csqio::io_error());
            }
    if (rc) {
        csq_text_reader_close(reader);
        destroy_plans();
        return rc;
    }

    Shared sh;
    // batches in circulation: one with the reader, two per GPU in flight, one with the writer, and one spare per stage
    // boundary so that a slow batch in one stage does not stall the others
    // (every job pins ~100 MB of host memory, ~0.1 s each in a VM: short inputs get by with the minimum of one per stage)
    uint64_t in_bytes = 0;
    for (int m = 0; m < n_mates; m++) {
        struct stat sb;
        if (stat(files->in[m], &sb) == 0) in_bytes += (uint64_t)sb.st_size * (strlen(files->in[m]) > 3 && !strcmp(files->in[m] + strlen(files->in[m]) - 3, ".gz") ? 4 : 1);
    }
    const int n_jobs = in_bytes < (3ull << 30) ? 2 * n_dev + 2 : 3 * n_dev + 3;
    std::vector<std::unique_ptr<Job>> jobs;
    Queue<Job*> free_q, ready_q, done_q;
    sh.free_q = &free_q;
    sh.ready_q = &ready_q;
    for (int i = 0; i < n_jobs; i++) {
        jobs.emplace_back(new Job());
        free_q.push(jobs.back().get());
    }
    double t_read = 0, t_write = 0;
    std::atomic<long> n_batches{0};

    std::thread reader_thread([&] {
        long index = 0;
        Job* j = nullptr;
        while (!sh.stop && free_q.pop(j)) {
            const auto t0 = Clock::now();
            int r = csq_text_reader_next_into(reader, j->text, batch_reads, &j->in);
            t_read += seconds_since(t0);
            if (r) {
                sh.fail(r, csq_last_error());
                break;
            }
            if (j->in.n_reads == 0) {
                free_q.push(j);
                break;
            }
            j->index = index++;
            if (trace) fprintf(stderr, "[csq_run_files] %8.3f s  batch %ld read in %.1f ms\n", seconds_since(t_start), j->index, seconds_since(t0) * 1e3);
            ready_q.push(j);
        }
        n_batches = index;
        stamp("reader done");
        ready_q.close();
    });

    std::vector<std::thread> workers;
    std::vector<double> gpu_total((size_t)n_dev, 0.0), gpu_kernel((size_t)n_dev, 0.0);
    for (int d = 0; d < n_dev; d++) {
        workers.emplace_back([&, d] {
            csq_plan* plan = plans[(size_t)d];
            std::deque<Job*> inflight;
            int next_slot = 0;
            auto finish_one = [&]() -> bool {
                Job* j = inflight.front();
                inflight.pop_front();
                int r = csq_wait(plan, j->slot);
                if (r == CSQ_ERR_CAPACITY) {  // grow to the reported sizes and fetch again
                    if (!size_job_outputs(*j, false)) {
                        sh.fail(CSQ_ERR_NOMEM, "out of host memory for output buffers");
                        return false;
                    }
                    r = csq_wait(plan, j->slot);
                }
                if (r) {
                    sh.fail(r, csq_last_error());
                    return false;
                }
                csq_slot_times(plan, j->slot, &j->total_ms, &j->kernel_ms);
                gpu_total[(size_t)d] += j->total_ms * 1e-3;
                gpu_kernel[(size_t)d] += j->kernel_ms * 1e-3;
                done_q.push(j);
                return true;
            };
            Job* j = nullptr;
            while (!sh.stop && ready_q.pop(j)) {
                if (!size_job_outputs(*j, true)) {
                    sh.fail(CSQ_ERR_NOMEM, "out of host memory for output buffers");
                    break;
                }
                j->slot = next_slot;
                next_slot ^= 1;
                int r = csq_submit_text(plan, j->slot, &j->in, &j->out);
                if (r) {
                    sh.fail(r, csq_last_error());
                    break;
                }
                if (trace) fprintf(stderr, "[csq_run_files] %8.3f s  batch %ld submitted\n", seconds_since(t_start), j->index);
                inflight.push_back(j);
                if (inflight.size() == 2 && !finish_one()) break;
            }
            while (!sh.stop && !inflight.empty())
                if (!finish_one()) break;
        });
    }

    std::thread writer_thread([&] {
        std::map<long, Job*> pending;
        long next = 0;
        Job* j = nullptr;
        while (done_q.pop(j)) {
            pending[j->index] = j;
            while (!pending.empty() && pending.begin()->first == next) {
                Job* w = pending.begin()->second;
                pending.erase(pending.begin());
                next++;
                if (!sh.stop) {
                    const auto t0 = Clock::now();
                    // tasks: (dest, mate, chunk) -> optional gzip member; written in order afterwards
                    struct Task {
                        int d, m;
                        const uint8_t* src;
                        size_t n;
                        std::vector<uint8_t> z;
                        int rc = 0;
                    };
                    std::vector<Task> tasks;
                    const size_t CH = 4u << 20;
                    for (int d = 0; d < CSQ_N_DEST; d++)
                        for (int m = 0; m < n_mates; m++) {
                            // paired --auto-rc on '-' strand: R1 goes to the R2 file and vice versa (trimmed only)
                            const int fm = (d == CSQ_DEST_TRIMMED && files->swap_sink && n_mates == 2) ? 1 - m : m;
                            if (!outs[d][fm].f) continue;
                            const csq_text_out& t = w->out.text[d][m];
                            for (size_t off = 0; off < t.bytes; off += CH) {
                                Task k;
                                k.d = d;
                                k.m = fm;
                                k.src = t.data + off;
                                k.n = (size_t)std::min<uint64_t>(CH, t.bytes - off);
                                tasks.push_back(std::move(k));
                            }
                        }
                    std::atomic<size_t> cursor{0};
                    auto compress = [&] {
                        for (;;) {
                            size_t i = cursor++;
                            if (i >= tasks.size()) return;
                            Task& k = tasks[i];
                            if (outs[k.d][k.m].gzip) k.rc = csqio::gzip_member(k.src, k.n, outs[k.d][k.m].gz_level, k.z);
                        }
                    };
                    std::vector<std::thread> pool;
                    const int nt = (int)std::min<size_t>((size_t)n_threads, tasks.size());
                    for (int t = 1; t < nt; t++) pool.emplace_back(compress);
                    compress();
                    for (auto& th : pool) th.join();
                    // the files are independent: one writing thread per output file, its tasks in order
                    std::vector<std::thread> wpool;
                    std::mutex werr_m;
                    auto write_file = [&](int d, int m) {
                        for (Task& k : tasks) {
                            if (k.d != d || k.m != m) continue;
                            int r = k.rc;
                            if (!r) r = outs[d][m].gzip ? outs[d][m].write_raw(k.z.data(), k.z.size()) : outs[d][m].write_raw(k.src, k.n);
                            if (r) {
                                std::lock_guard<std::mutex> g(werr_m);
                                sh.fail(r, csqio::io_error());
                                return;
                            }
                        }
                    };
                    bool first_file = true;
                    int fd0 = -1, fm0 = -1;
                    for (int d = 0; d < CSQ_N_DEST; d++)
                        for (int m = 0; m < n_mates; m++) {
                            bool any = false;
                            for (Task& k : tasks) any = any || (k.d == d && k.m == m);
                            if (!any) continue;
                            if (first_file) {
                                first_file = false;
                                fd0 = d;
                                fm0 = m;
                            } else {
                                wpool.emplace_back(write_file, d, m);
                            }
                        }
                    if (fd0 >= 0) write_file(fd0, fm0);
                    for (auto& th : wpool) th.join();
                    t_write += seconds_since(t0);
                    if (trace) fprintf(stderr, "[csq_run_files] %8.3f s  batch %ld written\n", seconds_since(t_start), w->index);
                }
                free_q.push(w);
            }
        }
    });

    reader_thread.join();
    for (auto& w : workers) w.join();
    stamp("workers done");
    done_q.close();
    writer_thread.join();
    stamp("writer done");
    free_q.close();

    for (int d = 0; d < CSQ_N_DEST; d++)
        for (int m = 0; m < 2; m++) {
            int r = outs[d][m].close();
            if (r && !sh.err_code) sh.fail(r, csqio::io_error());
        }
    if (counters) {
        memset(counters, 0, sizeof(*counters));
        for (csq_plan* p : plans) {
            csq_counters c;
            if (csq_stats(p, &c) == 0) {
                uint64_t* dst = (uint64_t*)counters;
                const uint64_t* src = (const uint64_t*)&c;
                for (size_t w = 0; w < sizeof(csq_counters) / sizeof(uint64_t); w++) dst[w] += src[w];
            }
        }
    }
    if (timing) {
        memset(timing, 0, sizeof(*timing));
        timing->read_inflate = t_read;  // inflate + parse of both mates (mates run on two threads)
        timing->parse = 0;
        for (int d = 0; d < n_dev; d++) {
            timing->h2d_kernels_d2h += gpu_total[(size_t)d];
            timing->kernels += gpu_kernel[(size_t)d];
        }
        timing->write_deflate = t_write;
        timing->total = seconds_since(t_start);
    }
    stamp("outputs closed");
    // tear down side by side: un-pinning the jobs' host buffers and freeing the plans' device buffers are both slow
    // driver calls (together ~0.3 s behind a 2 M-pair run when done one after the other)
    {
        std::thread unpin([&] { jobs.clear(); });
        csq_text_reader_close(reader);
        destroy_plans();
        unpin.join();
    }
    stamp("teardown done");
    if (sh.err_code) {
        csq_set_error(sh.err_msg.c_str());
        return sh.err_code;
    }
    return 0;
}
