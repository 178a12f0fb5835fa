#!/usr/bin/env python
"""bench.py - read-pairs/s of the trimming chain on BASELINE.json's workload.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this implementation
    python bench.py --impl reference [...]                       # CPU arm (oracle port, all host cores)
    torchrun --nproc-per-node N bench.py --gpus N ...            # one rank per GPU

One "step" = one pass of the whole hot path (all ALIGN ops, cuts, quality trim, filters, FASTQ
emission) over one batch of synthetic 2x150 read pairs (config 2: `-A TAKARAV3 --trim-polyA`).
`value` is kernel throughput with the batches resident in HBM (CUDA events around exactly K steps,
max over ranks); `e2e` is the same chain through csq_submit_text/csq_wait with pinned HOST buffers,
host->device and device->host copies inside the timed region.  Prints ONE JSON line (rank 0).
Input form: --input text (default; raw FASTQ bytes, the device builds the record index - what the
file driver does) or --input soa (host-parsed packed struct-of-arrays batches).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read-pairs/s (2x150, TAKARAV3)"
UNIT = "pairs/s"
BATCH_PAIRS = 2_000_000      # pairs per step and GPU
N_BATCHES = 5                # distinct resident batches per GPU: 5 x 2M = the 10M pairs of config 2
ARGV = ["-A", "TAKARAV3", "--trim-polyA"]
OPS_PER_CELL = 10            # algorithmic integer ops per DP cell


def takara_program():
    from cutseq_b200 import program, run
    from cutseq_b200.common import BarcodeConfig

    args = run.build_parser().parse_args(ARGV + ["r1.fq", "r2.fq"])
    scheme = run.resolve_scheme(args)
    return program.compile_paired(BarcodeConfig(scheme), run.settings_from_args(args))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        for l in rows:
            parts = [p.strip() for p in l.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def h2d_bytes(batch):
    total = 0
    for m in range(batch.n_mates):
        mi = batch.mate[m]
        total += 2 * mi.seq_bytes + 8 * batch.n_reads + mi.name_bytes + 4 * (batch.n_reads + 1)
    return int(total)


def cpu_leg(prog, steps, warmup, sample_pairs, first_index=0):
    """The oracle port on all host cores over bounded samples of the same workload."""
    from cutseq_b200 import native
    from oracle import oracle

    threads = oracle.lib().orc_max_threads()
    batch = native.synth_batch(2, sample_pairs, first_index=first_index, buffer=15)
    for _ in range(max(0, warmup)):
        oracle.run_batch(prog, batch, n_threads=threads, want_matches=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.run_batch(prog, batch, n_threads=threads, want_matches=False)
    dt = time.perf_counter() - t0
    return steps * sample_pairs / dt, threads, dt / steps


def files_leg(prog, pairs):
    """FASTQ files on disk -> csq_run_files -> trimmed FASTQ files, plain and .gz; host stages reported separately."""
    import shutil
    import tempfile

    from cutseq_b200 import native
    from scripts import bench_files

    tmp = tempfile.mkdtemp()
    out = {}
    try:
        batch = native.synth_batch(2, pairs, first_index=0, buffer=14)
        for variant in ("plain", "gz"):
            ext = ".fq.gz" if variant == "gz" else ".fq"
            ins = [os.path.join(tmp, f"in_R{m}{ext}") for m in (1, 2)]
            bench_files.write_fastq_from_batch(batch, ins, variant == "gz")
            outs = {"trimmed": [os.path.join(tmp, f"out_trimmed_R{m}{ext}") for m in (1, 2)],
                    "short": [os.path.join(tmp, f"out_short_R{m}{ext}") for m in (1, 2)]}
            best = None
            for rep in range(2):
                for q in outs["trimmed"] + outs["short"]:  # a fresh run writes new files
                    if os.path.exists(q):
                        os.remove(q)
                t0 = time.time()
                counters, timing = native.run_files(prog, ins, outs, gpus=1, threads=os.cpu_count() or 4)
                wall = time.time() - t0
                if best is None or wall < best[0]:
                    best = (wall, timing)
            wall, timing = best
            out[variant] = {"pairs_per_s": pairs / wall, "wall_s": wall, "read_inflate_s": timing.read_inflate,
                            "gpu_h2d_kernels_d2h_s": timing.h2d_kernels_d2h, "gpu_kernels_s": timing.kernels,
                            "deflate_write_s": timing.write_deflate, "input_bytes": sum(os.path.getsize(p) for p in ins)}
            for p in ins + outs["trimmed"] + outs["short"]:
                if os.path.exists(p):
                    os.remove(p)
        out["pairs"] = pairs
        out["host_threads"] = os.cpu_count()
        out["note"] = ("whole run incl. plan set-up and pinned allocations (~0.5 s fixed); stage seconds are busy times of "
                       "overlapping stages; .gz = one gzip member per input file (level 1), own inflate, zlib level-1 members out")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    prog = takara_program()
    sample = 200_000
    # the real `cutseq -t <cores>` would be the preferred CPU arm, but it needs the cutadapt package
    # (and the reference tree, which does not travel to the GPU box); probed here so that the line says which it is
    try:
        import cutadapt  # noqa: F401
        have_cutadapt = True
    except Exception:
        have_cutadapt = False
    value, threads, step_s = cpu_leg(prog, args.steps, min(args.warmup, 1), sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"config2: synthetic 2x150 read pairs, cutseq {' '.join(ARGV)}; step = {sample} pairs (bounded sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x {sample} pairs of the config-2 generator, oracle/cutseq_oracle.c (restated "
                                   "cutadapt chain; the reference needs the absent cutadapt package), gzip/parse excluded",
                         "cutadapt_importable": have_cutadapt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-pairs", type=int, default=BATCH_PAIRS)
    ap.add_argument("--batches", type=int, default=N_BATCHES)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the GPU's NUMA node (A/B runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the whole-file leg (FASTQ files on disk -> csq_run_files -> files)")
    ap.add_argument("--file-pairs", type=int, default=2_000_000, help="pairs in the whole-file leg")
    ap.add_argument("--no-prefilter", action="store_true", help="exact DP on every read (CSQ_PLAN_NO_PREFILTER)")
    ap.add_argument("--emit", default="stage", choices=["stage", "g16", "g32", "g8", "rec"], help="emit kernel variant (A/B runs); stage (k_emit_stage, through shared memory) is the product default")
    ap.add_argument("--homo", default="split", choices=["split", "one-lane", "v1"], help="poly-A / poly-T exact DP: column split over two lanes (default), whole column per thread with 2-4 columns side by side, or one column at a time (A/B runs)")
    ap.add_argument("--no-exact-stop", action="store_true", help="exact DP walks on after an error-free full match (CSQ_PLAN_NO_EXACT_STOP, A/B runs)")
    ap.add_argument("--parse", default="v1", choices=["onepass", "v1"], help="text-batch parse (A/B runs); v1 (four kernels) is the product default, onepass = single look-back kernel")
    ap.add_argument("--one-stream", action="store_true", help="mate chains on one stream (CSQ_PLAN_ONE_STREAM), for A/B runs")
    ap.add_argument("--input", default="text", choices=["text", "soa"], help="batch form handed to the library")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: anything libraries print there (e.g. NCCL's version banner)
    # is sent to stderr by pointing fd 1 at fd 2 until the result is written to the saved descriptor.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)

    import torch

    from cutseq_b200 import _abi as A
    from cutseq_b200 import build, native

    from cutseq_b200 import dist as csq_dist

    group = csq_dist.Group()  # NCCL process group when launched by torchrun with WORLD_SIZE > 1
    rank, local_rank, world = group.rank, group.local_rank, group.world
    dist = group.dist
    torch.cuda.set_device(local_rank)
    build.build()
    native.lib()  # fails loudly when the CUDA library is missing
    # one process per GPU: this rank's threads and pinned buffers stay on the GPU's NUMA node (-1: nothing to bind to)
    numa_node = -1 if args.no_numa else native.bind_host_to_device(local_rank)

    def barrier():
        torch.cuda.synchronize()
        group.barrier()
        torch.cuda.synchronize()

    max_over_ranks = group.max

    prog = takara_program()
    P, B = args.batch_pairs, args.batches
    plan = native.Plan(prog, local_rank, (A.PLAN_NO_PREFILTER if args.no_prefilter else 0) | {"stage": 0, "g16": A.PLAN_EMIT_G16, "g32": A.PLAN_EMIT_G32, "g8": A.PLAN_EMIT_G8, "rec": A.PLAN_EMIT_REC}[args.emit] | (A.PLAN_ONE_STREAM if args.one_stream else 0)
                       | (A.PLAN_PARSE_ONEPASS if args.parse == "onepass" else 0) | {"split": 0, "one-lane": A.PLAN_HOMO_ONE_LANE, "v1": A.PLAN_HOMO_V1}[args.homo] | (A.PLAN_NO_EXACT_STOP if args.no_exact_stop else 0))
    # this rank's contiguous index range of the workload: [rank*B*P, (rank+1)*B*P)
    # Host copies: batches 0 and 1 stay in pinned memory for the end-to-end leg; later batches reuse one
    # staging buffer (csq_upload is synchronous), so a rank pins three batches, not B.
    batches = []
    text_mode = args.input == "text"
    for b in range(B):
        batch = native.synth_batch(2, P, first_index=(rank * B + b) * P, buffer=min(b, 2))
        if text_mode:  # the FASTQ bytes of the batch, as a file would hold them, in pinned host memory
            texts = []
            for m in range(2):
                mi = batch.mate[m]
                buf = torch.empty(int(mi.name_bytes) + 2 * int(mi.seq_bytes) + 6 * P + 64, dtype=torch.uint8, pin_memory=True)
                n_bytes = native.format_fastq(batch, m, out=buf).numel()
                texts.append(buf[:n_bytes])
            tb = native.TextBatch(texts, P, first_record=(rank * B + b) * P)
            plan.upload_text(b, tb.c)
            if b < 2:
                batches.append(tb)
        else:
            plan.upload(b, batch)
            if b < 2:
                batches.append(batch)
    slots = list(range(B))

    def in_bytes(bt):
        return int(sum(bt.c.mate[m].bytes for m in range(2))) if text_mode else h2d_bytes(bt)

    def submit(slot, bt, out):
        if text_mode:
            plan.submit_text(slot, bt.c, out)
        else:
            plan.submit(slot, bt, out)

    # ---- kernel throughput, inputs resident in HBM ----
    if args.warmup > 0:
        plan.run_steps(slots, args.warmup)
    c0 = plan.stats()
    l0 = plan.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.time()
    ms = plan.run_steps(slots, args.steps)  # untimed sizing pass inside, then K steps between two CUDA events
    t1 = time.time()
    barrier()
    launches = plan.launch_count() - l0
    c1 = plan.stats()
    job_counters = group.sum_counters(c1)  # the one reduction of trim statistics (NCCL all-reduce when N > 1)
    ktimes = plan.kernel_times(slots[0])
    ms = max_over_ranks(ms)
    value = world * args.steps * P / (ms * 1e-3)
    # launches of the sizing pass are outside the CUDA-event bracket: count the timed ones only
    # csq_run_steps = B sizing steps + K timed steps + 1 per-kernel profiling step
    per_step_launches = launches // (args.steps + B + 1)
    timed_launches = per_step_launches * args.steps

    # nominal DP cells (device counters) over the timed + sizing steps -> cells per step
    def cells(c):
        return sum(sum(c.dp_cells[m]) for m in range(2))

    cells_per_step = (cells(c1) - cells(c0)) / (args.steps + B + 1)
    gcups_whole_chain = cells_per_step / (ms / args.steps * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H timed) ----
    e2e = None
    clocks = sampler.stop(t0, t1)
    if not args.no_e2e:
        cap = int(in_bytes(batches[0]) // 2 + 64 * P + 4096)
        outs, keep = [], []
        n_e2e = 3 if B + 2 < A.CSQ_N_SLOTS else 2  # batches in flight: keeps the H2D engine busy while a D2H drains
        for s in range(n_e2e):
            out = A.csq_batch_out()
            for d in range(A.CSQ_N_DEST):
                for m in range(2):
                    size = cap if d == 0 else cap // 4
                    buf = torch.empty(size, dtype=torch.uint8, pin_memory=True)
                    keep.append(buf)
                    out.text[d][m].data = buf.data_ptr()
                    out.text[d][m].capacity = size
            outs.append(out)
        e2e_slots = [B + i for i in range(n_e2e)] if B + n_e2e <= A.CSQ_N_SLOTS else list(range(n_e2e))

        def e2e_steps(k):
            """k steps, up to n_e2e batches in flight (submit of step i+n-1 is issued before the wait of step i)."""
            d2h = 0
            submitted = 0
            for i in range(k):
                while submitted < k and submitted < i + n_e2e:
                    submit(e2e_slots[submitted % n_e2e], batches[submitted % len(batches)], outs[submitted % n_e2e])
                    submitted += 1
                plan.wait(e2e_slots[i % n_e2e])
                o = outs[i % n_e2e]
                d2h += sum(o.text[d][m].bytes for d in range(A.CSQ_N_DEST) for m in range(2))
            return d2h

        e2e_steps(max(2, min(args.warmup, 3)))
        barrier()
        w0 = time.perf_counter()
        d2h = e2e_steps(args.steps)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        barrier()
        e2e_s = max_over_ranks(w1 - w0)
        e2e = {"value": world * args.steps * P / e2e_s, "unit": UNIT, "h2d_bytes_per_step": in_bytes(batches[0]),
               "d2h_bytes_per_step": int(d2h // args.steps), "ms_per_step": e2e_s / args.steps * 1e3,
               "timing": f"wall clock between device-synchronised points, {n_e2e} batches in flight through csq_submit/csq_wait, max over ranks"}

    native.unbind_host()  # the host legs below (files, CPU baseline) use every core of the box
    if rank != 0:
        plan.close()
        group.close()
        return 0

    # ---- roofline of the dominant kernel (per-kernel CUDA events of the last timed step) ----
    alu_peak, mixed_peak = native.int_peak(local_rank)
    peak = max(alu_peak, mixed_peak)
    step_ms = ms / args.steps
    dom_name, dom_ms = max(ktimes, key=lambda kv: kv[1]) if ktimes else ("n/a", 0.0)
    # nominal cells of each ALIGN launch, in launch order (mate 1 ops then mate 2 ops)
    align_cells = []
    for m, ops in enumerate((prog.ops_r1, prog.ops_r2)):
        for t, op in enumerate(ops):
            if op.kind == A.OP_ALIGN:
                align_cells.append((c1.dp_cells[m][t] - c0.dp_cells[m][t]) / (args.steps + B + 1))
    # one entry per ALIGN op: its k_prefilter launch (if any) plus its k_align launch
    align_ops = []
    for name, kms in ktimes:
        if name.startswith("k_prefilter"):
            align_ops.append([name.replace("k_prefilter", "k_prefilter+k_align"), kms, True])
        elif name.startswith("k_align"):
            if align_ops and align_ops[-1][2]:
                align_ops[-1][1] += kms
                align_ops[-1][2] = False
            else:
                align_ops.append([name, kms, False])
    per_kernel = []
    for (name, kms, _), cl in zip(align_ops, align_cells):
        per_kernel.append({"kernel": name, "ms": kms, "gcups": cl / (kms * 1e-3) / 1e9 if kms > 0 else None})
    hbm_peak, hbm_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = json.load(f)["hbm_gbs"], "measured"
    except Exception:
        pass
    d2h_per_step = e2e["d2h_bytes_per_step"] if e2e else int(0.72 * in_bytes(batches[0]))
    emit_bytes = in_bytes(batches[0]) + d2h_per_step  # input read once + FASTQ text written once (~1.3 KB / pair)
    emit_ms = sum(kms for name, kms in ktimes if name == "k_emit")
    roofline_dp = None
    dom_dp = max(per_kernel, key=lambda r: r["ms"]) if per_kernel else None
    if dom_dp:
        achieved = dom_dp["gcups"] * OPS_PER_CELL  # Gop/s (algorithmic integer lane-ops)
        roofline_dp = {
            "bound": "int_issue", "kernel": dom_dp["kernel"], "achieved": achieved, "peak": peak / 1e9, "unit": "Gop/s",
            "frac": achieved / (peak / 1e9), "traffic": None,
            "how": f"nominal DP cells of the op (m x columns, counted on the device) x {OPS_PER_CELL} ops/cell / CUDA-event duration of its "
                   f"launches; peak = csq_int_peak measured live (ALU-only {alu_peak / 1e12:.2f}, ALU+FMA mix {mixed_peak / 1e12:.2f} "
                   "T lane-op/s). With the bit-parallel prefilter most nominal cells are never visited, so this can exceed 1.",
        }
    # the same op by EXECUTED integer lane-ops: instructions per read of each kernel from the committed ncu capture of
    # this workload (deterministic for the generator), times the reads of a launch, over the live CUDA-event time
    roofline_dp_executed = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            ops_tab = json.load(f)["dp_lane_ops_per_read"]
        if dom_dp and not args.no_prefilter:
            kind = dom_dp["kernel"][dom_dp["kernel"].index("("):]
            lane_ops = ops_tab["k_prefilter" + kind] + ops_tab["k_align" + kind]
            achieved = lane_ops * P / (dom_dp["ms"] * 1e-3) / 1e9
            roofline_dp_executed = {"bound": "int_issue", "kernel": dom_dp["kernel"], "achieved": achieved, "peak": peak / 1e9, "unit": "Gop/s",
                                    "frac": achieved / (peak / 1e9), "traffic": None,
                                    "how": "executed integer lane-ops per read (smsp__inst_executed x 32 of the committed ncu capture, "
                                           + ops_tab["source"] + ") x reads per launch / CUDA-event duration of the op's launches; peak = csq_int_peak measured live"}
    except Exception:
        pass
    roofline_hbm = {"bound": "hbm", "kernel": {"stage": "k_emit_stage", "rec": "k_emit_rec"}.get(args.emit, "k_emit<%s>" % args.emit[1:]), "achieved": (emit_bytes / (emit_ms * 1e-3) / 1e9) if emit_ms else None,
                    "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src, "traffic": None,
                    "how": "algorithmic bytes per launch (input records read once + FASTQ text written once) / CUDA-event duration"}
    if roofline_hbm["achieved"]:
        roofline_hbm["frac"] = roofline_hbm["achieved"] / hbm_peak
    try:  # DRAM bytes of one launch from the committed ncu capture (per pair, scaled to this batch size)
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tr = json.load(f)[{"stage": "k_emit", "g16": "k_emit_g16"}[args.emit]]
        roofline_hbm["traffic"] = tr["dram_bytes_per_pair"] * P
        roofline_hbm["traffic_source"] = tr["source"]
    except Exception:
        pass
    # the dominant kernel of the step decides which of the two is THE roofline line
    roofline = roofline_hbm if (emit_ms and (not dom_dp or emit_ms >= dom_dp["ms"])) else roofline_dp

    # ---- whole files: read / inflate, GPU chain, deflate / write (N = 1, rank 0; reported separately) ----
    files = None
    if not args.no_files and world == 1:
        try:
            files = files_leg(prog, args.file_pairs)
        except Exception as exc:  # the headline numbers must not die with a full /tmp
            files = {"error": repr(exc)}

    cpu = None
    if not args.no_cpu and world == 1:
        v, threads, step_s = cpu_leg(prog, 3, 1, 100_000)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "3 x 100000 pairs of the same config-2 generator through oracle/cutseq_oracle.c (restated cutadapt chain), "
                         "gzip/parse excluded"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": {"workload": f"config2: synthetic 2x150 read pairs, cutseq {' '.join(ARGV)}; {B} resident batches x {P} pairs per GPU "
                               f"(= {B * P} pairs), step = one batch; consecutive steps use different batches "
                               f"({in_bytes(batches[0]) / 1e9:.2f} GB in each, far larger than the 126 MB L2, no flush needed); "
                               f"input form: {'raw FASTQ text, record index built on the device' if text_mode else 'host-parsed SoA'}",
                   "parallelism": f"dp{world} (contiguous index ranges per GPU, no collective on the data path)",
                   "prefilter": not args.no_prefilter, "emit": args.emit, "parse": args.parse, "homo_dp": args.homo, "exact_stop": not args.no_exact_stop, "numa_node": numa_node},
        "gcups": gcups_whole_chain, "cells_per_pair": cells_per_step / P,
        "roofline": roofline, "roofline_dp": roofline_dp, "roofline_dp_executed": roofline_dp_executed, "roofline_hbm": roofline_hbm, "kernels": [{"kernel": n, "ms": t} for n, t in ktimes],
        "dp_kernels": per_kernel, "cpu_baseline": cpu, "e2e": e2e, "files": files, "gpu_launches": int(timed_launches), "clocks": clocks,
        "job_counters": {"pairs": int(job_counters.n), "written": int(job_counters.written), "too_short": int(job_counters.too_short)},
    }
    sys.stdout.flush()
    os.write(result_fd, (json.dumps(line) + "\n").encode())
    plan.close()
    group.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
