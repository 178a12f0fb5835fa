#!/usr/bin/env python
"""End-to-end FILE benchmark (BASELINE.json config 5, single box): synthetic FASTQ files on disk ->
csq_run_files -> trimmed FASTQ files, with the host-side cost reported separately
(read+inflate / GPU section / deflate+write), plain and .gz variants, 1..N GPUs, output hashes.

    python scripts/bench_files.py --pairs 20000000 --gpus 1 --threads 16 --out gpurun_out/files.json

Also the fixture / hashing helpers of bench.py's files leg.
"""

import argparse
import hashlib
import json
import os
import shutil
import struct
import sys
import tempfile
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BLOCK_PAIRS = 4_000_000   # pairs generated at a time; larger fixtures repeat this block (trimming does not care)
BGZF_PIECE = 0xFF00       # uncompressed bytes per BGZF member, as bgzip


def bgzf_member(piece: bytes, level: int = 1) -> bytes:
    """One BGZF member (SAM/BAM specification 4.1): gzip header with the 'BC' extra field = member size - 1."""
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = c.compress(piece) + c.flush()
    size = 18 + len(body) + 8
    head = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", size - 1)
    return head + body + struct.pack("<II", zlib.crc32(piece) & 0xFFFFFFFF, len(piece) & 0xFFFFFFFF)


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_compress(text, threads: int) -> bytes:
    view = memoryview(text)
    pieces = [view[o:o + BGZF_PIECE] for o in range(0, len(view), BGZF_PIECE)]
    with ThreadPoolExecutor(max(1, threads)) as ex:
        members = list(ex.map(lambda p: bgzf_member(bytes(p)), pieces, chunksize=64))
    return b"".join(members)


def gzip_single_member(text, threads: int, level: int = 6, piece: int = 4 << 20) -> bytes:
    """ONE gzip member holding `text` - what `gzip` / `pigz` and the sequencers' software write - compressed by several
    threads the way pigz does it: every piece is deflated on its own and ends in a sync flush (an empty stored block,
    byte aligned, not final), only the last piece carries the final block; CRC-32 over the whole text."""
    view = memoryview(text)
    pieces = [view[o:o + piece] for o in range(0, len(view), piece)] or [view]

    def one(arg):
        i, p = arg
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        return c.compress(bytes(p)) + c.flush(zlib.Z_FINISH if i == len(pieces) - 1 else zlib.Z_SYNC_FLUSH)

    with ThreadPoolExecutor(max(1, threads)) as ex:
        bodies = list(ex.map(one, enumerate(pieces)))
    crc = 0
    for p in pieces:
        crc = zlib.crc32(p, crc)
    head = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff"
    return head + b"".join(bodies) + struct.pack("<II", crc & 0xFFFFFFFF, len(view) & 0xFFFFFFFF)


def write_fastq_from_soa(batch, paths):
    """Serialise a synthetic SoA batch to plain FASTQ files (csq_format_fastq of the product library is NOT used here:
    this also serves the CPU arm)."""
    import ctypes as C

    import numpy as np

    for m, path in enumerate(paths):
        mi = batch.mate[m]
        n = batch.n_reads
        noff = np.ctypeslib.as_array(C.cast(mi.name_off, C.POINTER(C.c_uint32)), (n + 1,))
        soff = np.ctypeslib.as_array(C.cast(mi.seq_off, C.POINTER(C.c_uint32)), (n,))
        slen = np.ctypeslib.as_array(C.cast(mi.seq_len, C.POINTER(C.c_uint32)), (n,))
        name = C.string_at(mi.name, mi.name_bytes)
        seq = C.string_at(mi.seq, mi.seq_bytes)
        qual = C.string_at(mi.qual, mi.seq_bytes)
        with open(path, "wb") as f:
            for i in range(n):
                f.write(b"@" + name[noff[i]:noff[i + 1]] + b"\n" + seq[soff[i]:soff[i] + slen[i]] + b"\n+\n" + qual[soff[i]:soff[i] + slen[i]] + b"\n")


def write_fixture(paths, pairs: int, compress, threads: int):
    """Config-2 FASTQ files of `pairs` pairs: blocks of up to 4 M generated pairs, the first block repeated (names
    repeat too; both mates stay in step).  compress = True / "bgzf": BGZF members of 0xFF00 bytes at level 1, then the
    BGZF EOF marker; "gzip": ordinary gzip members (level 6, one per generated block - a file of up to 4 M pairs is a
    single member, like the output of `gzip`)."""
    from cutseq_b200 import native

    block = min(pairs, BLOCK_PAIRS)
    reps, rest = divmod(pairs, block)
    handles = [open(p, "wb") for p in paths]
    try:
        for n, times in ((block, reps), (rest, 1 if rest else 0)):
            if not times:
                continue
            batch = native.synth_batch(2, n, first_index=0, buffer=14)
            for m, f in enumerate(handles):
                text = native.format_fastq(batch, m)
                data = gzip_single_member(text, threads) if compress == "gzip" else bgzf_compress(text, threads) if compress else memoryview(text)
                for _ in range(times):
                    f.write(data)
        if compress and compress != "gzip":
            for f in handles:
                f.write(BGZF_EOF)
    finally:
        for f in handles:
            f.close()
        native.lib().csq_synth_free()


def hash_outputs(paths, gz: bool):
    """sha256 of the (decompressed) bytes of every output file, one thread per file."""

    def one(path):
        h = hashlib.sha256()
        if not os.path.exists(path):
            return None
        with open(path, "rb") as f:
            if not gz:
                while True:
                    b = f.read(8 << 20)
                    if not b:
                        break
                    h.update(b)
            else:
                d = zlib.decompressobj(31)
                while True:
                    b = f.read(4 << 20)
                    if not b:
                        break
                    while b:
                        h.update(d.decompress(b))
                        if d.eof:  # next gzip member
                            b = d.unused_data
                            d = zlib.decompressobj(31)
                        else:
                            b = b""
        return h.hexdigest()

    with ThreadPoolExecutor(len(paths)) as ex:
        return dict(zip([os.path.basename(p).split("_", 1)[1] for p in paths], ex.map(one, paths)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=8_000_000)
    ap.add_argument("--pairs-gz", type=int, default=None)
    ap.add_argument("--pairs-gzip", type=int, default=4_000_000, help="pairs of the ordinary-gzip variant (single member in, .gz out)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--out", default=None)
    ap.add_argument("--variants", default="plain,gz", help="comma separated: plain, gz")
    args = ap.parse_args()
    import bench

    prog = bench.takara_program()
    v = args.variants.split(",")
    res = bench.files_leg(prog, args.pairs if "plain" in v else 0, (args.pairs_gz or args.pairs) if "gz" in v else 0, args.gpus, args.threads,
                          pairs_gzip=args.pairs_gzip if "gzip" in v else 0)
    print(json.dumps(res, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
