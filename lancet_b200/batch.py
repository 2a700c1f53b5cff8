"""Window batches: the host-side container for the C ABI's ``lb2_batch`` (include/lancet_b200.h).

A batch is what ``Graph_t::addAlignment`` would have accumulated for many windows
(reference src/Graph.cc:487-501) plus each window's ``Ref_t`` (reference src/Lancet.cc:283-300),
laid out as flat numpy arrays: one read pool, per-window index lists (tumor reads first, then
normal reads, each in BAM order -- reference src/Microassembler.cc:833-834).

The same arrays are the on-disk ``.lb2b`` format consumed by ``oracle/_ref/ref_windows``.
"""
from __future__ import annotations

import ctypes
import dataclasses
import struct

import numpy as np

READ_NORMAL = 0x01
READ_REVERSE = 0x02
READ_MATE_SHIFT = 2
READ_UNMAPPED = 0x10


@dataclasses.dataclass
class Batch:
    ref_off: np.ndarray    # u32 [W+1]
    ref_start: np.ndarray  # i32 [W]
    chr_id: np.ndarray     # u32 [W]
    wr_off: np.ndarray     # u32 [W+1]
    wr_idx: np.ndarray     # u32 [n_wr]
    base_off: np.ndarray   # u64 [R+1]
    flags: np.ndarray      # u8  [R]
    name_rank: np.ndarray  # u32 [R]
    ref_seq: np.ndarray    # u8  bytes
    seq: np.ndarray        # u8  bytes
    qual: np.ndarray       # u8  bytes

    def __post_init__(self):
        c = np.ascontiguousarray
        self.ref_off = c(self.ref_off, dtype=np.uint32)
        self.ref_start = c(self.ref_start, dtype=np.int32)
        self.chr_id = c(self.chr_id, dtype=np.uint32)
        self.wr_off = c(self.wr_off, dtype=np.uint32)
        self.wr_idx = c(self.wr_idx, dtype=np.uint32)
        self.base_off = c(self.base_off, dtype=np.uint64)
        self.flags = c(self.flags, dtype=np.uint8)
        self.name_rank = c(self.name_rank, dtype=np.uint32)
        self.ref_seq = c(self.ref_seq, dtype=np.uint8)
        self.seq = c(self.seq, dtype=np.uint8)
        self.qual = c(self.qual, dtype=np.uint8)
        assert len(self.ref_off) == self.n_windows + 1 and len(self.wr_off) == self.n_windows + 1
        assert len(self.base_off) == self.n_reads + 1
        assert int(self.wr_off[-1]) == len(self.wr_idx)
        assert int(self.ref_off[-1]) == len(self.ref_seq)
        assert int(self.base_off[-1]) == len(self.seq) == len(self.qual)

    @property
    def n_windows(self) -> int:
        return len(self.ref_start)

    @property
    def n_reads(self) -> int:
        return len(self.flags)

    # ---- algorithmic bytes (B_win) -------------------------------------------
    def algorithmic_bytes(self, n_variants: int = 0) -> int:
        """sum over windows of sum_reads(ceil(len/4)+len) + L_ref + 32, plus 64 per variant."""
        lens = (self.base_off[1:] - self.base_off[:-1]).astype(np.int64)
        per_read = (lens + 3) // 4 + lens
        tot = int(per_read[self.wr_idx].sum())
        tot += int(len(self.ref_seq)) + 32 * self.n_windows + 64 * n_variants
        return tot

    def window_reads(self, w: int):
        idx = self.wr_idx[self.wr_off[w]:self.wr_off[w + 1]]
        out = []
        for r in idx:
            o0, o1 = int(self.base_off[r]), int(self.base_off[r + 1])
            out.append((bytes(self.seq[o0:o1]).decode(), bytes(self.qual[o0:o1]).decode(), int(self.flags[r]),
                        int(self.name_rank[r])))
        return out

    def window_ref(self, w: int) -> str:
        return bytes(self.ref_seq[self.ref_off[w]:self.ref_off[w + 1]]).decode()

    def subset(self, windows) -> "Batch":
        """A new batch holding only ``windows`` (pool compacted to the reads they use)."""
        windows = np.asarray(windows, dtype=np.int64)
        wr_lists = [self.wr_idx[self.wr_off[w]:self.wr_off[w + 1]] for w in windows]
        used = np.unique(np.concatenate(wr_lists)) if wr_lists and sum(map(len, wr_lists)) else np.zeros(0, np.uint32)
        remap = np.full(self.n_reads, -1, dtype=np.int64)
        remap[used] = np.arange(len(used))
        lens = (self.base_off[1:] - self.base_off[:-1]).astype(np.int64)[used]
        base_off = np.zeros(len(used) + 1, dtype=np.uint64)
        base_off[1:] = np.cumsum(lens)
        gather = np.concatenate([np.arange(int(self.base_off[r]), int(self.base_off[r + 1])) for r in used]) \
            if len(used) else np.zeros(0, np.int64)
        ref_parts = [self.ref_seq[self.ref_off[w]:self.ref_off[w + 1]] for w in windows]
        ref_off = np.zeros(len(windows) + 1, dtype=np.uint32)
        ref_off[1:] = np.cumsum([len(p) for p in ref_parts])
        wr_off = np.zeros(len(windows) + 1, dtype=np.uint32)
        wr_off[1:] = np.cumsum([len(x) for x in wr_lists])
        return Batch(
            ref_off=ref_off, ref_start=self.ref_start[windows], chr_id=self.chr_id[windows],
            wr_off=wr_off,
            wr_idx=remap[np.concatenate(wr_lists)] if len(used) else np.zeros(0, np.uint32),
            base_off=base_off, flags=self.flags[used], name_rank=self.name_rank[used],
            ref_seq=np.concatenate(ref_parts) if ref_parts else np.zeros(0, np.uint8),
            seq=self.seq[gather], qual=self.qual[gather])

    # ---- .lb2b file -------------------------------------------------------------------------
    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(b"LB2B")
            f.write(struct.pack("<IIIIQQ", 2, self.n_windows, self.n_reads, len(self.wr_idx),
                                len(self.ref_seq), len(self.seq)))
            for a in (self.ref_off, self.ref_start, self.chr_id, self.wr_off, self.wr_idx, self.base_off,
                      self.flags, self.name_rank, self.ref_seq, self.seq, self.qual):
                a.tofile(f)

    @staticmethod
    def load(path: str) -> "Batch":
        with open(path, "rb") as f:
            assert f.read(4) == b"LB2B"
            ver, W, R, nwr, nref, nbase = struct.unpack("<IIIIQQ", f.read(32))
            assert ver == 2

            def rd(dt, n):
                return np.fromfile(f, dtype=dt, count=n)
            return Batch(ref_off=rd(np.uint32, W + 1), ref_start=rd(np.int32, W), chr_id=rd(np.uint32, W),
                         wr_off=rd(np.uint32, W + 1), wr_idx=rd(np.uint32, nwr), base_off=rd(np.uint64, R + 1),
                         flags=rd(np.uint8, R), name_rank=rd(np.uint32, R), ref_seq=rd(np.uint8, nref),
                         seq=rd(np.uint8, nbase), qual=rd(np.uint8, nbase))

    # ---- ctypes view for the C ABI ----------------------------------------------------------
    def as_struct(self) -> "LB2Batch":
        s = LB2Batch()
        s.n_windows = self.n_windows
        s.n_reads = self.n_reads
        s.n_wr = len(self.wr_idx)
        s.n_ref_bytes = len(self.ref_seq)
        s.n_base_bytes = len(self.seq)
        for name in ("ref_off", "ref_start", "chr_id", "wr_off", "wr_idx", "base_off", "flags", "name_rank",
                     "ref_seq", "seq", "qual"):
            setattr(s, name, getattr(self, name).ctypes.data)
        s._keep = self  # keep the arrays alive
        return s


class LB2Batch(ctypes.Structure):
    _fields_ = [
        ("n_windows", ctypes.c_uint32), ("n_reads", ctypes.c_uint32), ("n_wr", ctypes.c_uint32),
        ("n_ref_bytes", ctypes.c_uint64), ("n_base_bytes", ctypes.c_uint64),
        ("ref_off", ctypes.c_void_p), ("ref_start", ctypes.c_void_p), ("chr_id", ctypes.c_void_p),
        ("wr_off", ctypes.c_void_p), ("wr_idx", ctypes.c_void_p), ("base_off", ctypes.c_void_p),
        ("flags", ctypes.c_void_p), ("name_rank", ctypes.c_void_p), ("ref_seq", ctypes.c_void_p),
        ("seq", ctypes.c_void_p), ("qual", ctypes.c_void_p),
    ]
