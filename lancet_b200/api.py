"""ctypes binding of the C ABI in include/lancet_b200.h (lancet_b200/_lb2.so, built by __graft_entry__.build()).

The package has no CPU path: importing this module without the CUDA extension raises, and creating a
Context without an sm_100 GPU raises.  The host-side mirror of the reference interface is
``Context.process(batch)`` = "addAlignment* ; Microassembler::processGraph" over a batch of windows
(reference src/Microassembler.cc:73-249), returning the Variant_t constructor tuples the reference
hands to VariantDB_t::addVar (src/Graph.cc:1184-1188).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .batch import Batch, LB2Batch

# LB2_SO selects another build of the same library (tools/phase_profile.py: the -DLB2_PROFILE variant _lb2_prof.so)
_SO = os.environ.get("LB2_SO") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lb2.so")
if not os.path.exists(_SO):
    raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(there is no CPU fallback for the lancet_b200 hot path)")
_lib = ctypes.CDLL(_SO)

VARIANT_DTYPE = np.dtype([
    ("window", "<u4"), ("pos", "<i4"), ("str_off", "<u4"), ("ref_len", "<u2"), ("alt_len", "<u2"),
    ("motif_len", "<u2"), ("str_len", "<u2"), ("rcn_fwd", "<u2"), ("rcn_rev", "<u2"), ("rct_fwd", "<u2"),
    ("rct_rev", "<u2"), ("acn_fwd", "<u2"), ("acn_rev", "<u2"), ("act_fwd", "<u2"), ("act_rev", "<u2"),
    ("code", "u1"), ("prev_bp_ref", "u1"), ("prev_bp_alt", "u1"), ("kmer", "u1")])
WINDOW_DTYPE = np.dtype([("status", "u1"), ("final_k", "u1"), ("n_k_tried", "<u2"), ("n_variants", "<u4"),
                         ("n_nodes", "<u4"), ("detail", "<u4")])
assert VARIANT_DTYPE.itemsize == 40 and WINDOW_DTYPE.itemsize == 16

WIN_OK, WIN_SKIP_REPEAT, WIN_NO_READS, WIN_OVERFLOW, WIN_UNSUPPORTED = 0, 1, 2, 3, 4


class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "min_k", "max_k", "min_qual_trim", "min_qual_call", "cov_threshold", "low_cov_threshold", "max_tip_len",
        "dfs_limit", "max_indel_len", "max_mismatch", "max_unit_len", "min_report_units", "min_report_len",
        "dist_from_str")] + [("min_cov_ratio", ctypes.c_double)]

    @staticmethod
    def default(**over) -> "Params":
        p = Params()
        _lib.lb2_default_params(ctypes.byref(p))
        for k, v in over.items():
            setattr(p, k, v)
        return p


class _Result(ctypes.Structure):
    _fields_ = [("n_windows", ctypes.c_uint32), ("n_variants", ctypes.c_uint32), ("windows", ctypes.c_void_p),
                ("variants", ctypes.c_void_p), ("strings", ctypes.c_void_p), ("n_string_bytes", ctypes.c_uint64),
                ("kernel_ms", ctypes.c_float)]


_lib.lb2_create.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(Params), ctypes.c_int]
_lib.lb2_destroy.argtypes = [ctypes.c_void_p]
_lib.lb2_strerror.argtypes = [ctypes.c_void_p, ctypes.c_int]
_lib.lb2_strerror.restype = ctypes.c_char_p
for _n in ("lb2_process",):
    getattr(_lib, _n).argtypes = [ctypes.c_void_p, ctypes.POINTER(LB2Batch), ctypes.POINTER(_Result)]
_lib.lb2_upload.argtypes = [ctypes.c_void_p, ctypes.POINTER(LB2Batch)]
_lib.lb2_run.argtypes = [ctypes.c_void_p]
_lib.lb2_wait.argtypes = [ctypes.c_void_p]
_lib.lb2_download.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Result)]
_lib.lb2_kernel_launches.argtypes = [ctypes.c_void_p]
_lib.lb2_kernel_launches.restype = ctypes.c_uint64
for _n in ("lb2_last_h2d_bytes", "lb2_last_d2h_bytes"):
    getattr(_lib, _n).argtypes = [ctypes.c_void_p]
    getattr(_lib, _n).restype = ctypes.c_uint64
for _n in ("lb2_resident_ctas", "lb2_smem_per_cta"):
    getattr(_lib, _n).argtypes = [ctypes.c_void_p]
    getattr(_lib, _n).restype = ctypes.c_uint32
_lib.lb2_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
_lib.lb2_last_kernel_ms.argtypes = [ctypes.c_void_p]
_lib.lb2_last_kernel_ms.restype = ctypes.c_float
_lib.lb2_last_kernel_ms_of.argtypes = [ctypes.c_void_p, ctypes.c_int]
_lib.lb2_last_kernel_ms_of.restype = ctypes.c_float
_lib.lb2_kernel_version.restype = ctypes.c_char_p
KERNEL_VERSION = _lib.lb2_kernel_version().decode()


_lib.lb2_comm_unique_id.argtypes = [ctypes.c_char_p]
_lib.lb2_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
_lib.lb2_comm_gather.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint64,
                                 ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.c_int, ctypes.POINTER(_Result)]
COMM_ID_BYTES = 128


def comm_unique_id() -> bytes:
    """NCCL id for Context.comm_init (created on one rank, handed to the others by any means)"""
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    rc = _lib.lb2_comm_unique_id(buf)
    if rc != 0:
        raise RuntimeError(f"lb2_comm_unique_id failed ({rc}): no GPU or no NCCL")
    return buf.raw


class Result:
    """Host copy of an lb2_result."""

    def __init__(self, r: _Result):
        self.windows = np.ctypeslib.as_array(ctypes.cast(r.windows, ctypes.POINTER(ctypes.c_uint8)),
                                             (r.n_windows * 16,)).view(WINDOW_DTYPE).copy() if r.n_windows else np.zeros(0, WINDOW_DTYPE)
        self.variants = np.ctypeslib.as_array(ctypes.cast(r.variants, ctypes.POINTER(ctypes.c_uint8)),
                                              (r.n_variants * 40,)).view(VARIANT_DTYPE).copy() if r.n_variants else np.zeros(0, VARIANT_DTYPE)
        self.strings = ctypes.string_at(r.strings, r.n_string_bytes) if r.n_string_bytes else b""
        self.kernel_ms = float(r.kernel_ms)

    def records(self):
        """Variant_t constructor normalisation (reference src/Variant.hh:133-153) applied to the raw tuples:
        (window, pos, type, len, ref, alt, kmer, str, (8 counts), prev_bp_ref, prev_bp_alt)."""
        out = []
        s = self.strings
        for v in self.variants:
            o = int(v["str_off"]); rl, al, ml = int(v["ref_len"]), int(v["alt_len"]), int(v["motif_len"])
            ref = s[o:o + rl].decode(); alt = s[o + rl:o + rl + al].decode(); motif = s[o + rl + al:o + rl + al + ml].decode()
            code = chr(v["code"]); pos = int(v["pos"]); pba = chr(v["prev_bp_alt"])
            if code == "^":
                typ, ref, ln = "I", "", len(alt)
            elif code == "v":
                typ, alt, ln = "D", "", len(ref)
            elif code == "x":
                typ, pos, ln = "S", pos + 1, 1
            else:
                typ = "C"; ref = ref.replace("-", ""); alt = alt.replace("-", "")
                ln = len(alt) if len(ref) == len(alt) else abs(len(ref) - len(alt))
            if typ != "S":
                ref, alt = pba + ref, pba + alt
            st = f"{int(v['str_len'])}{motif}" if int(v["str_len"]) else "."
            counts = tuple(int(v[k]) for k in ("rcn_fwd", "rcn_rev", "rct_fwd", "rct_rev", "acn_fwd", "acn_rev", "act_fwd", "act_rev"))
            out.append((int(v["window"]), pos, typ, ln, ref, alt, int(v["kmer"]), st, counts, chr(v["prev_bp_ref"]), pba))
        return out


class Context:
    def __init__(self, params: Params | None = None, device: int = 0):
        self._h = ctypes.c_void_p()
        self.params = params or Params.default()
        rc = _lib.lb2_create(ctypes.byref(self._h), ctypes.byref(self.params), device)
        if rc != 0:
            raise RuntimeError(f"lb2_create failed ({rc}): {_lib.lb2_strerror(None, rc).decode()}")

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError(f"lancet_b200 call failed ({rc}): {_lib.lb2_strerror(self._h, rc).decode()}")

    def process(self, batch: Batch) -> Result:
        s = batch.as_struct(); r = _Result()
        self._ck(_lib.lb2_process(self._h, ctypes.byref(s), ctypes.byref(r)))
        return Result(r)

    def upload(self, batch: Batch) -> None:
        self._struct = batch.as_struct()
        self._ck(_lib.lb2_upload(self._h, ctypes.byref(self._struct)))

    def run(self) -> None:
        self._ck(_lib.lb2_run(self._h))

    def wait(self) -> None:
        self._ck(_lib.lb2_wait(self._h))

    def download(self) -> Result:
        r = _Result()
        self._ck(_lib.lb2_download(self._h, ctypes.byref(r)))
        return Result(r)

    # ---- several GPUs, one process each: the record gather on one rank over NCCL (include/lancet_b200.h) ----
    def comm_init(self, comm_id: bytes, rank: int, world: int) -> None:
        self._ck(_lib.lb2_comm_init(self._h, comm_id, rank, world))
        self.rank, self.world = rank, world

    def comm_gather(self, variants: np.ndarray, strings: bytes, window_offset: int = 0, stats=(), root: int = 0):
        """every rank passes its records (local window indices + window_offset = global); on `root` returns
        (variants, strings, summed stats) of all ranks in rank order, elsewhere (None, None, None)"""
        v = np.ascontiguousarray(variants.copy()); v["window"] += window_offset
        st = (ctypes.c_uint64 * max(len(stats), 1))(*[int(x) for x in stats])
        r = _Result()
        self._ck(_lib.lb2_comm_gather(self._h, v.ctypes.data if len(v) else None, len(v), strings, len(strings), st, len(stats), root, ctypes.byref(r)))
        if self.rank != root:
            return None, None, None
        gv = np.ctypeslib.as_array(ctypes.cast(r.variants, ctypes.POINTER(ctypes.c_uint8)), (r.n_variants * 40,)).view(VARIANT_DTYPE).copy() if r.n_variants else np.zeros(0, VARIANT_DTYPE)
        gs = ctypes.string_at(r.strings, r.n_string_bytes) if r.n_string_bytes else b""
        return gv, gs, [int(st[i]) for i in range(len(stats))]

    kernel_launches = property(lambda self: int(_lib.lb2_kernel_launches(self._h)))
    last_h2d_bytes = property(lambda self: int(_lib.lb2_last_h2d_bytes(self._h)))
    last_d2h_bytes = property(lambda self: int(_lib.lb2_last_d2h_bytes(self._h)))
    resident_ctas = property(lambda self: int(_lib.lb2_resident_ctas(self._h)))
    smem_per_cta = property(lambda self: int(_lib.lb2_smem_per_cta(self._h)))
    last_kernel_ms = property(lambda self: float(_lib.lb2_last_kernel_ms(self._h)))

    def last_kernel_ms_of(self, which: str) -> float:
        """device time of one stage of the last run(): 'pack' (pool pre-pack), 'windows' (first pass), 'escalation', 'compaction'"""
        return float(_lib.lb2_last_kernel_ms_of(self._h, ("pack", "windows", "escalation", "compaction").index(which)))

    PHASES = ("stage", "prescan_maxk", "ref_repeat_scan", "kmer_walk", "compact_sort", "mate_replay", "lowq_deficits",
              "table_clear", "ref_coverage", "order_emulation", "lowcov_components", "component_sequential", "bfs_loadpath",
              "path_repeat_scan", "align", "column_scan_emit", "other", "anchors_cycle1", "compress_sweep", "compress_layout", "compress_cleandead", "compact_links", "compact_fold", "bfs_lane0")

    def phase_cycles(self, reset: bool = True) -> dict:
        """lane-0 cycles per pipeline phase, summed over windows since the last reset (debugging aid)."""
        buf = (ctypes.c_ulonglong * 24)()
        self._ck(_lib.lb2_phase_cycles(self._h, buf, 1 if reset else 0))
        return {n: int(buf[i]) for i, n in enumerate(self.PHASES)}

    def close(self):
        if self._h:
            _lib.lb2_destroy(self._h); self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
