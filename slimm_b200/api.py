"""ctypes binding of the C ABI in include/slimm_gpu.h (the drop-in boundary of the hot path).

The names follow the reference's members (reference src/slimm.hpp:92-165).  There is no CPU
fallback: importing works anywhere (the library loads without a GPU), but creating a
:class:`SlimmGpu` raises when no CUDA device is usable or the native library was not built.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLIMM_GPU_LIB") or os.path.join(_HERE, "libslimm_gpu.so")   # SLIMM_GPU_LIB: another build of the same library (A/B runs)

KEEP_UNIQ_COV2 = 1
READ_RESULTS = 2
SKIP_BINS = 4
TIMING_NAMES = ["sort", "zero", "split", "coverage", "accumulate", "stats", "cutoff", "assign", "tail_host"]

EXPORTED_SYMBOLS = [
    "slimm_gpu_strerror", "slimm_gpu_last_error", "slimm_gpu_device_count", "slimm_gpu_create", "slimm_gpu_destroy",
    "slimm_gpu_reset", "slimm_gpu_set_stream", "slimm_gpu_host_alloc", "slimm_gpu_host_free", "slimm_gpu_push",
    "slimm_gpu_push_device", "slimm_gpu_sync_uploads", "slimm_gpu_coverage", "slimm_gpu_bins_device",
    "slimm_gpu_counters_device", "slimm_gpu_set_global_hits", "slimm_gpu_filter", "slimm_gpu_assign",
    "slimm_gpu_assign_device", "slimm_gpu_run", "slimm_gpu_get_summary", "slimm_gpu_get_ref_stats",
    "slimm_gpu_get_lca_counts", "slimm_gpu_get_lca_children", "slimm_gpu_fetch_bins", "slimm_gpu_get_uniq2_nz",
    "slimm_gpu_read_results", "slimm_gpu_enable_timing", "slimm_gpu_get_timings", "slimm_gpu_get_launch_count",
    "slimm_profile_rows", "slimm_gpu_set_scatter_mode", "slimm_gpu_set_taxa", "slimm_gpu_profile",
    "slimm_profile_db_is_tree_consistent", "slimm_gpu_set_shard", "slimm_gpu_get_slice_counts", "slimm_gpu_items_device",
    "slimm_gpu_accumulate_items", "slimm_gpu_stats_device", "slimm_gpu_profile_failed",
    "slimm_gpu_p2p_reserve", "slimm_gpu_p2p_connect", "slimm_gpu_split_to_peers", "slimm_gpu_accumulate_received", "slimm_gpu_p2p_disable",
    "slimm_gpu_slice_counts_device", "slimm_gpu_split_to_peers_device", "slimm_gpu_push_packed",
    "slimm_gpu_run_sharded_local",
]


class SlimmGpuError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [("n_refs", C.c_uint32), ("ref_len", C.c_void_p), ("lineage", C.c_void_p), ("bin_width", C.c_uint32),
                ("avg_read_length", C.c_uint32), ("reserve_records", C.c_uint64), ("device", C.c_int32),
                ("flags", C.c_uint32)]


class Summary(C.Structure):
    _fields_ = [("hits_count", C.c_uint32), ("matches_count", C.c_uint32), ("uniq_matches_count", C.c_uint32),
                ("uniq_matches_count2", C.c_uint32), ("reference_count", C.c_uint32), ("n_valid", C.c_uint32),
                ("failed_by_cov", C.c_uint32), ("failed_by_uniq_cov", C.c_uint32), ("failed_by_min_read", C.c_uint32),
                ("min_reads", C.c_uint32), ("coverage_cut_off", C.c_float), ("uniq_coverage_cut_off", C.c_float),
                ("n_pairs", C.c_uint64), ("n_bins", C.c_uint64), ("input_was_sorted", C.c_uint32), ("reserved", C.c_uint32)]


class _Row(C.Structure):
    _fields_ = [("taxon", C.c_uint32), ("kind", C.c_uint32), ("read_count", C.c_uint32), ("first_child", C.c_uint32),
                ("abundance", C.c_double)]


class _ProfileInput(C.Structure):
    _fields_ = [("n_refs", C.c_uint32), ("ref_len", C.c_void_p), ("lineage", C.c_void_p), ("n_taxa", C.c_uint64),
                ("taxa_id", C.c_void_p), ("taxa_rank", C.c_void_p), ("taxa_has_name", C.c_void_p),
                ("n_direct", C.c_uint64), ("direct_taxon", C.c_void_p), ("direct_count", C.c_void_p),
                ("n_children", C.c_uint64), ("child_taxon", C.c_void_p), ("child_ref", C.c_void_p),
                ("uniq_reads_count2", C.c_void_p), ("matches_count", C.c_uint32), ("avg_read_length", C.c_uint32),
                ("coverage_cut_off", C.c_float), ("abundance_cut_off", C.c_float), ("rank", C.c_uint32)]


_lib = None


def load_library():
    """Loads libslimm_gpu.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SlimmGpuError(f"{LIB_PATH} is missing: run `python -m slimm_b200.build` (needs nvcc); "
                            "there is no CPU fallback for the hot path")
    lib = C.CDLL(LIB_PATH)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    lib.slimm_gpu_strerror.restype = C.c_char_p
    lib.slimm_gpu_strerror.argtypes = [C.c_int]
    lib.slimm_gpu_last_error.restype = C.c_char_p
    lib.slimm_gpu_last_error.argtypes = [vp]
    lib.slimm_gpu_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.slimm_gpu_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    lib.slimm_gpu_destroy.argtypes = [vp]
    lib.slimm_gpu_reset.argtypes = [vp, u32, u32]
    lib.slimm_gpu_set_stream.argtypes = [vp, vp]
    lib.slimm_gpu_host_alloc.argtypes = [C.POINTER(vp), u64]
    lib.slimm_gpu_host_free.argtypes = [vp]
    lib.slimm_gpu_push.argtypes = [vp, vp, vp, vp, u64]
    lib.slimm_gpu_push_device.argtypes = [vp, vp, vp, vp, u64]
    lib.slimm_gpu_push_packed.argtypes = [vp, vp, vp, vp, u64]
    lib.slimm_gpu_sync_uploads.argtypes = [vp]
    lib.slimm_gpu_coverage.argtypes = [vp]
    lib.slimm_gpu_bins_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    lib.slimm_gpu_counters_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    lib.slimm_gpu_set_global_hits.argtypes = [vp, u64]
    lib.slimm_gpu_filter.argtypes = [vp, C.c_float, u32]
    lib.slimm_gpu_assign.argtypes = [vp]
    lib.slimm_gpu_assign_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    lib.slimm_gpu_run.argtypes = [vp, C.c_float, u32]
    lib.slimm_gpu_get_summary.argtypes = [vp, C.POINTER(Summary)]
    lib.slimm_gpu_get_ref_stats.argtypes = [vp] + [vp] * 8
    lib.slimm_gpu_get_lca_counts.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    lib.slimm_gpu_get_lca_children.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    lib.slimm_gpu_fetch_bins.argtypes = [vp, C.c_int, u32, vp, u32]
    lib.slimm_gpu_get_uniq2_nz.argtypes = [vp, vp]
    lib.slimm_gpu_read_results.argtypes = [vp, vp, vp, vp, u64, C.POINTER(u64)]
    lib.slimm_gpu_enable_timing.argtypes = [vp, C.c_int]
    lib.slimm_gpu_get_timings.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    lib.slimm_gpu_get_launch_count.argtypes = [vp, C.POINTER(u64)]
    lib.slimm_profile_rows.argtypes = [C.POINTER(_ProfileInput), C.POINTER(_Row), u64, C.POINTER(u64)]
    lib.slimm_gpu_set_scatter_mode.argtypes = [vp, C.c_int]
    lib.slimm_gpu_set_taxa.argtypes = [vp, u64, vp, vp, vp]
    lib.slimm_gpu_profile.argtypes = [vp, u32, C.c_float, C.POINTER(_Row), u64, C.POINTER(u64)]
    lib.slimm_gpu_profile_failed.argtypes = [vp, C.POINTER(u32)]
    lib.slimm_gpu_set_shard.argtypes = [vp, u32, u32]
    lib.slimm_gpu_get_slice_counts.argtypes = [vp, vp, u32, C.POINTER(u32)]
    lib.slimm_gpu_items_device.argtypes = [vp, C.POINTER(vp)]
    lib.slimm_gpu_accumulate_items.argtypes = [vp, vp, u64]
    lib.slimm_gpu_p2p_reserve.argtypes = [vp, u64, vp]
    lib.slimm_gpu_p2p_connect.argtypes = [vp, vp, u32]
    lib.slimm_gpu_split_to_peers.argtypes = [vp, vp, C.POINTER(u64)]
    lib.slimm_gpu_accumulate_received.argtypes = [vp]
    lib.slimm_gpu_slice_counts_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u32)]
    lib.slimm_gpu_split_to_peers_device.argtypes = [vp, vp]
    lib.slimm_gpu_p2p_disable.argtypes = [vp]
    lib.slimm_gpu_stats_device.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    lib.slimm_profile_db_is_tree_consistent.argtypes = [u32, vp, u64, vp, vp, vp, C.POINTER(C.c_int)]
    _lib = lib
    return lib


def device_count() -> int:
    n = C.c_int(0)
    load_library().slimm_gpu_device_count(C.byref(n))
    return n.value


@dataclass
class RefStats:
    reads_count: np.ndarray
    uniq_reads_count: np.ndarray
    uniq_reads_count2: Optional[np.ndarray]
    nz_bins: np.ndarray
    uniq_nz_bins: np.ndarray
    cov_percent: np.ndarray
    uniq_cov_percent: np.ndarray
    valid: np.ndarray


class SlimmGpu:
    """One profiling context on one GPU (one sample at a time; ``reset`` starts the next)."""

    def __init__(self, ref_len, lineage, bin_width: int, avg_read_length: int, device: int = 0,
                 flags: int = 0, reserve_records: int = 0):
        self._lib = load_library()
        self.ref_len = np.ascontiguousarray(ref_len, dtype=np.uint32)
        self.lineage = np.ascontiguousarray(lineage, dtype=np.uint32).reshape(-1, 8)
        self.n_refs = int(self.ref_len.size)
        if self.lineage.shape[0] != self.n_refs:
            raise ValueError("lineage must be [n_refs, 8]")
        self.bin_width, self.avg_read_length = int(bin_width), int(avg_read_length)
        cfg = _Config(self.n_refs, self.ref_len.ctypes.data, self.lineage.ctypes.data, self.bin_width,
                      self.avg_read_length, reserve_records, device, flags)
        self._ctx = C.c_void_p()
        rc = self._lib.slimm_gpu_create(C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.slimm_gpu_last_error(self._ctx).decode() if self._ctx else ""
            if self._ctx:
                self._lib.slimm_gpu_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise SlimmGpuError(f"slimm_gpu_create: {self._lib.slimm_gpu_strerror(rc).decode()} {msg}")
        self._keep = []   # host arrays that must outlive asynchronous uploads

    # -- plumbing -----------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            raise SlimmGpuError(f"{what}: {self._lib.slimm_gpu_strerror(rc).decode()}: "
                                f"{self._lib.slimm_gpu_last_error(self._ctx).decode()}")

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.slimm_gpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int):
        """Run the stages on the caller's stream.  A framework's default stream has handle 0, which the C ABI reads
        as "library-owned stream": it is passed as cudaStreamLegacy (0x1) instead, so that the library's kernels and
        the caller's work (NCCL collectives issued on that stream) stay ordered."""
        self._check(self._lib.slimm_gpu_set_stream(self._ctx, C.c_void_p(cuda_stream if cuda_stream else 1)), "set_stream")

    def reset(self, bin_width: int = 0, avg_read_length: int = 0):
        self._check(self._lib.slimm_gpu_reset(self._ctx, bin_width, avg_read_length), "reset")
        self._keep.clear()

    # -- ingest -------------------------------------------------------------------------------
    def push(self, read_id, ref_id, begin_pos):
        """Append a batch of kept records (host arrays; replaces reference src/slimm.hpp:194-213)."""
        a = np.ascontiguousarray(read_id, dtype=np.uint32)
        b = np.ascontiguousarray(ref_id, dtype=np.uint32)
        c = np.ascontiguousarray(begin_pos, dtype=np.int32)
        if not (a.size == b.size == c.size):
            raise ValueError("record arrays differ in length")
        self._keep += [a, b, c]
        self._check(self._lib.slimm_gpu_push(self._ctx, a.ctypes.data, b.ctypes.data, c.ctypes.data, a.size), "push")

    def push_ptrs(self, read_id_ptr: int, ref_id_ptr: int, begin_pos_ptr: int, n: int):
        self._check(self._lib.slimm_gpu_push(self._ctx, read_id_ptr, ref_id_ptr, begin_pos_ptr, n), "push")

    def push_packed(self, new_read_bits, ref_id16, begin_pos):
        """Grouped input in the 6.125-byte wire format (include/slimm_gpu.h): one "new read" bit, a 16-bit reference id and
        the position per record."""
        a = np.ascontiguousarray(new_read_bits, dtype=np.uint32)
        b = np.ascontiguousarray(ref_id16, dtype=np.uint16)
        c = np.ascontiguousarray(begin_pos, dtype=np.int32)
        if b.size != c.size or a.size * 32 < b.size:
            raise ValueError("record arrays differ in length")
        self._keep += [a, b, c]
        self._check(self._lib.slimm_gpu_push_packed(self._ctx, a.ctypes.data, b.ctypes.data, c.ctypes.data, b.size), "push_packed")

    def push_packed_ptrs(self, bits_ptr: int, ref16_ptr: int, begin_pos_ptr: int, n: int):
        self._check(self._lib.slimm_gpu_push_packed(self._ctx, bits_ptr, ref16_ptr, begin_pos_ptr, n), "push_packed")

    def push_device(self, read_id_ptr: int, ref_id_ptr: int, begin_pos_ptr: int, n: int):
        self._check(self._lib.slimm_gpu_push_device(self._ctx, read_id_ptr, ref_id_ptr, begin_pos_ptr, n), "push_device")

    def sync_uploads(self):
        self._check(self._lib.slimm_gpu_sync_uploads(self._ctx), "sync_uploads")
        self._keep.clear()

    # -- stages -------------------------------------------------------------------------------
    def coverage(self):
        self._check(self._lib.slimm_gpu_coverage(self._ctx), "coverage")

    def filter(self, cov_cut_off: float = 0.95, min_reads: int = 0):
        self._check(self._lib.slimm_gpu_filter(self._ctx, cov_cut_off, min_reads), "filter")

    def assign(self):
        self._check(self._lib.slimm_gpu_assign(self._ctx), "assign")

    def run(self, cov_cut_off: float = 0.95, min_reads: int = 0):
        self._check(self._lib.slimm_gpu_run(self._ctx, cov_cut_off, min_reads), "run")

    def bins_device(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.slimm_gpu_bins_device(self._ctx, C.byref(p), C.byref(n)), "bins_device")
        return p.value, n.value

    def counters_device(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.slimm_gpu_counters_device(self._ctx, C.byref(p), C.byref(n)), "counters_device")
        return p.value, n.value

    def assign_device(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.slimm_gpu_assign_device(self._ctx, C.byref(p), C.byref(n)), "assign_device")
        return p.value, n.value

    # -- sharded runs (several GPUs) -----------------------------------------------------------
    def set_shard(self, rank: int, n_ranks: int):
        self._check(self._lib.slimm_gpu_set_shard(self._ctx, rank, n_ranks), "set_shard")

    def slice_counts(self) -> np.ndarray:
        n = C.c_uint32()
        self._check(self._lib.slimm_gpu_get_slice_counts(self._ctx, None, 0, C.byref(n)), "get_slice_counts")
        counts = np.zeros(n.value, dtype=np.uint32)
        self._check(self._lib.slimm_gpu_get_slice_counts(self._ctx, counts.ctypes.data, n.value, C.byref(n)), "get_slice_counts")
        return counts

    def items_device(self) -> int:
        p = C.c_void_p()
        self._check(self._lib.slimm_gpu_items_device(self._ctx, C.byref(p)), "items_device")
        return p.value or 0

    def accumulate_items(self, items_ptr: int, n_items: int):
        self._check(self._lib.slimm_gpu_accumulate_items(self._ctx, C.c_void_p(items_ptr), n_items), "accumulate_items")

    def p2p_reserve(self, cap_items: int) -> bytes:
        """Allocates the receive buffer of the peer-to-peer item exchange; returns its 64-byte CUDA IPC handle."""
        h = C.create_string_buffer(64)
        self._check(self._lib.slimm_gpu_p2p_reserve(self._ctx, cap_items, h), "p2p_reserve")
        return h.raw

    def p2p_connect(self, handles: bytes, n_ranks: int):
        """Maps every rank's receive buffer (handles of all ranks in rank order, 64 bytes each)."""
        if len(handles) != 64 * n_ranks:
            raise ValueError("need one 64-byte handle per rank")
        buf = C.create_string_buffer(handles, len(handles))
        self._check(self._lib.slimm_gpu_p2p_connect(self._ctx, buf, n_ranks), "p2p_connect")
        self.p2p = True

    def split_to_peers(self, all_counts: np.ndarray) -> int:
        """all_counts[n_ranks][n_slices]: items per slice of every rank.  Splits this rank's items straight into the
        owners' receive buffers; returns how many items this rank receives."""
        t = np.ascontiguousarray(all_counts, dtype=np.uint32)
        n = C.c_uint64(0)
        self._check(self._lib.slimm_gpu_split_to_peers(self._ctx, t.ctypes.data, C.byref(n)), "split_to_peers")
        return int(n.value)

    def slice_counts_device(self) -> Tuple[int, int]:
        """Device pointer to this rank's items per histogram slice (u32) and the number of slices."""
        p, n = C.c_void_p(), C.c_uint32()
        self._check(self._lib.slimm_gpu_slice_counts_device(self._ctx, C.byref(p), C.byref(n)), "slice_counts_device")
        return p.value, n.value

    def split_to_peers_device(self, all_counts_ptr: int):
        """The peer-to-peer split planned on the device from the all-gathered counts ([n_ranks][n_slices] u32, device memory)."""
        self._check(self._lib.slimm_gpu_split_to_peers_device(self._ctx, C.c_void_p(all_counts_ptr)), "split_to_peers_device")

    def p2p_disable(self):
        self._check(self._lib.slimm_gpu_p2p_disable(self._ctx), "p2p_disable")
        self.p2p = False

    def accumulate_received(self):
        self._check(self._lib.slimm_gpu_accumulate_received(self._ctx), "accumulate_received")

    def stats_device(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.slimm_gpu_stats_device(self._ctx, C.byref(p), C.byref(n)), "stats_device")
        return p.value, n.value

    def set_global_hits(self, hits: int):
        self._check(self._lib.slimm_gpu_set_global_hits(self._ctx, hits), "set_global_hits")

    # -- results ------------------------------------------------------------------------------
    def summary(self) -> Summary:
        s = Summary()
        self._check(self._lib.slimm_gpu_get_summary(self._ctx, C.byref(s)), "get_summary")
        return s

    def ref_stats(self, with_uniq2: bool = True) -> RefStats:
        G = self.n_refs
        u = [np.zeros(G, dtype=np.uint32) for _ in range(5)]
        f = [np.zeros(G, dtype=np.float32) for _ in range(2)]
        v = np.zeros(G, dtype=np.uint8)
        self._check(self._lib.slimm_gpu_get_ref_stats(
            self._ctx, u[0].ctypes.data, u[1].ctypes.data, u[2].ctypes.data if with_uniq2 else None,
            u[3].ctypes.data, u[4].ctypes.data, f[0].ctypes.data, f[1].ctypes.data, v.ctypes.data), "get_ref_stats")
        return RefStats(u[0], u[1], u[2] if with_uniq2 else None, u[3], u[4], f[0], f[1], v)

    def lca_counts(self) -> Dict[int, int]:
        n = C.c_uint64()
        self._check(self._lib.slimm_gpu_get_lca_counts(self._ctx, None, None, 0, C.byref(n)), "get_lca_counts")
        t = np.zeros(n.value, dtype=np.uint32)
        c = np.zeros(n.value, dtype=np.uint32)
        self._check(self._lib.slimm_gpu_get_lca_counts(self._ctx, t.ctypes.data, c.ctypes.data, n.value, C.byref(n)),
                    "get_lca_counts")
        return dict(zip(t.tolist(), c.tolist()))

    def lca_children(self) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._lib.slimm_gpu_get_lca_children(self._ctx, None, None, 0, C.byref(n)), "get_lca_children")
        t = np.zeros(n.value, dtype=np.uint32)
        r = np.zeros(n.value, dtype=np.uint32)
        self._check(self._lib.slimm_gpu_get_lca_children(self._ctx, t.ctypes.data, r.ctypes.data, n.value, C.byref(n)),
                    "get_lca_children")
        return np.stack([t, r], axis=1)

    def fetch_bins(self, which: int, ref: int) -> np.ndarray:
        nb = int(self.ref_len[ref]) // self.bin_width + 1
        out = np.zeros(nb, dtype=np.uint32)
        self._check(self._lib.slimm_gpu_fetch_bins(self._ctx, which, ref, out.ctypes.data, nb), "fetch_bins")
        return out

    def uniq2_nz(self) -> np.ndarray:
        out = np.zeros(self.n_refs, dtype=np.uint32)
        self._check(self._lib.slimm_gpu_get_uniq2_nz(self._ctx, out.ctypes.data), "get_uniq2_nz")
        return out

    def read_results(self):
        n = C.c_uint64()
        self._check(self._lib.slimm_gpu_read_results(self._ctx, None, None, None, 0, C.byref(n)), "read_results")
        rid = np.zeros(n.value, dtype=np.uint32)
        kind = np.zeros(n.value, dtype=np.uint8)
        val = np.zeros(n.value, dtype=np.uint32)
        self._check(self._lib.slimm_gpu_read_results(self._ctx, rid.ctypes.data, kind.ctypes.data, val.ctypes.data,
                                                     n.value, C.byref(n)), "read_results")
        return rid, kind, val

    # -- profile tail -------------------------------------------------------------------------
    def set_taxa(self, taxa):
        """``taxa``: dict taxid -> (rank, name) (db.taxid__name) or the arrays from :func:`taxa_arrays`."""
        tid, trank, tname = taxa_arrays(taxa) if isinstance(taxa, dict) else taxa
        self._taxa_keep = (tid, trank, tname)
        self._rows_cap = 2 * int(tid.size) + 8
        self._rows_buf = (_Row * self._rows_cap)()
        self._check(self._lib.slimm_gpu_set_taxa(self._ctx, tid.size, tid.ctypes.data, trank.ctypes.data,
                                                 tname.ctypes.data), "set_taxa")

    def profile_raw(self, rank: int = 1, abundance_cut_off: float = 0.01):
        """Rank aggregation + abundances straight from the device results; returns (ctypes rows, n)."""
        n = C.c_uint64()
        self._check(self._lib.slimm_gpu_profile(self._ctx, rank, abundance_cut_off, self._rows_buf, self._rows_cap,
                                                C.byref(n)), "profile")
        return self._rows_buf, n.value

    def profile(self, rank: int = 1, abundance_cut_off: float = 0.01) -> List["ProfileRow"]:
        rows, n = self.profile_raw(rank, abundance_cut_off)
        return [ProfileRow(r.taxon, r.kind, r.read_count, r.first_child, r.abundance) for r in rows[:n]]

    def set_scatter_mode(self, mode: int):
        self._check(self._lib.slimm_gpu_set_scatter_mode(self._ctx, mode), "set_scatter_mode")

    # -- instrumentation ----------------------------------------------------------------------
    def enable_timing(self, on: bool = True):
        self._check(self._lib.slimm_gpu_enable_timing(self._ctx, int(on)), "enable_timing")

    def timings(self) -> Dict[str, float]:
        ms = (C.c_float * len(TIMING_NAMES))()
        self._check(self._lib.slimm_gpu_get_timings(self._ctx, ms, len(TIMING_NAMES)), "get_timings")
        return dict(zip(TIMING_NAMES, [float(x) for x in ms]))

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.slimm_gpu_get_launch_count(self._ctx, C.byref(n)), "get_launch_count")
        return n.value


@dataclass
class ProfileRow:
    taxon: int
    kind: int            # 0 plain, 1 "<parent>*", 2 "0*"
    read_count: int
    first_child: int
    abundance: float


def taxa_arrays(taxa: Dict[int, Tuple[int, str]]):
    """taxid -> (rank, name) as the three arrays slimm_profile_rows takes."""
    tid = np.fromiter(taxa.keys(), dtype=np.uint32, count=len(taxa))
    trank = np.fromiter((v[0] for v in taxa.values()), dtype=np.uint8, count=len(taxa))
    tname = np.fromiter((1 if v[1] != "" else 0 for v in taxa.values()), dtype=np.uint8, count=len(taxa))
    return tid, trank, tname


def profile_rows_arrays(ref_len, lineage, taxa_arr, direct: Dict[int, int], children: np.ndarray, uniq_reads_count2,
                        matches_count: int, avg_read_length: int, coverage_cut_off: float,
                        abundance_cut_off: float = 0.01, rank: int = 1) -> List[ProfileRow]:
    """Host tail of the path: rank aggregation + abundances (slimm_profile_rows; replaces reference
    src/slimm.hpp:560-610 and the numeric part of :733-843)."""
    lib = load_library()
    ref_len = np.ascontiguousarray(ref_len, dtype=np.uint32)
    lineage = np.ascontiguousarray(lineage, dtype=np.uint32)
    tid, trank, tname = taxa_arr
    dt = np.fromiter(direct.keys(), dtype=np.uint32, count=len(direct))
    dc = np.fromiter(direct.values(), dtype=np.uint32, count=len(direct))
    ch = np.ascontiguousarray(children, dtype=np.uint32).reshape(-1, 2)
    ct, cr = np.ascontiguousarray(ch[:, 0]), np.ascontiguousarray(ch[:, 1])
    u2 = np.ascontiguousarray(uniq_reads_count2, dtype=np.uint32)
    inp = _ProfileInput(ref_len.size, ref_len.ctypes.data, lineage.ctypes.data, tid.size, tid.ctypes.data,
                        trank.ctypes.data, tname.ctypes.data, dt.size, dt.ctypes.data, dc.ctypes.data, ct.size,
                        ct.ctypes.data, cr.ctypes.data, u2.ctypes.data, matches_count, avg_read_length,
                        coverage_cut_off, abundance_cut_off, rank)
    n = C.c_uint64()
    cap = 2 * int(tid.size) + 8
    rows = (_Row * cap)()
    rc = lib.slimm_profile_rows(C.byref(inp), rows, cap, C.byref(n))
    if rc != 0:
        raise SlimmGpuError(f"slimm_profile_rows: {lib.slimm_gpu_strerror(rc).decode()}")
    return [ProfileRow(r.taxon, r.kind, r.read_count, r.first_child, r.abundance) for r in rows[: n.value]]


def profile_rows(ref_len, lineage, taxa: Dict[int, Tuple[int, str]], direct, children, uniq_reads_count2,
                 matches_count: int, avg_read_length: int, coverage_cut_off: float,
                 abundance_cut_off: float = 0.01, rank: int = 1) -> List[ProfileRow]:
    return profile_rows_arrays(ref_len, lineage, taxa_arrays(taxa), direct, children, uniq_reads_count2,
                               matches_count, avg_read_length, coverage_cut_off, abundance_cut_off, rank)


def db_is_tree_consistent(lineage, taxa) -> bool:
    """True when slimm_gpu_profile can run the rank reduction on the device for this database."""
    lib = load_library()
    lin = np.ascontiguousarray(lineage, dtype=np.uint32).reshape(-1, 8)
    tid, trank, tname = taxa_arrays(taxa) if isinstance(taxa, dict) else taxa
    out = C.c_int(0)
    rc = lib.slimm_profile_db_is_tree_consistent(lin.shape[0], lin.ctypes.data, tid.size, tid.ctypes.data, trank.ctypes.data,
                                                 tname.ctypes.data, C.byref(out))
    if rc != 0:
        raise SlimmGpuError(f"slimm_profile_db_is_tree_consistent: {lib.slimm_gpu_strerror(rc).decode()}")
    return bool(out.value)


class _DevicePointer:
    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_tensor(ptr: int, n: int, dtype, device):
    """Zero-copy torch view of library-owned device memory (for NCCL reductions across ranks)."""
    import torch
    typestr = {torch.int32: "<i4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    t = torch.as_tensor(_DevicePointer(ptr, n, typestr), device=device)
    if n and t.data_ptr() != ptr:
        raise SlimmGpuError("torch.as_tensor copied the device buffer instead of wrapping it")
    return t
