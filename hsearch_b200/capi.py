"""ctypes binding of include/hsearch_b200.h (the C ABI).

This is the only way Python reaches the product: there is no Python or CPU
implementation of the path.  If libhsearch_b200.so is missing it is built
in-tree (hsearch_b200/build.py); if it cannot be loaded the import fails loudly.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HS_LIBRARY") or os.path.join(HERE, "libhsearch_b200.so")  # HS_LIBRARY: experiments only

HS_OK = 0
HS_ERR_INVALID, HS_ERR_CUDA, HS_ERR_CAPACITY, HS_ERR_UNSUPPORTED, HS_ERR_NOMEM, HS_ERR_COMM = -1, -2, -3, -4, -5, -6
HS_TABLE_FULL, HS_TABLE_PRINT6 = 0, 1
HS_METRIC_EUCLID_FP64, HS_METRIC_BLOSUM_INT = 0, 1
HS_PRED_D2_LE_R2, HS_PRED_SQRT_LE_R = 0, 1
HS_FLAG_SORT_HITS, HS_FLAG_HASH_EXACT, HS_FLAG_HASH_AUDIT, HS_FLAG_SCALAR_FILTER = 1, 2, 4, 8

# every symbol include/hsearch_b200.h declares
EXPORTS = [
    "hs_create", "hs_destroy", "hs_last_error", "hs_get_stats", "hs_device_available", "hs_get_stream", "hs_get_coordinates",
    "hs_get_blosum_metric", "hs_set_coordinates", "hs_letter_to_code", "hs_proteindb_code",
    "hs_generate_projection", "hs_set_projection", "hs_load_fragments", "hs_load_fragments_dev",
    "hs_extract_windows", "hs_num_fragments", "hs_hash", "hs_get_keys", "hs_pack_key_string", "hs_build_index",
    "hs_table_sizes", "hs_get_table", "hs_search_points", "hs_search_codes", "hs_search_points_dev",
    "hs_bruteforce_codes", "hs_bruteforce_points", "hs_bruteforce_points_dev", "hs_cluster", "hs_comm_init", "hs_comm_unique_id",
    "hs_greedy_cluster", "hs_union_find", "hs_parse_fasta", "hs_klsh_generate", "hs_kmer3_klsh", "hs_orf6",
    "hs_evaluate_recall", "hs_evaluate_recall_dev",
    "hs_search_points_compact", "hs_expand_hits", "hs_hits_checksum", "hs_hits_checksum_dev", "hs_hash_audit", "hs_comm_reserve", "hs_comm_result", "hs_protein_id", "hs_fragment_name",
    "hs_parse_fasta_gpu", "hs_get_blosum_filter_embedding",
]


class Params(C.Structure):
    _fields_ = [("len", C.c_uint32), ("K", C.c_uint32), ("L", C.c_uint32), ("W", C.c_double), ("R", C.c_double),
                ("table_variant", C.c_uint32), ("metric", C.c_uint32), ("predicate", C.c_uint32),
                ("flags", C.c_uint32)]


RECALL_BINS = 500


class Recall(C.Structure):
    """hs_recall"""
    _fields_ = [("tp", C.c_double), ("fn", C.c_double), ("n_tp", C.c_uint64), ("n_fn", C.c_uint64),
                ("n_extra", C.c_uint64), ("tp_bin", C.c_uint64 * RECALL_BINS), ("fn_bin", C.c_uint64 * RECALL_BINS)]


class CompactHits(C.Structure):
    """hs_compact_hits"""
    _fields_ = [("offsets", C.POINTER(C.c_uint64)), ("idt", C.POINTER(C.c_uint32)), ("dist2", C.POINTER(C.c_double)),
                ("cap", C.c_uint64), ("id_bits", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("n_fragments", C.c_uint64), ("guard_hits", C.c_uint64), ("guard_corrected", C.c_uint64),
                ("residual_flips", C.c_uint64), ("n_candidates", C.c_uint64), ("n_survivors", C.c_uint64),
                ("n_hits", C.c_uint64), ("n_edges", C.c_uint64), ("n_work_items", C.c_uint64),
                ("key_words", C.c_uint32), ("sort_passes", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("rank_path", C.c_uint32),
                ("ms_hash", C.c_float), ("ms_sort", C.c_float), ("ms_group", C.c_float), ("ms_permute", C.c_float),
                ("ms_sort_upsweep", C.c_float), ("ms_sort_scan", C.c_float), ("ms_sort_downsweep", C.c_float),
                ("ms_qhash", C.c_float), ("ms_probe", C.c_float), ("ms_filter", C.c_float), ("ms_exact", C.c_float),
                ("ms_hitsort", C.c_float), ("ms_total", C.c_float), ("ms_filter_tc", C.c_float),
                ("ms_host", C.c_float), ("n_candidates_tc", C.c_uint64), ("hash_sort_fallbacks", C.c_uint64),
                ("segsort_lists", C.c_uint64), ("segsort_fallbacks", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


HIT_DTYPE = np.dtype([("query", "<u4"), ("table_first", "<u4"), ("db_id", "<u8"), ("dist2", "<f8")])
assert HIT_DTYPE.itemsize == 24

_lib = None


def load(build_if_missing=True):
    """Load libhsearch_b200.so (building it first when absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise ImportError(f"{LIB_PATH} is missing; run python -m hsearch_b200.build")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    vp, u8p, u32p, u64p, i32p, dblp = (C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_double))
    lib.hs_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Params)]
    lib.hs_destroy.argtypes = [vp]
    lib.hs_destroy.restype = None
    lib.hs_last_error.restype = C.c_char_p
    lib.hs_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.hs_get_stream.argtypes = [vp, C.POINTER(vp)]
    lib.hs_get_coordinates.argtypes = [C.c_uint32, dblp]
    lib.hs_get_blosum_metric.argtypes = [i32p]
    lib.hs_set_coordinates.argtypes = [vp, dblp]
    lib.hs_letter_to_code.argtypes = [C.c_char]
    lib.hs_proteindb_code.argtypes = [C.c_char]
    lib.hs_generate_projection.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, dblp, dblp]
    lib.hs_set_projection.argtypes = [vp, dblp, dblp]
    lib.hs_load_fragments.argtypes = [vp, u8p, C.c_uint64, C.c_uint64]
    lib.hs_load_fragments_dev.argtypes = [vp, vp, C.c_uint64, C.c_uint64]
    lib.hs_extract_windows.argtypes = [vp, u8p, u32p, C.c_uint32, C.c_uint32, C.c_uint64, u32p, C.c_uint64, u64p]
    lib.hs_num_fragments.argtypes = [vp]
    lib.hs_num_fragments.restype = C.c_uint64
    lib.hs_hash.argtypes = [vp, i32p]
    lib.hs_get_keys.argtypes = [vp, C.c_uint32, u64p]
    lib.hs_pack_key_string.argtypes = [C.c_char_p, C.c_uint32, u64p]
    lib.hs_build_index.argtypes = [vp]
    lib.hs_table_sizes.argtypes = [vp, u64p]
    lib.hs_get_table.argtypes = [vp, C.c_uint32, u32p, u32p]
    lib.hs_search_points.argtypes = [vp, dblp, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hs_search_codes.argtypes = [vp, u8p, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hs_search_points_dev.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hs_bruteforce_codes.argtypes = [vp, u8p, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hs_bruteforce_points.argtypes = [vp, dblp, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hs_bruteforce_points_dev.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hs_cluster.argtypes = [vp, u32p]
    lib.hs_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.hs_comm_unique_id.argtypes = [vp]
    lib.hs_greedy_cluster.argtypes = [vp, u32p, u32p, u8p]
    lib.hs_union_find.argtypes = [vp, C.c_uint32, u32p, u32p, C.c_uint64, u32p]
    lib.hs_parse_fasta.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, u64p, C.c_uint64, u64p, u32p,
                                   C.c_uint64, u32p, u32p, u64p]
    lib.hs_get_blosum_filter_embedding.argtypes = [dblp]
    lib.hs_parse_fasta_gpu.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, u64p, C.c_uint64, u64p, u32p,
                                       C.c_uint64, u32p, u32p, u64p]
    lib.hs_klsh_generate.argtypes = [C.c_uint32, C.c_uint32, C.c_double, dblp, dblp, dblp]
    lib.hs_kmer3_klsh.argtypes = [vp, C.c_char_p, u64p, C.c_uint32, dblp, dblp, dblp, C.c_uint32, u32p, u64p, u8p,
                                  u64p]
    lib.hs_orf6.argtypes = [vp, C.c_char_p, u64p, C.c_uint32, C.c_char_p, C.c_uint64, i32p]
    lib.hs_evaluate_recall.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.POINTER(Recall)]
    lib.hs_evaluate_recall_dev.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.POINTER(Recall)]
    lib.hs_search_points_compact.argtypes = [vp, dblp, C.c_uint32, C.POINTER(CompactHits), u64p]
    lib.hs_expand_hits.argtypes = [C.POINTER(CompactHits), C.c_uint32, C.c_uint64, vp]
    lib.hs_hash_audit.argtypes = [vp, u64p]
    lib.hs_protein_id.argtypes = [vp, u32p, C.c_uint32, u32p, C.c_uint64, u32p]
    lib.hs_fragment_name.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint64, C.c_char_p,
                                     C.c_uint64]
    lib.hs_comm_reserve.argtypes = [vp, C.c_uint64]
    lib.hs_comm_result.argtypes = [vp, C.POINTER(vp), u64p]
    lib.hs_hits_checksum.argtypes = [vp, C.c_uint64, u64p]
    lib.hs_hits_checksum_dev.argtypes = [vp, vp, C.c_uint64, u64p]
    _lib = lib
    return lib


class HsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"hsearch_b200 error {code}: {msg}")
        self.code = code


def check(rc):
    if rc != HS_OK:
        raise HsError(rc, load().hs_last_error().decode(errors="replace"))


def ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))
