"""ctypes binding of libdiagon_b200.so (include/diagon_b200_c_api.h + include/dgpu_engine.h).

The shared library is the product; this module only declares its prototypes. Loading fails loudly
when the library has not been built (run `make` or `python -c "import __graft_entry__ as g; g.build()"`):
there is no Python or CPU fallback for any entry point.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DGPU_LIB: another build of the same library (debugging only, e.g. the -DDGPU_CHECK build with device-side bounds checks)
LIB_PATH = os.environ.get("DGPU_LIB") or os.path.join(_HERE, "libdiagon_b200.so")


class CorpusSpec(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("num_docs", C.c_uint32),
        ("vocab", C.c_uint32),
        ("zipf_s", C.c_double),
        ("len_mu", C.c_double),
        ("len_sigma", C.c_double),
        ("len_min", C.c_uint32),
        ("len_max", C.c_uint32),
        ("num_segments", C.c_uint32),
        ("with_price", C.c_int32),
    ]


class Results(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("counts", C.c_void_p), ("total_hits", C.c_void_p)]


# name -> (restype, argtypes); every symbol the two headers declare
PROTOTYPES = {
    # ---- include/diagon_b200_c_api.h, part 1 (mirrors diagon_c_api.h)
    "diagon_last_error": (C.c_char_p, []),
    "diagon_clear_error": (None, []),
    "diagon_open_fs_directory": (C.c_void_p, [C.c_char_p]),
    "diagon_open_mmap_directory": (C.c_void_p, [C.c_char_p]),
    "diagon_close_directory": (None, [C.c_void_p]),
    "diagon_open_index_reader": (C.c_void_p, [C.c_void_p]),
    "diagon_reader_num_docs": (C.c_int64, [C.c_void_p]),
    "diagon_reader_max_doc": (C.c_int64, [C.c_void_p]),
    "diagon_reader_get_segment_count": (C.c_int, [C.c_void_p]),
    "diagon_close_index_reader": (None, [C.c_void_p]),
    "diagon_create_index_searcher": (C.c_void_p, [C.c_void_p]),
    "diagon_search": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int]),
    "diagon_count": (C.c_int, [C.c_void_p, C.c_void_p]),
    "diagon_free_index_searcher": (None, [C.c_void_p]),
    "diagon_create_term": (C.c_void_p, [C.c_char_p, C.c_char_p]),
    "diagon_free_term": (None, [C.c_void_p]),
    "diagon_create_term_query": (C.c_void_p, [C.c_void_p]),
    "diagon_create_numeric_range_query": (C.c_void_p, [C.c_char_p, C.c_double, C.c_double, C.c_bool, C.c_bool]),
    "diagon_create_bool_query": (C.c_void_p, []),
    "diagon_bool_query_add_must": (None, [C.c_void_p, C.c_void_p]),
    "diagon_bool_query_add_should": (None, [C.c_void_p, C.c_void_p]),
    "diagon_bool_query_add_filter": (None, [C.c_void_p, C.c_void_p]),
    "diagon_bool_query_add_must_not": (None, [C.c_void_p, C.c_void_p]),
    "diagon_bool_query_set_minimum_should_match": (None, [C.c_void_p, C.c_int]),
    "diagon_bool_query_build": (C.c_void_p, [C.c_void_p]),
    "diagon_free_query": (None, [C.c_void_p]),
    "diagon_free_bool_query_builder": (None, [C.c_void_p]),
    "diagon_top_docs_total_hits": (C.c_int64, [C.c_void_p]),
    "diagon_top_docs_max_score": (C.c_float, [C.c_void_p]),
    "diagon_top_docs_score_docs_length": (C.c_int, [C.c_void_p]),
    "diagon_top_docs_score_doc_at": (C.c_void_p, [C.c_void_p, C.c_int]),
    "diagon_score_doc_get_doc": (C.c_int, [C.c_void_p]),
    "diagon_score_doc_get_score": (C.c_float, [C.c_void_p]),
    "diagon_free_top_docs": (None, [C.c_void_p]),
    # ---- part 2 (dgpu_*)
    "dgpu_builder_create": (C.c_void_p, []),
    "dgpu_builder_free": (None, [C.c_void_p]),
    "dgpu_builder_add_segment": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "dgpu_builder_set_field_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "dgpu_builder_add_term": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "dgpu_builder_add_numeric_doc_values": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]),
    "dgpu_builder_finish": (C.c_void_p, [C.c_void_p, C.c_int]),
    "dgpu_open_dump": (C.c_void_p, [C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "dgpu_open_index": (C.c_void_p, [C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "dgpu_reader_image_hash": (C.c_uint64, [C.c_void_p]),
    "dgpu_reader_save_image": (C.c_int, [C.c_void_p, C.c_char_p]),
    "dgpu_open_image": (C.c_void_p, [C.c_char_p, C.c_int]),
    "dgpu_named_corpus": (C.c_int, [C.c_char_p, C.c_double, C.POINTER(CorpusSpec)]),
    "dgpu_write_synthetic_dump": (C.c_int, [C.POINTER(CorpusSpec), C.c_char_p]),
    "dgpu_query_log_text": (C.c_void_p, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(C.c_int64)]),
    "dgpu_free_text": (None, [C.c_void_p]),
    "dgpu_open_synthetic": (C.c_void_p, [C.POINTER(CorpusSpec), C.c_int, C.c_int, C.c_int]),
    "dgpu_reader_num_terms": (C.c_int64, [C.c_void_p]),
    "dgpu_debug_set_fast_text_compile": (None, [C.c_int]),
    "dgpu_reader_term_id": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int64]),
    "dgpu_reader_term_bytes": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]),
    "dgpu_reader_dictionary_frozen": (C.c_int, [C.c_void_p]),
    "dgpu_reader_get_doc_freqs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "dgpu_reader_set_doc_freqs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "dgpu_reader_get_field_totals": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dgpu_reader_set_field_totals": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int64]),
    "dgpu_reader_image_bytes": (C.c_int64, [C.c_void_p]),
    "dgpu_reader_num_postings": (C.c_int64, [C.c_void_p]),
    "dgpu_reader_engine": (C.c_void_p, [C.c_void_p]),
    "dgpu_reader_decode_term": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]),
    "dgpu_search_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgpu_search_batch_text": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "dgpu_submit_batch_text": (C.c_void_p, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32]),
    "dgpu_batch_ticket_queries": (C.c_int32, [C.c_void_p]),
    "dgpu_sharded_submit_batch_text": (C.c_void_p, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32]),
    "dgpu_collect_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "dgpu_batch_ticket_free": (None, [C.c_void_p]),
    "dgpu_stage_batch_text": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, C.c_void_p]),
    "dgpu_compile_batch_text": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64]),
    "dgpu_comm_unique_id": (C.c_int, [C.c_void_p]),
    "dgpu_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "dgpu_comm_destroy": (None, [C.c_void_p]),
    "dgpu_comm_rank": (C.c_int, [C.c_void_p]),
    "dgpu_comm_world": (C.c_int, [C.c_void_p]),
    "dgpu_comm_allreduce_sum_i64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "dgpu_engine_exchange_topk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgpu_search_after": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float]),
    "dgpu_sharded_unique_id": (C.c_int, [C.c_void_p]),
    "dgpu_sharded_searcher_create": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "dgpu_sharded_searcher_free": (None, [C.c_void_p]),
    "dgpu_sharded_searcher_local": (C.c_void_p, [C.c_void_p]),
    "dgpu_sharded_search_batch_text": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "dgpu_sharded_search_staged": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dgpu_shm_exchange_selftest": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "dgpu_stage_compiled": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_int32]),
    "dgpu_create_long_range_query": (C.c_void_p, [C.c_char_p, C.c_int64, C.c_int64, C.c_bool, C.c_bool]),
    "dgpu_parse_query": (C.c_void_p, [C.c_char_p]),
    # ---- include/dgpu_engine.h
    "dgpu_engine_last_error": (C.c_char_p, []),
    "dgpu_engine_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "dgpu_engine_destroy": (None, [C.c_void_p]),
    "dgpu_engine_device": (C.c_int, [C.c_void_p]),
    "dgpu_engine_sm_count": (C.c_int, [C.c_void_p]),
    "dgpu_engine_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dgpu_engine_set_ktab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "dgpu_engine_decode_terms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "dgpu_engine_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(Results)]),
    "dgpu_engine_stage_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "dgpu_engine_search_staged": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dgpu_engine_device_results": (C.c_int, [C.c_void_p, C.POINTER(Results)]),
    "dgpu_engine_fetch_results": (C.c_int, [C.c_void_p, C.POINTER(Results)]),
    "dgpu_engine_sync": (C.c_int, [C.c_void_p]),
    "dgpu_engine_merge_parts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_uint32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dgpu_engine_launch_count": (C.c_uint64, [C.c_void_p]),
    "dgpu_engine_last_search_ms": (C.c_float, [C.c_void_p]),
    "dgpu_engine_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "dgpu_engine_last_phase_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float * 3)]),
    "dgpu_engine_batch_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64 * 16)]),
    "dgpu_engine_create_shadow": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "dgpu_engine_sync_options": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dgpu_engine_wait": (C.c_int, [C.c_void_p]),
    "dgpu_engine_pipeline": (None, [C.c_void_p, C.POINTER(C.c_int32 * 2)]),
}

_lib = None


def load():
    """Loads the shared library once and applies the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first (make, or __graft_entry__.build()). "
            "diagon_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    lib = load()
    msg = lib.diagon_last_error()
    return msg.decode() if msg else ""
