"""diagon_b200 — B200-native BM25 query scoring behind Diagon's IndexSearcher surface.

The product is libdiagon_b200.so (CUDA engine + C++20 host layer, C ABI in include/). This package is the
Python mirror of the reference's search API used by tests and bench.py; it contains no compute.
"""
from .api import (  # noqa: F401
    BatchResult, BatchTicket, BooleanClause, BooleanQuery, DiagonError, IndexBuilder, IndexReader, IndexSearcher,
    NumericRangeQuery, Occur, Query, ScoreDoc, ShardedSearcher, Term, TermQuery, TopDocs, TotalHits, and_query, named_corpus,
    or_query, parse_line, query_log_text, write_synthetic_dump,
)
from .dumpfile import read_dump  # noqa: F401
