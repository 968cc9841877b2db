"""CPU oracle for the RunbookAI vector-search path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  PARITY UNPINNED: see the header of rbk_oracle.c.
"""
from .oracle import (  # noqa: F401
    build,
    cosine,
    find_most_similar,
    host_threads,
    merge_lists,
    rrf,
    scores,
    search,
    search_batch_mt,
    search_batch_verify,
    search_chunked,
)
