"""oswald_b200 - B200-native Smith-Waterman database search (the hot path of enzorucci/OSWALD).

Product code: the CUDA library (csrc/cuda, C ABI in include/oswald_cuda.h), the host C tool
(csrc/host) and this thin Python mirror of the reference's host interface used by the tests
and the bench.  Nothing here imports oracle/.
"""
from .host import (ALPHABET, Database, Queries, Searcher, encode, load_query_sequences,  # noqa: F401
                   matrix, matrix_names, preprocess_db, read_fasta, write_db_file, db_file_info)
