"""Shared helpers of the test-suite (the oracle-side comparison lives in oracle/compare.py)."""
from oracle.compare import *  # noqa: F401,F403
from oracle.compare import DEFAULTS, explain_diff, fastq_records, first_diff, frag_table, group_counts, hap_sequences, lazy_hap_sequences, oracle_jobs, oracle_run  # noqa: F401
