"""SHIM, not Biopython: the minimum of `Bio.SeqIO` that the reference's unchanged test.py uses
(test.py:2,9-10), so it can run in this image where Biopython is absent.  Test infrastructure only."""
