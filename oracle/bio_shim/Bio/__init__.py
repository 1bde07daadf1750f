"""TEST INFRASTRUCTURE ONLY — a stand-in for Biopython, which is not installable here.

The unmodified reference imports ``Bio.SeqIO`` and ``Bio.BiopythonWarning``
(treetool/script.py:12-13, treetool/vcftree.py:8). On the population-phylogeny path the
only call is ``SeqIO.parse(path, "fasta")`` (treetool/vcftree.py:149). This package exists
so oracle/run_reference.py can import and run the reference's converters as they are; it
is never imported by the product.
"""


class BiopythonWarning(Warning):
    pass
