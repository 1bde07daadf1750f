"""Randomised parity run of the GPU VCF parsers against the restated reference converters
(oracle/ref_convert.py), many shapes / chunk sizes / seeds. Not part of the test suite."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import vcfgen
from oracle import cpu as oracle, ref_convert as rc
from phyloforge_b200 import vcf
from phyloforge_b200.engine import Engine

eng = Engine(0)
seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
rng = np.random.default_rng(seed0)
t_end = time.time() + budget
n_snp = n_sv = n_mis = 0


def fasta(names, chars):
    return b"".join(b">" + n.encode() + b"\n" + bytes(chars[i]) + b"\n" for i, n in enumerate(names))


with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "f.vcf")
    while time.time() < t_end:
        seed = int(rng.integers(0, 1 << 30))
        n = int(rng.choice([1, 2, 3, 31, 32, 33, 64, 65, 200, 513]))
        m = int(rng.integers(1, 1500))
        chunk = int(rng.choice([4096, 5000, 8192, 65536, 1 << 20]))
        text, _ = vcfgen.random_snp_vcf(seed, n, m, miss=float(rng.choice([0, 0.05, 0.5])), subfield_p=float(rng.choice([0, 0.5])))
        open(path, "w").write(text)
        chunk = max(chunk, 2 * max(map(len, text.split("\n"))) + 2)     # a line must fit the chunk buffer
        try:
            names, seqs, kept = rc.snp_convert(text)
        except rc.ReferenceWouldFail:
            names = None
        if names is not None:
            packed = vcf.load_snp_vcf(eng, path, non_biallelic="skip", chunk_bytes=chunk)
            assert fasta(packed.names, packed.chars) == rc.fasta_text(names, seqs).encode(), ("snp", seed, n, m, chunk)
            n_snp += 1
            if n >= 2 and packed.n_columns:
                # every column of the character matrix as mismatch counts (2-bit columns + one pass per character state)
                packed = vcf.load_snp_vcf(eng, path, non_biallelic="include", chunk_bytes=chunk)
                counts = eng.alloc_counts(n)
                eng.pair_counts(packed.planes, n, packed.n_words, counts, word_end=(packed.n_sites + 31) // 32)
                if packed.extra_chars is not None:
                    eng.add_extra_column_counts(packed.extra_chars, n, counts)
                want = oracle.char_pair_counts(packed.chars)
                assert np.array_equal(counts.cpu().numpy().astype(np.int64), want), ("include", seed, n, m, chunk)
                n_mis += 1
        text, _ = vcfgen.random_sv_vcf(seed, n, m, miss=float(rng.choice([0, 0.08, 0.6])))
        open(path, "w").write(text)
        chunk = max(chunk, 2 * max(map(len, text.split("\n"))) + 2)
        try:
            names, seqs = rc.sv_convert(text)
        except rc.ReferenceWouldFail:
            names = None
        if names is not None:
            packed = vcf.load_sv_vcf(eng, path, chunk_bytes=chunk)
            assert fasta(packed.names, packed.chars) == rc.fasta_text(names, seqs).encode(), ("sv", seed, n, m, chunk)
            n_sv += 1
print(f"fuzz ok: {n_snp} SNP VCFs, {n_sv} SV VCFs, {n_mis} count matrices over every column (non_biallelic = include)")
