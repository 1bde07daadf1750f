"""A `treebest`-named command for the unmodified PhyloForge (SURVEY.md §8(b) "alternative lower
boundary", §8(f).4): the reference runs ``{treebest} nj -b 1000 {out}/all_snp.fa > SNP_treebest.NHX``
through the shell (treetool/script.py:359-360), with the executable taken from cfg/software.ini. Pointing
that entry at ``bin/treebest`` of this repository makes the upstream package use the GPU for its tree
step without importing anything else from here.

    treebest nj [-b N] [-d p|ibs] [--non-biallelic include|skip|error] <alignment.fa>     -> Newick on stdout

Defaults are those of the slot it fills: the p-distance (mismatches / columns where both sequences are known:
what a mismatch model sees on all_snp.fa) over EVERY column of the file - columns with more
than two alleles or '*'-derived characters, which the reference's converter keeps (treetool/vcftree.py:109,
126-127, 56-63), are counted from their characters. ``-b`` (bootstrap replicates) is accepted and ignored: the
reference passes no seed, so treebest's support values are not a reproducible target; the tree
printed is the NJ tree of the un-resampled alignment."""
from __future__ import annotations

import argparse
import sys


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] != "nj":
        sys.stderr.write("phyloforge_b200 treebest shim: only `treebest nj [-b N] <fasta>` is provided\n")
        return 2
    ap = argparse.ArgumentParser(prog="treebest nj")
    ap.add_argument("-b", type=int, default=0, help="bootstrap replicates (accepted, ignored)")
    ap.add_argument("-d", default="p", choices=["p", "ibs"], help="distance (default: p)")
    ap.add_argument("--non-biallelic", default="include", choices=["include", "skip", "error"],
                    help="columns that do not fit the 2-bit genotype code: counted from their characters (default), "
                         "left out, or refused")
    ap.add_argument("--skip-non-biallelic", action="store_true", help="same as --non-biallelic skip")
    ap.add_argument("alignment")
    args = ap.parse_args(argv[1:])
    from .alignment import load_alignment
    from .engine import Engine
    eng = Engine()
    policy = "skip" if args.skip_non_biallelic else args.non_biallelic
    packed = load_alignment(eng, args.alignment, non_biallelic=policy)
    if packed.n_samples < 2:
        sys.stderr.write("at least two sequences are needed\n")
        return 1
    if packed.n_not_2bit:
        sys.stderr.write(f"{packed.n_not_2bit} column(s) do not fit the 2-bit genotype code: "
                         f"{'counted from their characters' if policy == 'include' else 'left out'}\n")
    res = eng.tree_from_planes(packed.planes, packed.n_samples, packed.n_words, packed.names, metric=args.d,
                               n_sites=packed.n_sites, extra_chars=packed.extra_chars)
    sys.stdout.write(res.newick + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
