"""SV tree workflow of the drop-in (mirror of treetool/svtree.py, class SvTree).

  convert_vcf_to_seq(vcf_file, out_file)    treetool/svtree.py:14-125 — presence/absence
      matrix of DEL/INS records, one column per all-homozygous record, two otherwise
  sv_tree()                                 treetool/svtree.py:127-151
``tree_software = iqtree`` keeps the reference's behaviour (write SV.matrix, run iqtree);
``tree_software = nj`` builds the IBS distance matrix and the NJ tree on the GPU and writes
``{out}/SV.treefile``.
"""
from __future__ import annotations

import sys
import time

from . import vcf
from .script import RunCmd, stamp


class SvTree(RunCmd):
    def __init__(self, cfg: str | None = None, tree: str | None = "sv"):
        super().__init__(cfg, tree if cfg is not None else None)

    def convert_vcf_to_seq(self, vcf_file, out_file):
        packed = vcf.load_sv_vcf(self.engine(), vcf_file, keep_chars=True)
        vcf.write_fasta(out_file, packed.names, packed.chars)
        self._packed = packed
        return packed

    def sv_tree(self):
        print(f"[{stamp()}] Converting a VCF file to a matrix...")
        infile = f"{self.out_path}/SV.matrix"
        t0 = time.perf_counter()
        try:
            packed = self.convert_vcf_to_seq(self.vcf_file, infile)
        except vcf.VcfFormatError as e:
            print(f"[{stamp()}] Failed to Convert VCF file to matrix, please check input file!")
            print(e)
            sys.exit(1)
        t1 = time.perf_counter()
        print(f"[{stamp()}] Finish Converting VCF file to matrix.")
        self.report.update({"vcf_file": self.vcf_file, "convert_seconds": t1 - t0, "lines": packed.n_lines,
                            "alignment_columns": packed.n_columns, "alignment_file": infile})
        print(f"[{stamp()}] Constructing tree, this may take a long time...")
        if self.tree_software in ("iqtree", "nj"):                      # svtree.py:142-151
            status = self.built_tree(infile, self.out_path, "SV", self.thread)
            if status == 0:
                print(f"[{stamp()}] Finish constructing the SV tree.")
            else:
                print(f"[{stamp()}] SV tree failed to construct.")
            return status
        print(f"tree_software: {self.tree_software} can not be recognized, please check it. exit......")
        sys.exit()
