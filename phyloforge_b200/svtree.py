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
        packed = vcf.load_sv_vcf(self.engine(), vcf_file, keep_chars=True, rank=self.rank, world=self.world)
        if self.world == 1:
            vcf.write_fasta(out_file, packed.names, packed.chars)
        self._packed = packed
        return packed

    def sv_tree(self):
        status = self.spawn_if_multi_gpu("-S")          # gpus = N: N worker processes run this method, one per GPU
        if status is not None:
            return status
        import os
        from . import multigpu
        chief = self.rank == 0
        if chief:
            print(f"[{stamp()}] Converting a VCF file to a matrix...")
        infile = f"{self.out_path}/SV.matrix"
        t0 = time.perf_counter()
        packed, failure = None, None
        try:
            eng = self.engine()
            packed = self.convert_vcf_to_seq(self.vcf_file, infile)
        except vcf.VcfFormatError as e:
            failure = e
        if not multigpu.agree_ok(failure is None):
            if chief:
                print(f"[{stamp()}] Failed to Convert VCF file to matrix, please check input file!")
            if failure is not None:
                print(failure if self.world == 1 else f"worker {self.rank} (lines are counted from its first line): {failure}")
            sys.exit(1)
        if self.world > 1:
            # one chunk matrix per worker, concatenated per sample in worker order into SV.matrix
            wd = os.path.join(self.out_path, "working_dir")
            os.makedirs(wd, exist_ok=True)
            chunk = [os.path.join(wd, f"SV{r}.matrix") for r in range(self.world)]
            shards = multigpu.gather_local(packed.chars)   # worker threads: the shards themselves; processes: None
            if shards is None:
                vcf.write_fasta(chunk[self.rank], packed.names, packed.chars)
            multigpu.barrier()
            if chief:
                if shards is None:
                    os.replace(multigpu.merge_chunk_fastas(chunk, os.path.join(wd, "SV_all"), "fa"), infile)
                else:
                    vcf.write_fasta(infile, packed.names, shards)
            multigpu.globalise(eng, packed, self.rank, self.world)
        t1 = time.perf_counter()
        if chief:
            print(f"[{stamp()}] Finish Converting VCF file to matrix.")
        self.report.update({"vcf_file": self.vcf_file, "convert_seconds": t1 - t0,
                            "lines": getattr(packed, "total_lines", packed.n_lines),
                            "alignment_columns": getattr(packed, "total_columns", packed.n_columns),
                            "alignment_file": infile})
        if chief:
            print(f"[{stamp()}] Constructing tree, this may take a long time...")
        multigpu.barrier()
        if self.tree_software in ("iqtree", "nj"):                      # svtree.py:142-151
            status = self.built_tree(infile, self.out_path, "SV", self.thread)
            if chief:
                if status == 0:
                    print(f"[{stamp()}] Finish constructing the SV tree.")
                else:
                    print(f"[{stamp()}] SV tree failed to construct.")
            return status
        print(f"tree_software: {self.tree_software} can not be recognized, please check it. exit......")
        sys.exit()
