"""SNP tree workflow of the drop-in (mirror of treetool/vcftree.py, class VcfTree).

Same public methods and outputs as the reference:
  generate_ambiguous_code(base1, base2) -> str          vcftree.py:26-65
  cutFile(vcf_file, num, out_path) -> list[str]          vcftree.py:67-94
  convert_vcf_to_seq(vcf_file) -> str  (path of .fa)     vcftree.py:96-143
  merge_seq(input_files, out_file, format)               vcftree.py:145-171
  change_file_extension(file_list, new_extension)        vcftree.py:204-210
  snp_tree()                                             vcftree.py:212-243
but the conversion runs on the GPU (csrc/vcf_parse.cu, csrc/pack.cu) and, for
``tree_software = treebest``, so do the distance matrix and the NJ tree. ``snp_tree`` does
not need to cut the VCF into per-process chunks (the GPU parser takes the whole file), so
``working_dir`` stays empty unless ``cutFile`` / ``convert_vcf_to_seq`` are called directly.
"""
from __future__ import annotations

import os
import shutil
import sys
import time

from . import vcf
from .script import RunCmd, stamp

_HET = {"AC": "M", "AG": "R", "AT": "W", "CG": "S", "CT": "Y", "GT": "K"}


class VcfTree(RunCmd):
    def __init__(self, cfg: str | None = None, tree: str | None = "snp"):
        super().__init__(cfg, tree if cfg is not None else None)
        wd = f"{self.out_path}/working_dir"      # vcftree.py:16-24: always start from an empty dir
        if self.rank == 0 and not (self.gpus > 1 and self.world == 1):   # (the workers' rank 0 does it, not their parent)
            if os.path.isdir(wd):
                shutil.rmtree(wd)
            os.makedirs(wd)
        self.working_dir = os.path.abspath(wd)

    # ------------------------------------------------------------------ helpers
    def generate_ambiguous_code(self, base1, base2):
        """Diploid genotype -> one character, as the table in BASELINE.md §5. Host-side
        restatement for API compatibility; the GPU parser carries the same table."""
        b1, b2 = base1.upper(), base2.upper()
        ok1, ok2 = b1 in ("A", "C", "G", "T"), b2 in ("A", "C", "G", "T")
        if b1 == b2:
            if ok1:
                return b1
            return {".": "-", "*": "N"}.get(b1)      # None for anything else, like the reference
        if ok1 and ok2:
            return _HET["".join(sorted((b1, b2)))]
        if ok1 != ok2 and (b2 if ok1 else b1) == "*":
            return (b1 if ok1 else b2).lower()
        return "N"

    def cutFile(self, vcf_file, num, out_path):
        """Split the VCF body into `num` files of n // num + 1 lines, each with the #CHROM
        header (kept for callers that use it; snp_tree itself does not need it)."""
        header = None
        body = []
        with open(vcf_file, encoding="utf-8") as f:
            for line in f.read().splitlines():
                if line.startswith("#"):
                    if header is None and line.startswith("#CHROM"):
                        header = line + "\n"
                    continue
                body.append(line)
        if header is None:
            raise vcf.VcfFormatError(f"{vcf_file}: no #CHROM header")
        per = len(body) // num + 1
        stem, ext = os.path.splitext(vcf_file)
        paths = []
        for i in range(num):
            path = os.path.join(out_path, os.path.basename(stem) + str(i) + ext)
            part = body[i * per:] if i == num - 1 else body[i * per:(i + 1) * per]
            with open(path, "w", encoding="utf-8") as out:
                out.write(header)
                out.writelines(ln + "\n" for ln in part)
            paths.append(path)
        return paths

    def convert_vcf_to_seq(self, vcf_file):
        """VCF -> FASTA of IUPAC characters next to it (``<stem>.fa``), on the GPU."""
        packed = vcf.load_snp_vcf(self.engine(), vcf_file, keep_chars=True, non_biallelic="skip")
        out_file = os.path.splitext(vcf_file)[0] + ".fa"
        vcf.write_fasta(out_file, packed.names, packed.chars)
        return out_file

    def merge_seq(self, input_files, out_file, format):
        """Concatenate per-chunk FASTA files per sample id, in file order."""
        seqs: dict[str, list] = {}
        for name in input_files:
            cur = None
            with open(name) as f:
                for line in f:
                    line = line.rstrip("\r\n")
                    if line.startswith(">"):
                        cur = line[1:].split()[0] if line[1:].split() else ""
                        seqs.setdefault(cur, [])
                    elif cur is not None:
                        seqs[cur].append(line.strip())
        joined = {k: "".join(v) for k, v in seqs.items()}
        if format == "phylip":
            with open(out_file + ".phy", "w") as f:
                f.write(f"{len(joined)} {len(next(iter(joined.values())))}\n")
                for k, v in joined.items():
                    f.write(f"{k}  {v}\n")
        else:
            with open(out_file + ".fa", "w") as f:
                for k, v in joined.items():
                    f.write(f">{k}\n{v}\n")

    def change_file_extension(self, file_list, new_extension):
        return [os.path.splitext(p)[0] + new_extension for p in file_list]

    # ------------------------------------------------------------------ workflow
    def snp_tree(self):
        status = self.spawn_if_multi_gpu("-s")          # gpus = N: N worker processes run this method, one per GPU
        if status is not None:
            return status
        from . import multigpu
        chief = self.rank == 0
        if chief:
            print(f"[{stamp()}] Converting a VCF file to a base sequence...")
        basename = "SNP"
        t0 = time.perf_counter()
        packed, failure = None, None
        try:
            multigpu.trace("snp_tree entered")
            eng = self.engine()
            multigpu.trace("engine ready")
            multigpu.barrier()                           # rank 0 has emptied working_dir before anybody writes there
            multigpu.trace("first barrier passed")
            packed = vcf.load_snp_vcf(eng, self.vcf_file, keep_chars=True, non_biallelic=self.non_biallelic,
                                      rank=self.rank, world=self.world)
            multigpu.trace("byte range converted")
        except vcf.VcfFormatError as e:
            failure = e
        if not multigpu.agree_ok(failure is None):
            if chief:
                print(f"[{stamp()}] Failed to Convert VCF file to base sequence, please check the input file!")
            if failure is not None:
                print(failure if self.world == 1 else f"worker {self.rank} (lines are counted from its first line): {failure}")
            sys.exit(1)
        self._packed = packed
        use_fasta = self.tree_software.upper() in ("TREEBEST", "FASTTREE")      # vcftree.py:229-235
        if self.world == 1:
            if use_fasta:
                infile = f"{self.out_path}/all_snp.fa"
                vcf.write_fasta(infile, packed.names, packed.chars)
            else:
                infile = f"{self.out_path}/all_snp.phy"
                vcf.write_phylip(infile, packed.names, packed.chars)
                print(infile)
        else:
            # worker processes, the reference's own layout: one chunk FASTA per worker in working_dir, concatenated
            # in chunk order (convert_vcf_to_seq -> merge_seq, vcftree.py:139-171); worker threads hand their shards
            # to rank 0 in memory and working_dir stays empty, as with one GPU
            stem = os.path.splitext(os.path.basename(self.vcf_file))[0]
            chunk = [os.path.join(self.working_dir, f"{stem}{r}.fa") for r in range(self.world)]
            shards = multigpu.gather_local(packed.chars)   # worker threads: the shards themselves; processes: None
            if shards is None:
                vcf.write_fasta(chunk[self.rank], packed.names, packed.chars)
                multigpu.trace("chunk FASTA written")
            multigpu.barrier()
            infile = f"{self.out_path}/all_snp" + (".fa" if use_fasta else ".phy")
            if chief:
                if shards is None:
                    multigpu.merge_chunk_fastas(chunk, f"{self.out_path}/all_snp", "fa" if use_fasta else "phylip")
                else:                                      # the same concatenation, without reading the chunks back
                    (vcf.write_fasta if use_fasta else vcf.write_phylip)(infile, packed.names, shards)
                if not use_fasta:
                    print(infile)
            multigpu.trace("chunks merged (rank 0)")
            multigpu.globalise(eng, packed, self.rank, self.world)
            multigpu.trace("globalised")
        t1 = time.perf_counter()
        n_unfit = getattr(packed, "total_not_2bit", packed.n_not_2bit)
        if chief:
            print(f"[{stamp()}] Finish Converting a VCF file to a base sequence.")
            if n_unfit and self.tree_software in ("treebest", "nj"):
                how = {"include": "counted from their characters", "skip": "left out of the distance matrix"}
                print(f"[{stamp()}] {n_unfit} site(s) are not biallelic A/C/G/T (more than two alleles, or '*'): "
                      f"{how.get(self.non_biallelic, self.non_biallelic)} (non_biallelic = {self.non_biallelic})")
        self.report.update({"vcf_file": self.vcf_file, "convert_seconds": t1 - t0,
                            "lines": getattr(packed, "total_lines", packed.n_lines),
                            "alignment_columns": getattr(packed, "total_columns", packed.n_columns),
                            "sites_not_2bit": n_unfit, "alignment_file": infile})
        if chief:
            print(f"[{stamp()}] Constructing tree...")
        multigpu.barrier()
        status = self.built_tree(infile, f"{self.out_path}", basename, 1)
        multigpu.trace("tree built")
        if chief:
            if status == 0:
                print(f"[{stamp()}] Finish constructing the SNP tree.")
            else:
                print(f"[{stamp()}] SNP tree failed to construct.")
        return status
