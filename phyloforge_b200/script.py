"""Configuration, tool registry and the tree-building step of the drop-in.

Mirror of the slice of the reference's ``RunCmd`` that the population-phylogeny path uses
(SURVEY.md §8(a) rows a1, a10):
  __init__ / get_config / get_soft_path   treetool/script.py:21-113
  run_command                             treetool/script.py:224-231
  built_tree                              treetool/script.py:345-364
Same attribute names, INI sections/keys and exit behaviour. The difference is where the
arithmetic runs: ``tree_software = treebest`` (SNP) and ``tree_software = nj`` (SV) build the
distance matrix and the neighbor-joining tree on the GPU instead of shelling out.
"""
from __future__ import annotations

import datetime
import json
import os
import subprocess
import sys
import time
from configparser import ConfigParser

from . import run

_CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")
_SECTIONS = {"snp": "snp_opt", "sv": "sv_opt"}
TREE_TOOLS = ("raxml", "iqtree", "fasttree", "treebest", "phyml", "nj")


def stamp() -> str:
    return datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S")


class RunCmd:
    """Holds the run options as attributes (``vcf_file``, ``out_path``, ``tree_software``,
    ``thread`` — all strings, as the reference leaves them) and the external tool paths."""

    def __init__(self, cfg: str | None = None, tree: str | None = None):
        self.aln_software = "mafft"
        self.tree_software = "raxml"            # script.py:23
        self.seq = "cds"
        self.thread = 10                        # script.py:25
        self.out_path = os.getcwd()
        # new optional keys; the defaults keep a configuration written for PhyloForge running unchanged
        self.distance = "p"                     # p = mismatches / valid columns: what a mismatch model sees on
                                                # all_snp.fa, the metric of the treebest slot; or ibs
        self.non_biallelic = "include"          # the reference keeps multi-allelic / '*' sites (vcftree.py:109,
                                                # 126-127, 56-63): they are counted from their characters; or skip / error
        self.bootstrap = 0                      # block-bootstrap replicates (the reference asks treebest for
        self.bootstrap_seed = 1                 # -b 1000, script.py:360); 0 = no support values
        self.gpus = 1                           # workers / GPUs the site axis of the VCF is cut over
        self.gpu_workers = "threads"            # one worker thread per GPU in this process; or processes
        from . import multigpu
        self.rank, self.world = multigpu.worker_env()
        self._engine = None
        self._packed = None
        self.report = {}

        if cfg is None or tree is None:         # the reference always reads sys.argv (script.py:31)
            cfg, tree = run.get_parser()
        if not os.path.isfile(cfg):             # script.py:32-35
            print(f"{cfg} is not a valid file, please check it")
            print("Exiting...")
            sys.exit(1)
        self.cfg_path, self.tree_kind = cfg, tree

        self.software_path = self.get_soft_path()
        for k, v in self.software_path.items():
            setattr(self, str(k).lower(), v)

        if tree not in _SECTIONS:
            print(f"The {tree} workflow is not part of phyloforge_b200")
            sys.exit(2)
        if tree == "sv":
            self.tree_software = "iqtree"       # script.py:50-52
        self.opt = self.get_config(cfg, _SECTIONS[tree])
        for k, v in self.opt.items():           # script.py:57-61
            setattr(self, str(k), v.lower() if "software" in k else v)

        os.makedirs(f"{self.out_path}", exist_ok=True)      # script.py:63-67
        self.out_path = os.path.abspath(f"{self.out_path}")

        try:                                    # script.py:69-74
            if int(self.thread) <= 1:
                self.thread = 2
        except ValueError:
            print("thread should be int, please check it")
            sys.exit()
        if str(self.tree_software) not in TREE_TOOLS:       # script.py:75-78
            print(f"The tree_software: {self.tree_software} is not recognized, please check it")
            sys.exit()
        self.distance = str(self.distance).strip().lower()
        if self.distance not in ("ibs", "p"):
            print(f"distance: {self.distance} is not recognized (ibs or p), please check it")
            sys.exit()
        self.non_biallelic = str(self.non_biallelic).strip().lower()
        if self.non_biallelic not in ("error", "skip", "include"):
            print(f"non_biallelic: {self.non_biallelic} is not recognized (error, skip or include), please check it")
            sys.exit()
        try:
            self.bootstrap = int(self.bootstrap)
            self.bootstrap_seed = int(self.bootstrap_seed)
            if self.bootstrap < 0:
                raise ValueError
        except ValueError:
            print(f"bootstrap: {self.bootstrap} / bootstrap_seed: {self.bootstrap_seed} must be non-negative integers")
            sys.exit()
        try:
            if str(self.gpus).strip().lower() in ("all", "auto"):
                import torch
                self.gpus = max(1, torch.cuda.device_count())
            self.gpus = int(self.gpus)
            if self.gpus < 1:
                raise ValueError
        except ValueError:
            print(f"gpus: {self.gpus} must be a positive integer (or all)")
            sys.exit()
        self.gpu_workers = str(self.gpu_workers).strip().lower()
        if self.gpu_workers not in ("threads", "processes"):
            print(f"gpu_workers: {self.gpu_workers} is not recognized (threads or processes), please check it")
            sys.exit()

    # ------------------------------------------------------------------ config
    def get_config(self, opt_cfg, opt):
        parser = ConfigParser()
        parser.read(opt_cfg)
        return dict(parser.items(opt))

    def get_soft_path(self):
        parser = ConfigParser()
        parser.read(os.path.join(_CFG_DIR, "software.ini"))
        return dict(parser.items("software"))

    # ------------------------------------------------------------------ engine
    def engine(self):
        """The GPU engine, created on first use. Raises if the CUDA library or a GPU is missing."""
        if self._engine is None:
            from .engine import Engine
            device = None
            if self.world > 1:
                from . import multigpu
                device = multigpu.init_worker(self.rank, self.world)
            self._engine = Engine(device)
        return self._engine

    def spawn_if_multi_gpu(self, flag: str):
        """``gpus = N`` (N > 1) in the parent: run this workflow as N workers, one per GPU - threads of this process
        (default) or processes of the same command line (``gpu_workers = processes``), phyloforge_b200/multigpu.py -
        and return their exit status; None when the caller is a worker itself and should do the work."""
        if self.gpus <= 1 or self.world > 1:
            return None
        from . import multigpu
        if self.gpu_workers == "threads":
            cls, cfg, kind = type(self), self.cfg_path, self.tree_kind
            method = {"-s": "snp_tree", "-S": "sv_tree"}[flag]
            status = multigpu.run_threads(lambda: getattr(cls(cfg, kind), method)(), self.gpus,
                                          will_reduce=self.tree_software in ("treebest", "nj"))
        else:
            status = multigpu.spawn_workers([flag, self.cfg_path], self.gpus)
        if status != 0:
            sys.exit(status)
        return status

    # ------------------------------------------------------------------ shell
    def run_command(self, command):
        try:
            subprocess.check_output(command, shell=True, stderr=subprocess.DEVNULL)
            return 0
        except subprocess.CalledProcessError as e:
            print(e)
            sys.exit(e.returncode)

    # ------------------------------------------------------------------ tree
    def gpu_tree(self, packed, out_file: str, basename: str) -> int:
        """Distance matrix + NJ on the GPU from packed genotypes; writes the tree to
        `out_file`, the distance matrix to ``{out}/{basename}.dist`` and a timing record. With several workers
        (gpus = N) every worker counts its own columns, the count matrices are summed over the workers inside
        the engine, and rank 0 writes the files."""
        import torch
        from .phylip import write_distance_matrix
        eng = self.engine()
        if packed.n_samples < 2:
            print("At least two samples are needed to build a tree")
            sys.exit(1)
        t0 = time.perf_counter()
        extra = getattr(packed, "extra_chars", None)
        if self.bootstrap > 0 and packed.n_samples >= 4:
            # main tree + seeded block-bootstrap support on its internal edges (labels in percent)
            res, support, _ = eng.bootstrap_tree(packed.planes, packed.n_samples, packed.n_words, packed.names,
                                                 replicates=self.bootstrap, seed=self.bootstrap_seed,
                                                 metric=self.distance, n_sites=packed.n_sites, extra_chars=extra,
                                                 word0=getattr(packed, "word0", 0),
                                                 total_cols=getattr(packed, "total_cols", None),
                                                 extra0=getattr(packed, "extra0", 0),
                                                 total_extra=getattr(packed, "total_extra", None))
            self.report.update({"bootstrap_replicates": self.bootstrap, "bootstrap_seed": self.bootstrap_seed,
                                "bootstrap_mean_support": float(support.mean()) if support.size else None})
        else:
            res = eng.tree_from_planes(packed.planes, packed.n_samples, packed.n_words, packed.names,
                                       metric=self.distance, n_sites=packed.n_sites, extra_chars=extra)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if self.rank != 0:
            return 0
        with open(out_file, "w") as f:
            f.write(res.newick + "\n")
        out_dir = os.path.dirname(out_file)
        write_distance_matrix(os.path.join(out_dir, f"{basename}.dist"), packed.names, res.dist.cpu().numpy())
        if res.n_zero_valid:
            print(f"[{stamp()}] Warning: {res.n_zero_valid // 2} sample pairs share no genotyped site; "
                  "their distance was set to 1.0")
        self.report.update({"samples": packed.n_samples, "sites": getattr(packed, "total_cols", packed.n_sites),
                            "distance": self.distance, "gpus": self.world, "tree_seconds": t1 - t0,
                            "tree_file": out_file})
        with open(os.path.join(out_dir, f"{basename}.phyloforge_b200.json"), "w") as f:
            json.dump(self.report, f, indent=1)
        return 0

    def built_tree(self, infile, outpath, basename, thread):
        """Construct the tree from the alignment file the converter wrote (script.py:345-364).
        treebest (SNP) / nj (SV): on the GPU from the genotypes packed by the converter.
        Other tools: the reference's own command lines, run through the shell."""
        if os.path.isfile(infile) and os.path.getsize(infile) == 0:
            print(f"{infile} is empty: no site survived the conversion, no tree is built")
            return 1
        tool = self.tree_software
        if tool in ("treebest", "nj"):
            if self._packed is None:
                # called on a file alone (the reference's contract): re-read the character matrix
                from .alignment import load_alignment
                from .vcf import VcfFormatError
                try:
                    self._packed = load_alignment(self.engine(), infile, non_biallelic=self.non_biallelic)
                except VcfFormatError as e:
                    print(e)
                    sys.exit(1)
            out = f"{outpath}/{basename}_treebest.NHX" if tool == "treebest" else f"{outpath}/{basename}.treefile"
            return self.gpu_tree(self._packed, out, basename)
        if self.rank != 0:
            return 0                            # the external tool runs once, on the files rank 0 wrote
        if tool == "iqtree":
            cmd = f"{self.iqtree} -s {infile} -pre {outpath}/{basename} -nt {thread} -m MFP --quiet -B 1000 -redo"
        elif tool == "fasttree":
            cmd = f"{self.fasttree} {infile} > {outpath}/{basename}.tre"
        elif tool == "raxml":
            model = "GTRGAMMA" if self.seq in ("cds", "codon", "codon1", "codon2") else "PROTGAMMAJTT"
            cmd = (f"{self.raxml} -f a -x 12345 -p 12345 -# 100 -m {model} -s {infile} -n {basename} "
                   f"-T {thread} -w {outpath}")
        else:  # phyml
            cmd = f"{self.phyml} -i {infile} -b 100 -m HKY85 -f m -v e -a e -o tlr"
        return self.run_command(cmd)
