"""Command line of the drop-in: ``phyloforge -s <cfg>`` / ``phyloforge -S <cfg>``.

Mirror of the reference's dispatcher for the population-phylogeny path only
(treetool/run.py:12-99 ``get_parser``; treetool/run.py:102-123 ``main``): same flags, same
return value ``[cfg_path, tree_kind]``, same lazy import of the workflow module. The flags of
the reference's other workflows (-l -o -w -g) are accepted and answered with a message: they
are outside the scope of this package.
"""
from __future__ import annotations

import argparse
import configparser
import os
import sys

VERSION = "phyloforge 1.0.0 (phyloforge_b200 0.1.0)"
_CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")
_OUT_OF_SCOPE = {"lcn_tree": "lcn", "organelle_tree": "organelle", "whole_genome_tree": "whole_genome",
                 "gene_tree": "gene"}


def _cat(name: str) -> None:
    with open(os.path.join(_CFG_DIR, name)) as f:
        sys.stdout.write(f.read())


def _update_software_ini(user_cfg: str) -> None:
    """``-sc file``: merge the [software] section of `file` into cfg/software.ini (run.py:74-92)."""
    new = configparser.ConfigParser()
    new.read(user_cfg)
    if not new.has_section("software"):
        print("Invalid config file format")
        return
    target = os.path.join(_CFG_DIR, "software.ini")
    cur = configparser.ConfigParser()
    cur.read(target)
    for section in new.sections():
        if not cur.has_section(section):
            cur.add_section(section)
        for key, value in new.items(section):
            cur.set(section, key, value)
    with open(target, "w") as f:
        cur.write(f)
    print("Software config file has been updated")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(
        description="PhyloForge population-phylogeny path on B200 GPUs (drop-in for phyloforge -s / -S)",
        formatter_class=argparse.RawTextHelpFormatter)
    p.add_argument("-l", "--lcn_bs", type=str, dest="lcn_tree", help="(not part of this package)")
    p.add_argument("-s", "--snp", type=str, dest="snp_tree", help="construct phylogenetic tree based on SNP sites")
    p.add_argument("-S", "--sv", type=str, dest="sv_tree",
                   help="construct phylogenetic tree based on SV (structural variation)")
    p.add_argument("-o", "--organelle", type=str, dest="organelle_tree", help="(not part of this package)")
    p.add_argument("-w", "--whole_genome", type=str, dest="whole_genome_tree", help="(not part of this package)")
    p.add_argument("-g", "--gene", type=str, dest="gene_tree", help="(not part of this package)")
    p.add_argument("-c", action="store_true", help="get the run configuration file: phyloforge -c >run.cfg")
    p.add_argument("-gsc", action="store_true", help="get the software configuration file")
    p.add_argument("-sc", help="perform software configuration: phyloforge -sc software.cfg")
    p.add_argument("-v", "--version", action="version", version=VERSION)
    return p


def get_parser(argv=None):
    """Parse the command line and return ``[cfg_path, tree_kind]`` like the reference."""
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.c:
        _cat("run.config")
        sys.exit()
    if args.gsc:
        _cat("software.cfg")
        sys.exit()
    if args.snp_tree:
        return [args.snp_tree, "snp"]
    if args.sv_tree:
        return [args.sv_tree, "sv"]
    for attr, kind in _OUT_OF_SCOPE.items():
        if getattr(args, attr):
            return [getattr(args, attr), kind]
    if args.sc:
        _update_software_ini(args.sc)
        sys.exit()
    print(parser.format_usage())
    sys.exit()


def main(argv=None):
    cfg, tree = get_parser(argv)
    try:
        if tree == "snp":
            from . import vcftree
            vcftree.VcfTree(cfg).snp_tree()
        elif tree == "sv":
            from . import svtree
            svtree.SvTree(cfg).sv_tree()
        else:
            print(f"The {tree} workflow is not part of phyloforge_b200 (only -s and -S are); use PhyloForge itself.")
            sys.exit(2)
    finally:
        _leave_process_group()


def _leave_process_group():
    """A worker of a `gpus = N` run (phyloforge_b200/multigpu.py) leaves the workers' process group on its way out."""
    if "torch" not in sys.modules:
        return
    import torch.distributed as dist
    from . import multigpu
    if multigpu.worker_env()[1] > 1 and dist.is_available() and dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001 - the run's own exit status matters, not the teardown's
            pass


if __name__ == "__main__":
    main()
