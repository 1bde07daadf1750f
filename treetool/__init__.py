"""Drop-in namespace: ``treetool.run``, ``treetool.script``, ``treetool.vcftree`` and
``treetool.svtree`` with the names PhyloForge users import, backed by phyloforge_b200.
Only the population-phylogeny path (-s / -S) exists here."""
