"""Re-export of phyloforge_b200.svtree under the reference module name treetool.svtree."""
from phyloforge_b200.svtree import *  # noqa: F401,F403
from phyloforge_b200 import svtree as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
