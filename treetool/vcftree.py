"""Re-export of phyloforge_b200.vcftree under the reference module name treetool.vcftree."""
from phyloforge_b200.vcftree import *  # noqa: F401,F403
from phyloforge_b200 import vcftree as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
