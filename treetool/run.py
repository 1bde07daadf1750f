"""Re-export of phyloforge_b200.run under the reference module name treetool.run."""
from phyloforge_b200.run import *  # noqa: F401,F403
from phyloforge_b200 import run as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
