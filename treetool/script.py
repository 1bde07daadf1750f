"""Re-export of phyloforge_b200.script under the reference module name treetool.script."""
from phyloforge_b200.script import *  # noqa: F401,F403
from phyloforge_b200 import script as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
