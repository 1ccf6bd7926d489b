"""Import shim: the package directory is named `spark-rapids_b200/` (not a valid Python identifier),
so this module turns itself into a package whose search path is that directory."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "spark-rapids_b200")]
from spark_rapids_b200._init import *  # noqa: E402,F401,F403
