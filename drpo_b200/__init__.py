"""Import alias: the package directory is named after the project
(``distributional-reachability-policy-optimization_b200/``), which is not a valid Python identifier, so ``drpo_b200``
points its search path there and re-exports it."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "distributional-reachability-policy-optimization_b200"))
_init = _os.path.join(__path__[0], "__init__.py")
exec(compile(open(_init).read(), _init, "exec"))
