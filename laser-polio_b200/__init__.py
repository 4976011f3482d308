"""laser-polio per-tick agent update, B200-native.

``import laser_polio_b200 as lp`` gives the reference's public surface for the hot path
(reference ``src/laser_polio/__init__.py``): ``lp.SEIR_ABM``, the five components,
``lp.default_pars`` / ``lp.default_run_order``, ``lp.PropertySet``, the ``lp.*`` distributions,
``lp.date`` / ``lp.daterange`` / ``lp.get_seasonality``.  The per-tick work runs as sm_100a CUDA
kernels behind the C ABI in ``include/lpk.h``; see DESIGN.md.
"""

from .abm import *  # noqa: F401,F403
from .core import *  # noqa: F401,F403
from .distributions import *  # noqa: F401,F403
from .pars import *  # noqa: F401,F403
from .utils import *  # noqa: F401,F403

__version__ = "0.1.0"
