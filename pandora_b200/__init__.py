"""pandora_b200 -- B200-native (sm_100a) implementation of CNES/Pandora's dense cost-volume hot path.

Census / SAD / SSD / ZNCC matching cost -> cross-based cost aggregation -> 8-path SGM -> winner-takes-all,
behind the reference's plugin API (``AbstractMatchingCost`` / ``AbstractAggregation`` /
``AbstractOptimization`` / ``AbstractDisparity``: same names, config keys and error behaviour) and a
C-ABI shared library (``include/pandora_b200.h``).  Importing this package does not need a GPU; any
compute call does -- there is no CPU fallback.
"""
from . import constants  # noqa: F401
from ._common import ConfigError, MachineError, get_engine  # noqa: F401
from ._native import LIB_PATH, Pb200Error, build, get_option, kernel_launches, last_path, load, option, set_option  # noqa: F401
from .aggregation import AbstractAggregation, CrossBasedCostAggregation  # noqa: F401
from .dataset import DataArray, Dataset, LazyVolume, add_disparity, create_image_dataset  # noqa: F401
from .disparity import AbstractDisparity, WinnerTakesAll  # noqa: F401
from .matching_cost import AbstractMatchingCost, Census, SadSsd, Zncc  # noqa: F401
from .optimization import AbstractOptimization, Sgm  # noqa: F401
from .pipeline import StereoPipeline, run  # noqa: F401
from .refinement import AbstractRefinement, Quadratic, Vfit  # noqa: F401
from .validation import AbstractValidation, CrossCheckingAccurate, right_disparity_fast  # noqa: F401
from .cost_volume_confidence import AbstractCostVolumeConfidence, Ambiguity, Risk  # noqa: F401
from .criteria import validity_mask  # noqa: F401
from .filter import AbstractFilter, MedianFilter  # noqa: F401

__version__ = "0.2.0"
