from .combine import CombineMove
from .distgen import DistributionGenerate
from .gaussian import GaussianMove
from .group import GroupStretchMove
from .rj import DistributionGenerateRJ, ReversibleJumpMove
from .move import Move
from .mtdistgen import MTDistGenMove
from .stretch import StretchMove
from .tempering import TemperatureControl, make_ladder

__all__ = ["Move", "CombineMove", "StretchMove", "GaussianMove", "GroupStretchMove", "ReversibleJumpMove", "DistributionGenerateRJ", "DistributionGenerate", "MTDistGenMove",
           "TemperatureControl", "make_ladder"]
