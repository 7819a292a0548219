"""crossscalepatchmatch_b200 — B200-native (sm_100a) PatchMatch-stereo +
cross-scale plane-cost engine behind the plugin surface of
rookiepig/CrossScalePatchMatch.  CUDA only; see include/cspm_c.h for the ABI."""
from ._lib import (CspmError, LIB_PATH, SCHEDULE_RASTER, SCHEDULE_REDBLACK, STAGE_INIT, STAGE_PLANE2DISP,
                   STAGE_POSTPROC, STAGE_REFINE, STAGE_SPATIAL, STAGE_VIEW)
from .api import (CCMethod, CenCC, CSPC, CSPatchMatch, GrdPC, GetCCType, GrdCC, IPlaneCost, Plane, PreCSPC, PreSSPC, kLeft,
                  kRight, set_thread_stream)
from . import _lib

__all__ = ["CspmError", "CCMethod", "GrdCC", "CenCC", "GetCCType", "IPlaneCost", "PreSSPC", "PreCSPC", "GrdPC", "CSPC", "Plane",
           "CSPatchMatch", "kLeft", "kRight", "LIB_PATH"]


def device_count() -> int:
    return _lib.dll.cspm_device_count()
