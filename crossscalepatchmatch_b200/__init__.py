"""crossscalepatchmatch_b200 — B200-native (sm_100a) PatchMatch-stereo +
cross-scale plane-cost engine behind the plugin surface of
rookiepig/CrossScalePatchMatch.  CUDA only; see include/cspm_c.h for the ABI.

Attribute access is lazy: `import crossscalepatchmatch_b200.synth` (the input
generator, host-only) does not load libcspm_b200.so, anything else does — and
fails loudly if the shared object is missing (there is no CPU fallback)."""
import importlib

_API = ("CCMethod", "CenCC", "CSPC", "CSPatchMatch", "GrdPC", "GetCCType", "GrdCC", "IPlaneCost", "Plane", "PreCSPC",
        "PreSSPC", "kLeft", "kRight", "set_thread_stream", "PatchMatchBatch")
_LIB = ("CspmError", "LIB_PATH", "SCHEDULE_RASTER", "SCHEDULE_REDBLACK", "SCHEDULE_REDBLACK8", "STAGE_INIT", "STAGE_PLANE2DISP", "STAGE_POSTPROC",
        "STAGE_REFINE", "STAGE_SPATIAL", "STAGE_VIEW")
__all__ = ["CspmError", "CCMethod", "GrdCC", "CenCC", "GetCCType", "IPlaneCost", "PreSSPC", "PreCSPC", "GrdPC", "CSPC", "Plane",
           "CSPatchMatch", "kLeft", "kRight", "LIB_PATH", "device_count"]


def __getattr__(name):
    if name in _API:
        return getattr(importlib.import_module(".api", __name__), name)
    if name in _LIB:
        return getattr(importlib.import_module("._lib", __name__), name)
    if name in ("api", "_lib", "shard", "synth"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def device_count() -> int:
    return importlib.import_module("._lib", __name__).dll.cspm_device_count()
