"""python -m crossscalepatchmatch_b200 — the reference's command line (CSPM/main.cc:23-34, 57-139) on the B200 engine.

Same flags, gflags syntax (--name=value, --name value, --flag / --noflag for booleans):

  --l_img_file --r_img_file   input images, anything cv2.imread decodes (the reference ships PNGs)
  --l_dis_file --r_dis_file   output disparity maps (8-bit, written with cv2.imwrite)
  --max_dis --dis_scale       disparity range and output scaling
  --cc_name                   GRD | CEN                      (GetCCType, main.cc:39-55)
  --use_cs --reg_lambda       PreCSPC with scale_num = 5     (main.cc:96-104), else PreSSPC
  --use_pp                    post-processing                (cs_patchmatch.cc:347-588)
plus three that the reference does not have: --seed (RNG contract of include/cspm_c.h; default 0x5EED), --device and
--schedule=raster|redblack|redblack8 (raster = the reference's order, the default; include/cspm_c.h).

Image decoding / encoding is host plumbing outside the hot path; everything between
imread and imwrite runs in libcspm_b200.so.  No CPU fallback: without a CUDA device the run fails.
"""
from __future__ import annotations

import sys
import time

DEFAULTS = {  # main.cc:23-34
    "l_img_file": "l_img.png", "r_img_file": "r_img.png", "l_dis_file": "l_dis.png", "r_dis_file": "r_dis.png",
    "max_dis": 0, "dis_scale": 0, "cc_name": "CCName", "use_cs": False, "use_pp": False, "reg_lambda": 0.0,
    "seed": 0x5EED, "device": 0, "schedule": "raster",
}
SCHEDULES = {"raster": 0, "redblack": 1, "redblack8": 2}


def parse_flags(argv):
    """gflags-style parsing; returns the flag dict or raises ValueError with the message to print."""
    flags = dict(DEFAULTS)
    i = 0
    while i < len(argv):
        a = argv[i].lstrip("-")
        i += 1
        if "=" in a:
            k, v = a.split("=", 1)
        elif a in flags and isinstance(DEFAULTS[a], bool):
            k, v = a, "true"
        elif a.startswith("no") and a[2:] in flags and isinstance(DEFAULTS[a[2:]], bool):
            k, v = a[2:], "false"
        elif a in flags and i < len(argv):
            k, v = a, argv[i]
            i += 1
        else:
            k, v = a, None
        if k not in flags or v is None:
            raise ValueError(f"unknown command line flag '{k}'" if k not in flags else f"flag '--{k}' is missing its argument")
        d = DEFAULTS[k]
        if isinstance(d, bool):
            if v.lower() not in ("true", "false", "1", "0", "t", "f", "y", "n", "yes", "no"):
                raise ValueError(f"illegal value '{v}' specified for bool flag '{k}'")
            flags[k] = v.lower() in ("true", "1", "t", "y", "yes")
        elif isinstance(d, int):
            flags[k] = int(v, 0)
        elif isinstance(d, float):
            flags[k] = float(v)
        else:
            flags[k] = v
    return flags


def main(argv=None) -> int:
    print("PatchMatch Stereo Matching")
    try:
        F = parse_flags(sys.argv[1:] if argv is None else argv)
    except ValueError as e:
        print(f"ERROR: {e}", file=sys.stderr)
        return 1
    import cv2

    bar = "--------------------------------------------------------"
    print(f"{bar}\nLoad Image: {F['l_img_file']} {F['r_img_file']}\n{bar}")
    l_img = cv2.imread(F["l_img_file"], cv2.IMREAD_COLOR)  # 8-bit BGR, as CV_LOAD_IMAGE_COLOR (main.cc:67-68)
    r_img = cv2.imread(F["r_img_file"], cv2.IMREAD_COLOR)
    if l_img is None or r_img is None:
        print("Error: can not open image")  # main.cc:69-75
        return 1
    from . import CSPatchMatch, CspmError, GetCCType, PreCSPC, PreSSPC

    try:
        cc_cost = GetCCType(F["cc_name"])  # main.cc:89
        if cc_cost is None:
            print(f"Error: unknown cost function '{F['cc_name']}' (GRD | CEN)")
            return 1
        t0 = time.perf_counter()  # timed region, main.cc:92
        max_iter, wnd_size = 3, 35
        if F["use_cs"]:
            scale_num = 5  # main.cc:100
            plane_cost = PreCSPC(l_img, r_img, F["max_dis"], wnd_size, scale_num, cc_cost, F["reg_lambda"], device=F["device"])
        else:
            plane_cost = PreSSPC(l_img, r_img, F["max_dis"], wnd_size, cc_cost, device=F["device"])
        if F["schedule"] not in SCHEDULES:
            print(f"Error: unknown schedule '{F['schedule']}' (raster | redblack | redblack8)")
            return 1
        patch_match = CSPatchMatch(l_img, r_img, F["max_dis"], F["dis_scale"], seed=F["seed"], device=F["device"],
                                   schedule=SCHEDULES[F["schedule"]])
        patch_match.PatchMatch(max_iter, plane_cost, F["use_pp"])
        l_dis, r_dis = patch_match.dis(0), patch_match.dis(1)
        print(f"{bar}\nTotal Time: {time.perf_counter() - t0:.6g}\n{bar}")  # main.cc:119-125
    except CspmError as e:
        print(f"ERROR: {e}", file=sys.stderr)
        return 1
    ok = cv2.imwrite(F["l_dis_file"], l_dis) and cv2.imwrite(F["r_dis_file"], r_dis)  # main.cc:132-133
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
