#!/usr/bin/env python
"""Build-time registration of the sm_100a platform in a PanguLU tree, the way a maintainer would do it:

    python register_platform.py <reference root> <repo root> <out dir>

Creates <out dir> (git-ignored, under oracle/_ref) as a build tree of SYMBOLIC LINKS to the reference's sources -- nothing
is copied -- plus the four files the registration touches, derived from the reference's own:
    src/platforms/platform_list.csv        + "0201001,GPU_SM100"
    build_list.csv                         CPU_NAIVE, GPU_SM100
    src/pangulu_common.h                   PANGULU_DEFAULT_PLATFORM -> PANGULU_PLATFORM_GPU_SM100 (src/pangulu_common.h:4-9)
    src/platforms/02_NONSHAREDMEM/01_GPU/001_SM100/   links to include/pangulu_platform_0201001.h (+ pangulu_sm100_abi.h)
and then runs the reference's OWN generator, `python3 build_helper.py generate_platform_helper`, inside that tree, which
writes src/pangulu_platform_helper.c and src/platforms/pangulu_platform_common.h with the GPU_SM100 cases."""
import os
import shutil
import subprocess
import sys


def main(ref, repo, out):
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(os.path.join(out, "src", "platforms"))
    os.symlink(os.path.join(ref, "include"), os.path.join(out, "include"))
    os.symlink(os.path.join(ref, "build_helper.py"), os.path.join(out, "build_helper.py"))
    derived = {"pangulu_common.h", "pangulu_platform_helper.c"}
    for f in sorted(os.listdir(os.path.join(ref, "src"))):
        p = os.path.join(ref, "src", f)
        if os.path.isfile(p) and f not in derived:
            os.symlink(p, os.path.join(out, "src", f))
    text = open(os.path.join(ref, "src", "pangulu_common.h")).read()
    old = "#define PANGULU_DEFAULT_PLATFORM PANGULU_PLATFORM_GPU_CUDA"
    if old not in text:
        raise SystemExit("default-platform line not found in pangulu_common.h")
    open(os.path.join(out, "src", "pangulu_common.h"), "w").write(text.replace(old, "#define PANGULU_DEFAULT_PLATFORM PANGULU_PLATFORM_GPU_SM100", 1))
    pdir = os.path.join(out, "src", "platforms")
    os.makedirs(os.path.join(pdir, "01_SHAREDMEM"))
    os.symlink(os.path.join(ref, "src", "platforms", "01_SHAREDMEM", "00_CPU"), os.path.join(pdir, "01_SHAREDMEM", "00_CPU"))
    mine = os.path.join(pdir, "02_NONSHAREDMEM", "01_GPU", "001_SM100")
    os.makedirs(mine)
    for h in ("pangulu_platform_0201001.h", "pangulu_sm100_abi.h"):
        os.symlink(os.path.join(repo, "include", h), os.path.join(mine, h))
    plist = open(os.path.join(ref, "src", "platforms", "platform_list.csv")).read().rstrip("\n")
    open(os.path.join(pdir, "platform_list.csv"), "w").write(plist + "\n0201001,GPU_SM100\n")
    open(os.path.join(out, "build_list.csv"), "w").write("CPU_NAIVE\nGPU_SM100\n")
    subprocess.run([sys.executable, "build_helper.py", "generate_platform_helper"], cwd=out, check=True)
    helper = open(os.path.join(out, "src", "pangulu_platform_helper.c")).read()
    if "case PANGULU_PLATFORM_GPU_SM100 : pangulu_platform_0201001_hybrid_batched" not in helper:
        raise SystemExit("the generated dispatch does not name the sm_100a platform")


if __name__ == "__main__":
    main(*sys.argv[1:4])
