"""A/B builds of libvoxeltoy_b200.so with extra -D flags, for one-call GPU comparisons:

    python tools/build_variants.py nohints:-DVT_MEM_HINTS=0 chunk12:-DVT_WF_STEP_CHUNK=12 ...

writes voxeltoy_b200/variants/<name>.so (git-ignored, travels to the GPU box); select one with VT_LIB_PATH=<path>."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxeltoy_b200 import build as vb   # noqa: E402


def main():
    out_dir = os.path.join(vb.HERE, "variants")
    os.makedirs(out_dir, exist_ok=True)
    cus = vb._sources(vb.CSRC, (".cu",)); cpps = vb._sources(vb.HOST, (".cpp",))
    procs = []
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(":")
        lib = os.path.join(out_dir, name + ".so")
        cmd = ["nvcc"] + vb.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-I", os.path.join(vb.ROOT, "include"), "-I", vb.HERE, "-o", lib] + cus + cpps + ["-lz"]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit("variant %s failed to build" % name)
        print("built", name)


if __name__ == "__main__":
    main()
