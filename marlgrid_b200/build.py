"""In-tree build of the CUDA library (nvcc cross-compiles sm_100a without a GPU).

One object per translation unit of marlgrid_b200/csrc (compiled in parallel, rebuilt only when its sources
changed), linked into marlgrid_b200/libmarlgrid_b200.so.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
HEADER = os.path.join(HERE, "..", "include", "marlgrid_b200.h")
LIB = os.path.join(HERE, "libmarlgrid_b200.so")

# translation unit -> headers it includes (besides mg_device.cuh / mg_common.cuh / the C ABI header)
UNITS = {
    "mg_abi.cu": [],
    "mg_env_kernels.cu": ["mg_env.cuh"],
    "mg_obs_kernels.cu": ["mg_obs.cuh"],
    "mg_fused_kernels.cu": ["mg_env.cuh", "mg_obs.cuh"],
    "mg_fused2.cu": [],
    "mg_pregen.cu": ["mg_world.cuh"],
    "mg_fused2_enc7.cu": ["mg_env.cuh", "mg_world.cuh", "mg_fused2.cuh"],
    "mg_fused2_enc5.cu": ["mg_env.cuh", "mg_world.cuh", "mg_fused2.cuh"],
    "mg_fused2_enc7h.cu": ["mg_env.cuh", "mg_world.cuh", "mg_fused2.cuh"],
    "mg_fused2_enc5h.cu": ["mg_env.cuh", "mg_world.cuh", "mg_fused2.cuh"],
    "mg_fused2_rgb7.cu": ["mg_env.cuh", "mg_world.cuh", "mg_fused2.cuh"],
    "mg_fused2_rgb5.cu": ["mg_env.cuh", "mg_world.cuh", "mg_fused2.cuh"],
}
COMMON = ["mg_device.cuh", "mg_common.cuh"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _deps(unit):
    return [os.path.join(CSRC, unit)] + [os.path.join(CSRC, h) for h in COMMON + UNITS[unit]] + [HEADER]


def _obj(unit):
    return os.path.join(OBJDIR, unit.replace(".cu", ".o"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return any(_stale(_obj(u), _deps(u)) for u in UNITS) or _stale(LIB, [_obj(u) for u in UNITS if os.path.exists(_obj(u))])


def build(force=False, verbose=False):
    """Compile marlgrid_b200/csrc/*.cu -> marlgrid_b200/libmarlgrid_b200.so."""
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [u for u in UNITS if force or _stale(_obj(u), _deps(u))]

    def compile_unit(u):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(u), os.path.join(CSRC, u)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return u, r.returncode, r.stdout + r.stderr

    failed = False
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1)) as ex:
            for u, rc, out in ex.map(compile_unit, todo):
                if out.strip() and (verbose or rc != 0):
                    sys.stderr.write(f"--- {u}\n{out}\n")
                failed = failed or rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if todo or _stale(LIB, [_obj(u) for u in UNITS]):
        subprocess.check_call([nvcc, "-shared", "-o", LIB] + [_obj(u) for u in UNITS] + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
