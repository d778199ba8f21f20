"""Build integration/pointops_C_shim.cpp into integration/_build/pointops_C_shim.so (a `pointops._C` replacement over
libpointops_b200.so) with a plain g++ command line against the torch / pybind11 / CUDA headers -- no ninja, no JIT cache
under ~/.cache, so the built module travels with the tree like the library itself.

    python integration/build_shim.py          -> path of the .so
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT_DIR = os.path.join(HERE, "_build")
NAME = "pointops_C_shim"
OUT = os.path.join(OUT_DIR, NAME + ".so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "pointops_C_shim.cpp")
    deps = [src, os.path.join(ROOT, "include", "pointops_b200.h")]   # the shim is compiled against the C header
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(d) for d in deps):
        return OUT
    import torch
    from torch.utils import cpp_extension as ce
    sys.path.insert(0, ROOT)
    from pointcloudpdf_b200 import build as libbuild
    libbuild.build()
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = ce.include_paths("cuda") if hasattr(ce, "include_paths") else ce.include_paths()
    try:
        inc = ce.include_paths(device_type="cuda")
    except TypeError:
        pass
    cuda_home = ce.CUDA_HOME or "/usr/local/cuda"
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *[f"-I{p}" for p in inc], f"-I{os.path.join(cuda_home, 'include')}", f"-I{sysconfig.get_paths()['include']}",
           f"-I{os.path.join(ROOT, 'include')}", src, "-o", OUT,
           f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
           f"-L{libbuild.LIB_DIR}", "-lpointops_b200", f"-Wl,-rpath,{tlib}", "-Wl,-rpath,$ORIGIN/../../pointcloudpdf_b200/lib"]
    subprocess.check_call(cmd)
    return OUT


def load():
    """Import the built module (building it first if needed)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    path = build()
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
