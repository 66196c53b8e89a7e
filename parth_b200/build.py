"""Compile the CUDA extension in-tree: parth_b200/libparth_b200.so (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container too.  The .so is
git-ignored but travels to the GPU box with the repository snapshot.
"""
import glob
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libparth_b200.so")

OBJ = os.path.join(CSRC, "_obj")
NVCC_COMPILE = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]
NVCC_LINK = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    src, obj, verbose = args
    cmd = [_nvcc()] + NVCC_COMPILE + ["-I" + os.path.join(ROOT, "include"), "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    return r.returncode, " ".join(cmd), r.stdout + r.stderr


def build(force=False, verbose=False):
    """one object per .cu (compiled in parallel, only when stale), then one link"""
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((src, obj, verbose))
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for rc, cmd, out in ex.map(_compile_one, jobs):
            if rc != 0:
                raise RuntimeError("nvcc failed:\n" + cmd + "\n" + out)
            if verbose:
                print(out)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd = [_nvcc()] + NVCC_LINK + objs + ["-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


def build_example(name="quick_start"):
    """Compile a C++ consumer of the drop-in class (include/parth/parth.h) against the library."""
    src = os.path.join(ROOT, "examples", name + ".cpp")
    out = os.path.join(ROOT, "examples", name)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), src, "-o", out,
           "-L" + HERE, "-lparth_b200", "-Wl,-rpath," + HERE, "-pthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
