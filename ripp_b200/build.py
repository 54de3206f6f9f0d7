"""In-tree build of libripp_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libripp_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-shared", "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return out


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def build_lib(force=False, verbose=False):
    """Compile every .cu to an object in parallel, then link the shared library."""
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    nvcc = os.environ.get("NVCC", "nvcc")
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("RIPP_B200_EXTRA_FLAGS", "").split()
    objdir = os.path.join(CSRC, os.environ.get("RIPP_B200_OBJDIR", "build"))
    out_lib = os.environ.get("RIPP_B200_OUT", LIB)
    os.makedirs(objdir, exist_ok=True)

    def obj_stale(obj, dep):
        # per-object staleness from the dependency file nvcc wrote with the object (-MD): only the translation
        # units that include an edited header are recompiled (pairing6.cu alone takes minutes)
        if force or not os.path.exists(obj) or not os.path.exists(dep):
            return True
        t = os.path.getmtime(obj)
        words = open(dep).read().replace("\\\n", " ").split()
        files = [w for w in words[1:] if not w.endswith(":")]
        return any((not os.path.exists(f)) or os.path.getmtime(f) > t for f in files)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        dep = obj[:-2] + ".d"
        if not obj_stale(obj, dep):
            return obj
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-MD", "-MF", dep, "-c", "-o", obj, src]
        print("[ripp_b200.build]", " ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", out_lib] + objs
    print("[ripp_b200.build]", " ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out_lib


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
