"""Build the native libraries in-tree (no JIT cache): nvcc for the CUDA C-ABI library, gcc for the host
world generator.  Used by __graft_entry__.build() and importable on a box without a GPU (nvcc
cross-compiles sm_100a)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("VOXPLAT_B200_LIB", os.path.join(HERE, "libvoxplat_b200.so"))
WORLDGEN = os.path.join(HERE, "libvpworldgen.so")
CU_SOURCES = ["vp_context.cu", "vp_splat.cu", "vp_mesh.cu", "vp_rle.cu", "vp_nodes.cu", "vp_edit.cu", "vp_worldfile.cu", "vp_worldgen_dev.cu", "vp_multi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "static"]
NVCC_FLAGS += os.environ.get("VP_NVCC_EXTRA", "").split()        # e.g. -DVP_PROFILE_PHASES for an instrumented build


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "voxplat_b200.h"))
    if force or _newer(LIB, deps):
        objs = []
        procs = []
        for s in srcs:
            o = s[:-3] + ".o"
            objs.append(o)
            cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        for cmd, p in procs:
            out, _ = p.communicate()
            if verbose or p.returncode:
                sys.stderr.write(out)
            if p.returncode:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
        cmd = ["nvcc", "-shared", "--cudart", "static", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs
        subprocess.check_call(cmd)
    wsrc = os.path.join(CSRC, "vp_worldgen.c")
    if force or _newer(WORLDGEN, [wsrc]):
        subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-fopenmp", "-fPIC", "-Wall", "-shared", "-o", WORLDGEN, wsrc])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
