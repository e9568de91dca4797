"""Builds libmmduet_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("MMD_LIB_PATH") or os.path.join(HERE, "libmmduet_b200.so")   # override: debug builds (tools/trace_attn.py)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"] + os.environ.get("MMD_NVCC_EXTRA", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mmduet_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    objdir = os.path.join(HERE, "build" + ("_dbg" if os.environ.get("MMD_LIB_PATH") else ""))
    os.makedirs(objdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


JITTER_LIB = os.path.join(HERE, "libmmduet_b200_jitter.so")


def build_jitter(force=False):
    """Diagnostic twin of the library: the attention kernel compiled with -DMMD_ATTN_JITTER=7 (pseudo-random sleeps in the
    loader, MMA-issuer and softmax roles), every other object shared with the product build.  Only the GPU test
    tests/test_gpu_kernels.py::test_attention_under_timing_jitter loads it (through MMD_LIB_PATH, in a subprocess)."""
    build(force=False)
    src = os.path.join(CSRC, "attn_tcgen05.cu")
    if not force and os.path.exists(JITTER_LIB) and os.path.getmtime(JITTER_LIB) >= max(os.path.getmtime(LIB), os.path.getmtime(src)):
        return JITTER_LIB
    objdir = os.path.join(HERE, "build")
    jobj = os.path.join(objdir, "attn_tcgen05_jitter.o")
    subprocess.check_call([NVCC] + FLAGS + ["-DMMD_ATTN_JITTER=7", "-c", src, "-o", jobj])
    objs = [os.path.join(objdir, os.path.basename(x)[:-3] + ".o") for x in sources() if not x.endswith("attn_tcgen05.cu")] + [jobj]
    subprocess.check_call([NVCC, "-shared", "-o", JITTER_LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return JITTER_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
