"""Build libddiffpg_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libddiffpg_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
SOURCES = ["abi.cu", "actor_pack.cu", "actor_sample_fma.cu", "actor_sample_tc.cu", "actor_train_fma.cu", "actor_train_tc.cu", "actor_train_chain_tc.cu", "q_fma.cu", "q_chain_tc.cu", "q_tc.cu", "replay_gather.cu", "tc_gemm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Xcompiler", "-fPIC,-ffp-contract=off",
              "-Xptxas", "-v"]


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build_variant(out, defines):
    """Experimental A/B build: same sources with extra -D flags into another .so (see DDP_LIB_PATH)."""
    objdir = os.path.join(HERE, "build", "variant_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        subprocess.run([nvcc_path(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj],
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        objs.append(obj)
    subprocess.run([nvcc_path(), "-shared", "-cudart", "static", "-o", out, *objs], check=True)
    return out


def build_variant_of(src, out, defines):
    """A/B build that recompiles ONE source with extra -D flags and links it against the objects of the main build."""
    build(verbose=False)
    objdir = os.path.join(HERE, "build")
    obj = os.path.join(objdir, "variant_" + os.path.basename(out) + "_" + src.replace(".cu", ".o"))
    subprocess.run([nvcc_path(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj],
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    objs = [obj if s == src else os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    subprocess.run([nvcc_path(), "-shared", "-cudart", "static", "-o", out, *objs], check=True)
    return out


def build(force=False, verbose=True):
    """Compile every .cu under csrc/ into one shared library.  Skips when sources are unchanged."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc_path(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        failed |= p.returncode != 0
    with open(os.path.join(HERE, "build", "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see ddiffpg_b200/build/nvcc.log")
    link = [nvcc_path(), "-shared", "-cudart", "static", "-o", LIB, *objs]
    subprocess.run(link, check=True)
    with open(STAMP, "w") as f:
        f.write(dig)
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
