"""Build recipe for libvitsdec.so (the C-ABI library declared in include/vitsdec.h).

nvcc cross-compiles for sm_100a without a GPU; the .so is built IN-TREE next to this file so that it
travels to the GPU box with the repo snapshot.  No torch headers are involved: the library only needs
the CUDA runtime (statically linked) and resolves cuTensorMapEncodeTiled from the driver at run time.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvitsdec.so")            # the product: tcgen05 / TMA kernels only
LIB_TEST = os.path.join(HERE, "libvitsdec_test.so")  # + the CUDA-core cross-check backend (option impl=1), tests only
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["decoder.cu", "conv_tc.cu", "conv_tc2.cu", "conv_pair.cu", "conv_pairf.cu", "conv_mrfp.cu", "conv_mrf128.cu", "pack.cu", "flow.cu"]
# Test build: the same objects, except that decoder.cu is compiled with -DVITSDEC_TESTING (which is what makes impl=1
# reachable) and conv_simt.cu, the CUDA-core restatement of the conv primitive, is linked in.  The product library has
# no second backend (north_star: "no multi-backend dispatch").
TEST_ONLY_SOURCES = ["conv_simt.cu"]
TEST_RECOMPILED = ["decoder.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
]
if os.environ.get("VITSDEC_TRACE") == "1":   # debug build with per-tile clock64 trace hooks (tools/trace_probe.py)
    NVCC_FLAGS = NVCC_FLAGS + ["-DVITSDEC_TRACE=1"]


if os.environ.get("VITSDEC_WAIT_HINT_NS"):   # experiment: suspend-time hint of mbarrier.try_wait (ptx.cuh)
    NVCC_FLAGS = NVCC_FLAGS + ["-DVITSDEC_WAIT_HINT_NS=" + os.environ["VITSDEC_WAIT_HINT_NS"]]


if os.environ.get("VITSDEC_SPIN_SLEEP"):   # experiment: back-off (ns) between failed mbarrier polls
    NVCC_FLAGS = NVCC_FLAGS + ["-DVITSDEC_SPIN_SLEEP=" + os.environ["VITSDEC_SPIN_SLEEP"]]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    paths = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if not f.startswith(".")]
    paths.append(os.path.join(HERE, "..", "include", "vitsdec.h"))
    for path in paths:
        if os.path.isfile(path):
            h.update(os.path.basename(path).encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh():
    return (os.path.exists(LIB) and os.path.exists(LIB_TEST) and os.path.exists(STAMP)
            and open(STAMP).read().strip() == _digest())


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libvitsdec.so.  Returns the library path."""
    if not force and is_fresh():
        return LIB
    nvcc = _nvcc()
    objs, objs_test = [], []
    log = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []

    def compile_(src, suffix, extra):
        obj = os.path.join(HERE, "build", src.replace(".cu", suffix + ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        return obj

    for src in SOURCES:
        obj = compile_(src, "", [])
        objs.append(obj)
        if src in TEST_RECOMPILED:
            objs_test.append(compile_(src, "_test", ["-DVITSDEC_TESTING=1"]))
        else:
            objs_test.append(obj)
    for src in TEST_ONLY_SOURCES:
        objs_test.append(compile_(src, "_test", ["-DVITSDEC_TESTING=1"]))
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s: %s" % (src, " ".join(cmd)))
    for lib, obj_list in ((LIB, objs), (LIB_TEST, objs_test)):
        cmd = [nvcc, "-shared", "-o", lib] + obj_list + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if out.returncode != 0:
            sys.stderr.write(out.stdout)
            raise RuntimeError("link failed: " + " ".join(cmd))
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    with open(STAMP, "w") as f:
        f.write(_digest())
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
