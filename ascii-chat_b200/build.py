"""Build libasciichat_b200.so (hand-written sm_100a CUDA + C-ABI) in-tree with nvcc.

    python ascii-chat_b200/build.py [--force]

Output: ascii-chat_b200/lib/libasciichat_b200.so  (git-ignored, travels with gpurun snapshots).
cudart is linked statically so the library depends only on the driver (libcuda.so.1).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib", os.environ.get("ACB200_LIB_NAME", "libasciichat_b200.so"))  # experiments: other name
SOURCES = ["render_kernels.cu", "engine.cu", "dropin.cu", "grid.cu", "server.cu", "effects.cu", "rain.cu"] + ["rk_mode%d.cu" % k for k in range(8)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall,-fvisibility=hidden", "--extended-lambda",
              "-cudart", "static"]


def _newest_source_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(os.path.dirname(HERE), "include", "asciichat_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source_mtime():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(HERE, "lib", os.environ.get("ACB200_OBJ_PREFIX", "") + s.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("ACB200_EXTRA_NVCC", "").split() + ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
        if verbose and out:
            print(out)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", OUT] + objs + ["-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
