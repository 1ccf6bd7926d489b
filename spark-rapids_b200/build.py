"""Build libb200sql.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

Every translation unit under csrc/ is compiled with
  -gencode arch=compute_100a,code=sm_100a -lineinfo
and linked into spark-rapids_b200/lib/libb200sql.so.  The built .so is git-ignored but travels to
the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libb200sql.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-Xcudafe", "--diag_suppress=177", "-DB2_BUILD", "-split-compile", "0",
]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(root, f))
    return out


_INC = None


def _includes(path, seen):
    """transitive closure of the quoted #include files of `path` (csrc/ and include/ only)"""
    import re
    if path in seen or not os.path.exists(path):
        return
    seen.add(path)
    with open(path, "r", errors="replace") as fh:
        for m in re.finditer(r'^\s*#\s*include\s+"([^"]+)"', fh.read(), flags=re.M):
            _includes(os.path.normpath(os.path.join(os.path.dirname(path), m.group(1))), seen)


def _deps_digest(src=None):
    """digest of the headers `src` really includes (so touching vm.cuh does not rebuild join.cu) + the flags"""
    seen = set()
    _includes(src, seen)
    seen.discard(src)
    h = hashlib.sha1()
    for f in sorted(seen):
        with open(f, "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, digest, verbose):
    rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
    obj = os.path.join(OBJ, rel + ".o")
    stamp = obj + ".stamp"
    with open(src, "rb") as fh:
        key = hashlib.sha1(fh.read() + _deps_digest(src).encode() + digest.encode()).hexdigest()
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == key:
        return obj, False
    cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as fh:
        fh.write(key)
    return obj, True


def build_all(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    digest = "force%d" % os.getpid() if force else ""
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        res = list(ex.map(lambda s: _compile(s, digest, verbose), srcs))
    objs = [o for o, _ in res]
    if any(changed for _, changed in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
