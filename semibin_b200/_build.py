"""Builds semibin_b200/libsbk.so in-tree with nvcc for sm_100a (no GPU needed).

The sha256 of every source the library is built from (csrc/*, include/sbk.h, the nvcc flags) is compiled into the
binary (`sbk_source_hash()`); `_lib.load()` recomputes it from the tree and refuses a binary that does not match,
so a stale prebuilt .so can never be the thing that is tested or benchmarked.

`python -m semibin_b200._build --probes` builds libsbk_probe.so, the same sources with -DSBK_ENABLE_PROBES: the
profiling switches (SBK_TC_PROBE, SBK_TC_CG, SBK_RERANK_BLOCK, ...) exist in that variant only; it is selected with
SBK_LIB_VARIANT=probe and is never what the tests or bench.py load by default.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
HEADER = os.path.normpath(os.path.join(HERE, '..', 'include', 'sbk.h'))
SOURCES = ['api.cu', 'knn_exact.cu', 'knn_tc.cu', 'edges.cu', 'radius.cu', 'scan.cu', 'integrate.cu', 'kmeans.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']
HASH_MARK = 'SBK_SOURCE_HASH='


def lib_path(probes=False):
    return os.path.join(HERE, 'libsbk_probe.so' if probes else 'libsbk.so')


LIB = lib_path(False)


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def source_hash(probes=False):
    """sha256 over the sorted source files (name + content), the header and the compile flags."""
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h')))
    for f in files:
        h.update(f.encode())
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    with open(HEADER, 'rb') as fh:
        h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS + (['-DSBK_ENABLE_PROBES'] if probes else [])).encode())
    return h.hexdigest()[:32]


def embedded_hash(path):
    """The hash compiled into a built library, read from the file (no dlopen); None if absent."""
    try:
        with open(path, 'rb') as fh:
            blob = fh.read()
    except OSError:
        return None
    i = blob.find(HASH_MARK.encode())
    if i < 0:
        return None
    j = i + len(HASH_MARK)
    return blob[j:j + 32].decode('ascii', 'replace')


def needs_build(probes=False):
    return embedded_hash(lib_path(probes)) != source_hash(probes)


def build(force=False, verbose=False, probes=False):
    lib = lib_path(probes)
    if not force and not needs_build(probes):
        return lib
    objdir = os.path.join(HERE, 'build_probe' if probes else 'build')
    os.makedirs(objdir, exist_ok=True)
    defs = ['-DSBK_SOURCE_HASH_STR="%s%s"' % (HASH_MARK, source_hash(probes))] + (['-DSBK_ENABLE_PROBES'] if probes else [])
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + defs + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f'nvcc failed on {src}')
    cmd = [_nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static', '-ldl', '-lrt', '-lpthread']
    subprocess.check_call(cmd)
    if embedded_hash(lib) != source_hash(probes):
        raise RuntimeError(f'{lib}: the source hash was not embedded')
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, probes='--probes' in sys.argv))
