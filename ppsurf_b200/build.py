"""Builds ``libppsurf_b200.so`` in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libppsurf_b200.so')
SOURCES = ['common.cu', 'knn.cu', 'linear.cu', 'decode.cu', 'decode_tc.cu', 'pointnet_tc.cu', 'chain_tc.cu', 'encoder.cu', 'fka_tc.cu', 'sampling.cu', 'volume.cu', 'mcubes.cu', 'train_gemm.cu', 'train_gemm_tc.cu',
           'train_ops.cu']
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Wno-deprecated-gpu-targets']


def _nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: ppsurf_b200 needs the CUDA toolkit to build its sm_100a kernels')
    return nvcc


def source_hash():
    """first 8 bytes (as an int) of the sha256 over the CUDA sources and the public header, in a fixed order"""
    import hashlib
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h')))
    for path in [os.path.join(CSRC, f) for f in files] + [os.path.join(os.path.dirname(HERE), 'include', 'ppsurf_b200.h')]:
        h.update(os.path.basename(path).encode())
        with open(path, 'rb') as f:
            h.update(f.read())
    return int.from_bytes(h.digest()[:8], 'little')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    obj_dir = os.path.join(HERE, 'build')
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'ppsurf_b200.h'))
    objs, procs = [], []
    digest = source_hash()
    stamp = os.path.join(obj_dir, 'source_hash.txt')
    old_digest = open(stamp).read().strip() if os.path.exists(stamp) else ''
    for src in SOURCES:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        objs.append(obj)
        # common.cu carries the digest of ALL sources (pps_source_hash): it is recompiled whenever any source changed
        extra = ['-DPPS_SOURCE_HASH={}ull'.format(digest)] if src == 'common.cu' else []
        if force or _stale(obj, [src_path] + headers) or (extra and old_digest != str(digest)):
            cmd = [nvcc] + ARCH + FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', src_path, '-o', obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('--- nvcc {}\n{}\n'.format(src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    with open(stamp, 'w') as f:
        f.write(str(digest))
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [nvcc] + ARCH + ['-shared', '-Wno-deprecated-gpu-targets', '-o', LIB_PATH] + objs
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
