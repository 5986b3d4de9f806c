"""Build libcase_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI), and the torch custom-op layer
libcase_b200_torch.so (TORCH_LIBRARY(case_b200, ...), csrc/torch_ops.cpp: host C++ only, links the C ABI + libtorch)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcase_b200.so')
TORCH_LIB = os.path.join(HERE, 'libcase_b200_torch.so')
SOURCES = ['rowops.cu', 'rowops_tc.cu', 'layer_cluster.cu', 'attention.cu', 'xattn_part.cu', 'additive_v2.cu', 'vocab.cu', 'tail.cu', 'sparse_tail.cu', 'select.cu', 'step.cu', 'gemm_tcgen05.cu', 'producers.cu', 'gemm_rows.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mt for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')] + \
           [os.path.join(HERE, '..', 'include', 'case_b200.h')]
    objs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, 'build', s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f'--- nvcc {s}\n{out}', file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-lcudart'])
    build_torch_ops(force=force, verbose=verbose)
    return LIB


def build_torch_ops(force: bool = False, verbose: bool = False) -> str:
    """g++ only (the ops call the C ABI; no device code here): seconds on the CPU box, the .so ships with the snapshot."""
    import torch
    from torch.utils import cpp_extension as ce
    src = os.path.join(CSRC, 'torch_ops.cpp')
    hdr = os.path.join(HERE, '..', 'include', 'case_b200.h')
    if not (force or _stale(TORCH_LIB, [src, hdr, LIB])):
        return TORCH_LIB
    tlib = ce.library_paths()[0]
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-o', TORCH_LIB, src,
           f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}', '-DTORCH_API_INCLUDE_EXTENSION_H']
    for inc in ce.include_paths() + [cuda_inc]:
        cmd += ['-isystem', inc]
    cmd += ['-L', tlib, '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch', '-L', HERE, '-l:libcase_b200.so',
            '-Wl,-rpath,$ORIGIN', f'-Wl,-rpath,{tlib}']
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0 or verbose:
        print(f'--- g++ torch_ops.cpp\n{p.stdout}', file=sys.stderr)
    if p.returncode != 0:
        raise RuntimeError('building the torch custom-op layer failed')
    return TORCH_LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
