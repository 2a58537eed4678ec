"""Build libkbnet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libkbnet_b200.so')
SOURCES = ['api.cu', 's2d_fused.cu', 's2d_generic.cu', 'conv_direct.cu', 'conv_umma.cu', 'conv_tma.cu', 'outlier.cu', 'conv_small.cu', 'conv_mma.cu', 'depth_io.cu', 'operand_format.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC', '-shared', '--threads', '0']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'kbnet_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    extra = (['-DKBNET_V2_STAMPS'] if os.environ.get('KBNET_V2_STAMPS') else []) + ['-D' + f for f in os.environ.get('KBNET_V2_DEFS', '').split() if f]   # role timelines (tools/v2_timeline.py)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
