"""Build the CPU-emulated copy of the CUDA library  --  TEST INFRASTRUCTURE ONLY.

Compiles py-debruijn_b200/csrc/dbg_b200.cu (host code AND kernel bodies) with g++ against
tests/emu/cuda_emu.h into tests/emu/_build/libdbg_b200_emu.so.  Only tests load it.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'py-debruijn_b200', 'csrc')
OUT = os.path.join(HERE, '_build', 'libdbg_b200_emu.so')


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, 'cuda_emu.h'),
                                                                os.path.join(ROOT, 'include', 'dbg_b200.h')]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = '%s.%d.tmp' % (OUT, os.getpid())             # concurrent builders (test workers) never see a partial file
    cmd = ['g++', '-x', 'c++', '-std=c++17', '-O2', '-g', '-w', '-DDBG_EMU', '-include',
           os.path.join(HERE, 'cuda_emu.h'), '-shared', '-fPIC', '-o', tmp, os.path.join(CSRC, 'dbg_b200.cu')]
    subprocess.check_call(cmd)
    os.replace(tmp, OUT)
    return OUT


if __name__ == '__main__':
    print(build(force=True))
