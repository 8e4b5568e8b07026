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
    cmd = ['g++', '-x', 'c++', '-std=c++17', '-O2', '-g', '-w', '-DDBG_EMU', '-include',
           os.path.join(HERE, 'cuda_emu.h'), '-shared', '-fPIC', '-o', OUT, os.path.join(CSRC, 'dbg_b200.cu')]
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force=True))
