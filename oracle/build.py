"""Build the C oracle (oracle/encode_ref.c) into oracle/_build/libescgnn_oracle.so with plain gcc.

ORACLE = test infrastructure. `__graft_entry__.build()` calls this; building the checker is not using it.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libescgnn_oracle.so')


def build(force=False):
    src = os.path.join(HERE, 'encode_ref.c')
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    cmd = ['gcc', '-O2', '-fopenmp', '-shared', '-fPIC', '-o', LIB, src, '-lm']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
