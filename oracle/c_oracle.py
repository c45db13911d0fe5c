"""ctypes front-end of the C oracle (oracle/encode_ref.c). ORACLE = test infrastructure, never product code."""
import ctypes
import os

import numpy as np

from . import build as _build

_ERR = {1: 'capacity', 2: 'sub_degree >= 200', 3: 'h outside [1,4]', 4: 'rd bin outside [0,100)',
        5: 'use_rd on a non-symmetric edge multiset', 6: 'node id outside [0,N)'}
_lib = None
_i64p = ctypes.POINTER(ctypes.c_int64)


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB if os.path.exists(_build.LIB) else _build.build()
        _lib = ctypes.CDLL(path)
        _lib.escgnn_oracle_encode.restype = ctypes.c_int
        _lib.escgnn_oracle_encode_batch_digest.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_i64p)


def encode_graph(edge_index, num_nodes, h, use_rd=False, self_loop=False):
    """Same contract as oracle.encode_ref.encode_graph (rd_mode='fp64')."""
    ei = np.ascontiguousarray(np.asarray(edge_index, dtype=np.int64).reshape(2, -1))
    e_in, n = ei.shape[1], int(num_nodes)
    eo = np.empty((2, e_in + n), dtype=np.int64)
    e_out = ctypes.c_int64(0)
    nnz = ctypes.c_int64(0)
    cap = max(64, (e_in + n) * 64)
    while True:
        buf = np.empty((3, cap), dtype=np.int64)
        rc = lib().escgnn_oracle_encode(_p(ei[0]), _p(ei[1]), ctypes.c_int64(e_in), ctypes.c_int64(n),
                                        int(h), int(use_rd), int(self_loop), _p(eo[0]), _p(eo[1]),
                                        ctypes.byref(e_out), _p(buf[0]), _p(buf[1]), _p(buf[2]),
                                        ctypes.c_int64(cap), ctypes.byref(nnz))
        if rc == 1:
            cap = int(nnz.value)
            continue
        if rc != 0:
            raise ValueError('oracle: ' + _ERR.get(rc, str(rc)))
        k = int(nnz.value)
        return eo[:, :e_out.value].copy(), buf[0, :k].copy(), buf[1, :k].copy(), buf[2, :k].copy()


def encode_batch_digest(src, dst, edge_ptr, node_ptr, h, use_rd=False, self_loop=False, threads=None):
    """(E_out, nnz, sum of counts, xor-hash) over a batch of graphs with graph-local node ids. OpenMP over graphs."""
    if threads:
        os.environ['OMP_NUM_THREADS'] = str(threads)
    src = np.ascontiguousarray(src, dtype=np.int64)
    dst = np.ascontiguousarray(dst, dtype=np.int64)
    edge_ptr = np.ascontiguousarray(edge_ptr, dtype=np.int64)
    node_ptr = np.ascontiguousarray(node_ptr, dtype=np.int64)
    out = np.zeros(4, dtype=np.int64)
    rc = lib().escgnn_oracle_encode_batch_digest(_p(src), _p(dst), _p(edge_ptr), _p(node_ptr),
                                                 ctypes.c_int64(len(node_ptr) - 1), int(h), int(use_rd),
                                                 int(self_loop), _p(out))
    if rc != 0:
        raise ValueError('oracle: ' + _ERR.get(rc, str(rc)))
    return tuple(int(x) for x in out)
